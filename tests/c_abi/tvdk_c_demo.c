/* tvdk_c_demo.c -- the C ABI used from plain C, without Python or PyTorch: what a non-Python host (or a
 * DarSIA maintainer binding libtvdk through cffi / ctypes / a C extension) has to do for one call of
 * darsia.split_bregman_tvd (restoration/split_bregman_tvd.py:14-177).
 *
 *   tvdk_c_demo <in.bin> <out.bin> <n0> <n1> <isotropic> <solver: 0 Jacobi / 1 CG> <solver_maxiter> <outer iters>
 *
 * in.bin / out.bin: n0*n1 float64 values, row-major.  mu = 0.1, omega = 1, ell = 1.
 * Build: gcc tvdk_c_demo.c -I../../include -I$CUDA/include -L../../darsia_b200/csrc -ltvdk -L$CUDA/lib64 -lcudart
 * (tests/test_gpu_c_abi.py builds and runs it on the GPU box and compares with the oracle). */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tvdk.h"

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                       \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)
#define TK(call)                                                                        \
  do {                                                                                  \
    int r_ = (call);                                                                    \
    if (r_ != 0) {                                                                      \
      fprintf(stderr, "%s -> %d: %s\n", #call, r_, tvdk_last_error());                  \
      return 3;                                                                         \
    }                                                                                   \
  } while (0)

int main(int argc, char** argv) {
  if (argc != 9) {
    fprintf(stderr, "usage: %s in.bin out.bin n0 n1 isotropic solver solver_maxiter outer_iters\n", argv[0]);
    return 1;
  }
  const int n0 = atoi(argv[3]), n1 = atoi(argv[4]);
  const size_t nvox = (size_t)n0 * n1;
  double* host = (double*)malloc(nvox * sizeof(double));
  FILE* fi = fopen(argv[1], "rb");
  if (!fi || fread(host, sizeof(double), nvox, fi) != nvox) {
    fprintf(stderr, "cannot read %zu doubles from %s\n", nvox, argv[1]);
    return 1;
  }
  fclose(fi);
  if (!tvdk_device_ok()) {
    fprintf(stderr, "%s\n", tvdk_last_error());
    return 4;
  }

  tvdk_problem_t p;
  memset(&p, 0, sizeof p);
  p.geom.dim = 2;
  p.geom.dtype = TVDK_F64;
  p.geom.n[0] = n0; p.geom.n[1] = n1; p.geom.n[2] = 1;
  p.geom.pitch = tvdk_pitch(TVDK_F64, n1);
  p.geom.nb = 1;
  p.geom.g0_off = 0; p.geom.g0_n = n0; p.geom.z_begin = 0; p.geom.z_end = n0;
  const int64_t fe = tvdk_field_elems(&p.geom);
  if (fe <= 0) { fprintf(stderr, "%s\n", tvdk_last_error()); return 3; }

  double *d_host_order, *d_f, *d_u, *d_t;
  CK(cudaMalloc((void**)&d_host_order, nvox * sizeof(double)));
  CK(cudaMalloc((void**)&d_f, fe * sizeof(double)));
  CK(cudaMalloc((void**)&d_u, fe * sizeof(double)));
  CK(cudaMalloc((void**)&d_t, 2 * fe * sizeof(double)));
  CK(cudaMemset(d_f, 0, fe * sizeof(double)));
  CK(cudaMemset(d_t, 0, 2 * fe * sizeof(double)));            /* d = b = 0            (:100-101) */
  CK(cudaMemcpy(d_host_order, host, nvox * sizeof(double), cudaMemcpyHostToDevice));
  TK(tvdk_pack(&p.geom, TVDK_F64, d_host_order, d_f, NULL));   /* host order -> padded planar field */
  CK(cudaMemcpy(d_u, d_f, fe * sizeof(double), cudaMemcpyDeviceToDevice)); /* u0 = img (:99) */

  p.isotropic = atoi(argv[5]);
  p.solver_kind = atoi(argv[6]) ? TVDK_SOLVER_CG : TVDK_SOLVER_JACOBI;
  p.solver_maxiter = atoi(argv[7]);
  p.solver_tol = p.solver_kind == TVDK_SOLVER_CG ? 0.0 : -1.0;  /* CG: fixed count; Jacobi: tol = None */
  p.max_num_iter = atoi(argv[8]);
  p.eps = -1.0;                                                 /* eps = None */
  p.mu.value = 0.1; p.omega.value = 1.0; p.ell.value = 1.0;     /* scalars: device_ptr = NULL */
  p.f = d_f; p.u = d_u; p.t = d_t; p.t_comp_stride = fe;
  p.fused = 1;
  p.check_every = 8;
  p.workspace_bytes = tvdk_workspace_bytes(&p);
  if (p.workspace_bytes == 0) { fprintf(stderr, "%s\n", tvdk_last_error()); return 3; }
  CK(cudaMalloc(&p.workspace, p.workspace_bytes));
  CK(cudaMalloc((void**)&p.stats, sizeof(tvdk_stat_t)));

  int launched = 0;
  TK(tvdk_solve(&p, NULL, &launched));                          /* the whole outer loop (:129-175) */
  TK(tvdk_unpack(&p.geom, TVDK_F64, d_u, d_host_order, NULL));
  CK(cudaMemcpy(host, d_host_order, nvox * sizeof(double), cudaMemcpyDeviceToHost));
  tvdk_stat_t st;
  CK(cudaMemcpy(&st, p.stats, sizeof st, cudaMemcpyDeviceToHost));
  printf("iterations %d (launched %d), last relative increment %.6e, kernels launched %lld\n", st.iters, launched,
         st.inc, (long long)tvdk_launch_count(0));

  FILE* fo = fopen(argv[2], "wb");
  if (!fo || fwrite(host, sizeof(double), nvox, fo) != nvox) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
  fclose(fo);
  cudaFree(d_host_order); cudaFree(d_f); cudaFree(d_u); cudaFree(d_t); cudaFree(p.workspace); cudaFree(p.stats);
  free(host);
  return 0;
}
