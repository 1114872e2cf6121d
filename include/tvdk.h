/*
 * tvdk.h -- C ABI of libtvdk, the B200-native (sm_100a) split-Bregman TV denoising kernels.
 *
 * This is the drop-in boundary of the hot path.  The DarSIA reference is pure Python and has
 * no FFI layer; the "interface" each entry point replaces is the Python callable cited beside
 * it (paths relative to /root/reference/src/darsia/).  The binding a maintainer would add on
 * the reference side (ctypes) is shown in INTEGRATION.md; darsia_b200/_cabi.py is that binding.
 *
 * Conventions
 *   - every pointer named "device" is a raw CUDA device pointer owned by the caller (PyTorch);
 *     the library allocates nothing persistent on the device and frees nothing it did not
 *     allocate; the workspace is caller-provided (size from tvdk_workspace_bytes);
 *   - every function returns 0 on success, a negative tvdk error code, or a positive
 *     cudaError_t; tvdk_last_error() returns a thread-local message;
 *   - "stream" is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - no global mutable state besides the thread-local error string: safe to drive 8 GPUs
 *     from 8 threads or 8 processes.
 *
 * Device data layout ("field"): batch-major planar, last axis contiguous and padded to
 * `pitch` elements (a multiple of 16 bytes so every row start is 128-bit aligned):
 *     2-D: field[b][i0][pitch]            3-D: field[b][i0][i1][pitch]
 * The split-Bregman state is kept COMPRESSED: instead of the reference's d and b arrays
 * (split_bregman_tvd.py:100-101) the library stores t_a = grad_a(u) + b_a (one field per
 * spatial axis), from which d_a = shrink(t)_a and b_a = t_a - d_a are recomputed with the
 * reference's own arithmetic (split_bregman_tvd.py:117-126, 142-146) wherever they are
 * needed.  Component a of t lives at t + a * comp_stride elements.
 */
#ifndef TVDK_H
#define TVDK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVDK_VERSION 100 /* 0.1.0 */

/* error codes (negative) */
#define TVDK_OK 0
#define TVDK_ERR_INVALID (-1)   /* bad argument */
#define TVDK_ERR_UNSUPPORTED (-2) /* valid in DarSIA, not implemented here */
#define TVDK_ERR_WORKSPACE (-3) /* workspace too small */
#define TVDK_ERR_NODEVICE (-4)  /* no CUDA device / wrong architecture */

/* dtype codes */
#define TVDK_F64 0
#define TVDK_F32 1

/* solver kinds: darsia.Jacobi (utils/linear_solvers/jacobi.py:14), darsia.CG (cg.py:10) */
#define TVDK_SOLVER_JACOBI 0
#define TVDK_SOLVER_CG 1

/* Geometry of one batch of independent dim-dimensional problems. */
typedef struct tvdk_geom {
  int32_t dim;      /* 2 or 3 (split_bregman_tvd.py:19 `dim`) */
  int32_t dtype;    /* TVDK_F64 / TVDK_F32 */
  int32_t n[3];     /* local extents, slowest first; n[2] unused for dim == 2 */
  int32_t pitch;    /* elements per row of the last axis, multiple of 16 bytes */
  int32_t nb;       /* number of independent problems (channels / frames), >= 1 */
  /* Slab decomposition along axis 0 (multi-GPU, one volume): the local array holds planes
   * [0, n[0]) of which [z_begin, z_end) are owned/computed, the rest are ghost planes filled by
   * the caller's halo exchange; local plane 0 is global plane g0_off of g0_n (g0_off may be
   * negative for the unused lower ghosts of the first slab).
   * Single GPU: g0_off = 0, g0_n = n[0], z_begin = 0, z_end = n[0]. */
  int32_t g0_off;
  int32_t g0_n;
  int32_t z_begin;
  int32_t z_end;
} tvdk_geom_t;

/* A coefficient (mu / omega / ell of split_bregman_tvd.py:16-18): scalar, or a per-voxel
 * field with the geometry's layout; batch_stride = 0 shares one map between all problems. */
typedef struct tvdk_coef {
  const void* device_ptr; /* NULL -> scalar `value` */
  double value;
  int64_t batch_stride;   /* elements; 0 or the field size */
} tvdk_coef_t;

/* Per-problem result record written by tvdk_solve (device memory, nb entries). */
typedef struct tvdk_stat {
  double nrm_f;     /* ||img||_2                       (split_bregman_tvd.py:58)  */
  double inc;       /* last relative increment          (split_bregman_tvd.py:161) */
  double energy;    /* last energy functional if asked  (split_bregman_tvd.py:71-77) */
  double bnrm, atol, rho, rho_prev, pq, alpha, beta; /* CG scalars (scipy cg recurrence) */
  double jac_x0nrm2, jac_err;                        /* Jacobi tol branch (jacobi.py:118-131) */
  int32_t iters;    /* outer iterations performed */
  int32_t done;     /* 1 once inc < eps fired       (split_bregman_tvd.py:173-175) */
  int32_t cg_stop, cg_first, cg_iters;
  int32_t jac_stop, jac_sel;
  int32_t cur_buf;  /* internal: which buffer holds the iterate */
  double e_init, e_prev; /* Chambolle: energy of iteration 0 and of the previous iteration */
  /* single-pass CG (K7): the predicted rho of the next iteration, and the largest relative deviation
   * |rho_predicted - rho_summed| / rho_summed seen during the solve (diagnostic: ~1e-15 on well-conditioned systems) */
  double rho_pred, cg_pred_dev;
  /* 1 once a relative increment came out NaN / Inf (the reference would silently carry it on); the problem is
   * marked done so that the remaining iterations leave the last finite-input state alone */
  int32_t nonfinite;
  int32_t pad_;
} tvdk_stat_t;

/* One call of darsia.split_bregman_tvd (restoration/split_bregman_tvd.py:14-27). */
typedef struct tvdk_problem {
  tvdk_geom_t geom;
  int32_t isotropic;       /* split_bregman_tvd.py:23 */
  int32_t solver_kind;     /* TVDK_SOLVER_* */
  int32_t solver_maxiter;  /* Solver.maxiter (solver.py:20) */
  int32_t max_num_iter;    /* split_bregman_tvd.py:20 */
  double solver_tol;       /* Solver.tol; < 0 means None (Jacobi only) */
  double eps;              /* < 0 means None (split_bregman_tvd.py:21) */
  tvdk_coef_t mu, omega, ell;
  const void* f;           /* device: image field (read only) */
  void* u;                 /* device: in = initial iterate (copy of f, or x0[0]); out = result */
  void* t;                 /* device: dim fields, compressed state; in = zeros unless warm start */
  int64_t t_comp_stride;   /* elements between components of t */
  /* Warm start x0 = (u0, d0, b0) (split_bregman_tvd.py:93-97): d0/b0 as planar fields; the
   * first outer iteration reads them explicitly, afterwards the state is compressed. */
  const void* x0_d;        /* device or NULL */
  const void* x0_b;        /* device or NULL */
  int64_t x0_comp_stride;
  void* workspace;         /* device, tvdk_workspace_bytes() bytes, 256-byte aligned */
  size_t workspace_bytes;
  tvdk_stat_t* stats;      /* device, nb entries */
  double* history;         /* device or NULL: max_num_iter * nb relative increments */
  double* energy_history;  /* device or NULL: energies (verbose), max_num_iter * nb */
  int32_t verbose_every;   /* 0 = never compute the energy functional */
  int32_t check_every;     /* host polls the device "all done" flag every N outer iterations */
  int32_t fused;           /* 1 = allow the fused per-iteration kernels (default), 0 = building blocks only,
                            * 2 = take every temporally blocked / single-pass variant wherever it is implemented,
                            *     not only where it measured faster (3-D single-pass CG and blocked sweeps,
                            *     K1x2 pairs for isotropic / per-voxel / fp32 problems and small images),
                            * 3 = as 1 but one outer iteration per launch (no K1x2 pairs),
                            * 4 = as 1, with the TMA-staged fused Jacobi(1) kernel for 2-D anisotropic problems with
                            *     scalar coefficients whatever the environment variable TVDK_TMA2D says */
  int32_t tile_x, tile_y, zchunk; /* 0 = library default; thread-block shape / planes per CTA */
  /* `adaptive` schedule (split_bregman_tvd.py:26, :149-158): HOST array of max_num_iter flags,
   * adaptive_mask[it] != 0 <=> adaptive(it) is true: after the shrinkage of iteration `it`
   * ell <- 1 / max(sum_a |grad_a u+|, 1e-12) per voxel.  NULL = no adaptivity. */
  const uint8_t* adaptive_mask;
} tvdk_problem_t;

int tvdk_version(void);
const char* tvdk_last_error(void);
/* 1 if a CUDA device of compute capability 10.x is current, else 0 (and sets last_error). */
int tvdk_device_ok(void);
/* Number of kernels this thread has launched through the library since the last reset. */
int64_t tvdk_launch_count(int reset);
/* Thread-local default launch shape used wherever a call passes 0 (tuning / tests):
 * thread-block x, thread-block y (3-D), planes per CTA. */
int tvdk_set_tune(int tile_x, int tile_y, int zchunk);

/* Recommended pitch (elements) for a last-axis extent. */
int tvdk_pitch(int dtype, int n_last);
/* Elements of one field (n0*[n1*]pitch) and of one batch of fields. */
int64_t tvdk_field_elems(const tvdk_geom_t* g);

/* Workspace needed by tvdk_solve for this problem (device pointers are ignored). */
size_t tvdk_workspace_bytes(const tvdk_problem_t* p);

/* The whole outer loop of darsia.split_bregman_tvd (split_bregman_tvd.py:129-175), device
 * resident; returns after enqueueing (and polling for convergence when eps >= 0).
 * `iters_launched` (host, may be NULL) receives the number of outer iterations enqueued. */
int tvdk_solve(const tvdk_problem_t* p, void* stream, int* iters_launched);

/* Layout conversion between the host-order array (spatial axes..., batch axes flattened last,
 * C-contiguous: element (i, b) at i*nb + b) and the planar padded device layout, with dtype
 * cast.  src/dst are device pointers.  (replaces img.copy()/img_as_float, :99, and the final
 * convert_dtype, split_bregman_tvd.py:177) */
int tvdk_pack(const tvdk_geom_t* g, int src_dtype, const void* src_interleaved, void* dst_field, void* stream);
int tvdk_unpack(const tvdk_geom_t* g, int dst_dtype, const void* src_field, void* dst_interleaved, void* stream);

/* ---- test-granularity operators (each one is one kernel launch) -------------------------- */

/* rhs = omega*f + sum_a H_a(ell*(b_a - d_a)) from the compressed state t
 * (split_bregman_tvd.py:84-89 `_rhs_function`). */
int tvdk_rhs(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* omega,
             const tvdk_coef_t* ell, const void* f, const void* t, int64_t t_comp_stride, void* rhs,
             void* workspace, size_t workspace_bytes, void* stream);

/* One Jacobi sweep x_out = rhs/diag + N(x_in)/(diag/ell) (jacobi.py:101-114). */
int tvdk_jacobi_sweep(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell,
                      const void* x_in, const void* rhs, void* x_out, void* workspace,
                      size_t workspace_bytes, void* stream);

/* y = omega*x - laplace(x, diffusion_coeff=ell) (cg.py:37-44, derivatives.py:88-104). */
int tvdk_cg_matvec(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell,
                   const void* x, void* y, void* workspace, size_t workspace_bytes, void* stream);

/* Shrinkage + Bregman update on the compressed state: t_out = grad(u_new) + b(t_in)
 * (split_bregman_tvd.py:136-146); sum_out[b] (device, doubles) = ||u_new - u_old||^2; if u_out is
 * not NULL the owned planes of u_new are copied into it (split_bregman_tvd.py:170; it may alias
 * u_old, t_out may alias t_in). */
int tvdk_shrink(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* ell,
                const void* u_new, const void* u_old, void* u_out, const void* t_in, void* t_out,
                int64_t t_comp_stride, double* sum_out, void* workspace, size_t workspace_bytes,
                void* stream);

/* The same two operators for the FIRST outer iteration after a warm start x0 = (u, d, b)
 * (split_bregman_tvd.py:93-97): an arbitrary (d, b) need not be the expansion of any compressed state, so
 * the right-hand side takes d and b and the shrinkage takes b as planar fields (dim components,
 * `*_comp_stride` elements apart); t_out is the compressed state the following iterations continue from.
 * (tvdk_solve does the same internally for x0_d / x0_b; these are for host-composed loops such as MG.) */
int tvdk_rhs_explicit(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* f,
                      const void* d, const void* b, int64_t db_comp_stride, void* rhs, void* workspace,
                      size_t workspace_bytes, void* stream);
int tvdk_shrink_explicit(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* ell,
                         const void* u_new, const void* u_old, void* u_out, const void* b_in,
                         int64_t b_comp_stride, void* t_out, int64_t t_comp_stride, double* sum_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Expand the compressed state into the reference's d and b (planar fields, dim each). */
int tvdk_expand_state(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu,
                      const tvdk_coef_t* ell, const void* t, int64_t t_comp_stride, void* d, void* b,
                      int64_t db_comp_stride, void* stream);

/* Deterministic reductions over the valid voxels of each problem: out[b] (device doubles). */
int tvdk_dot(const tvdk_geom_t* g, const void* x, const void* y, double* out, void* workspace,
             size_t workspace_bytes, void* stream);
int tvdk_norm_diff2(const tvdk_geom_t* g, const void* x, const void* y, double* out, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Energy functional 1/2||omega(x-f)||^2 + sum_a sum |mu G_a x| (split_bregman_tvd.py:71-77). */
int tvdk_energy(const tvdk_geom_t* g, const tvdk_coef_t* mu, const tvdk_coef_t* omega, const void* x,
                const void* f, double* out, void* workspace, size_t workspace_bytes, void* stream);

/* Workspace needed by the single-operator entry points above. */
size_t tvdk_op_workspace_bytes(const tvdk_geom_t* g);

/* One fused outer iteration with Jacobi(maxiter=1): (u_in, t_in) -> (u_out, t_out), and
 * sum_out[b] = ||u_out - u_in||^2 (the K1 kernel; used by the slab-decomposed driver). */
int tvdk_fused_jacobi1(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu,
                       const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* f,
                       const void* u_in, void* u_out, const void* t_in, void* t_out,
                       int64_t t_comp_stride, double* sum_out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* TWO fused outer iterations with Jacobi(maxiter=1) in one pass (2-D only; what tvdk_solve runs for
 * pairs of iterations of large anisotropic fp64 problems with scalar coefficients when eps is None):
 * (u_in, t_in) -> two iterations -> (u_out, t_out);
 * sum_out[2b], sum_out[2b+1] = the two increments' squared norms. */
int tvdk_fused_jacobi1_x2(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu,
                          const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* f,
                          const void* u_in, void* u_out, const void* t_in, void* t_out,
                          int64_t t_comp_stride, double* sum_out, void* workspace,
                          size_t workspace_bytes, void* stream);

/* CG building blocks for the slab-decomposed driver (scalars supplied by the caller after the
 * cross-rank reduction; each returns per-problem partial sums in sum_out, device doubles).
 *   cg_init:   r = rhs(t,f) - A u ; x = u ; sum_out[2b] = ||rhs||^2, sum_out[2b+1] = r.r
 *   cg_pq:     p_out = first ? r : beta[b]*p_in + r ; sum_out[b] = p_out . (A p_out)
 *   cg_update: x += alpha[b]*p ; r -= alpha[b]*(A p) ; sum_out[b] = r.r                      */
int tvdk_cg_init(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* omega,
                 const tvdk_coef_t* ell, const void* f, const void* t, int64_t t_comp_stride,
                 const void* u, void* x, void* r, double* sum_out, void* workspace,
                 size_t workspace_bytes, void* stream);
int tvdk_cg_pq(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell,
               const void* p_in, const void* r, void* p_out, const double* beta, int first,
               double* sum_out, void* workspace, size_t workspace_bytes, void* stream);
int tvdk_cg_update(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell,
                   const void* p, void* x, void* r, const double* alpha, double* sum_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Single-pass CG iterations (what tvdk_solve runs for CG with tol == 0; same scipy recurrence as
 * above, cg.py:50-56, reordered so that x, r and p stream through the GPU once per iteration):
 *   tvdk_cg_first:  sum_out[2b..] = { r.(A r), (A r).(A r) }          (p_0 = r_0 is never stored)
 *   tvdk_cg_fused:  x += alpha[b] p_in;  r_out = r_in - alpha[b] A p_in;  p_out = r_out + beta[b] p_in;
 *                   sum_out[4b..] = { r_out.r_out, p_out.(A p_out), r_out.(A p_out), (A p_out).(A p_out) }
 * p_in may equal r_in (first fused iteration); p_out / r_out must not alias p_in / r_in.
 * The caller derives alpha = rho / (p.q) and beta = rho_next / rho with the PREDICTED
 * rho_next = rho - 2 alpha (r.q) + alpha^2 (q.q)  (= ||r - alpha q||^2). */
int tvdk_cg_first(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* r,
                  double* sum_out, void* workspace, size_t workspace_bytes, void* stream);
int tvdk_cg_fused(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell,
                  const void* p_in, const void* r_in, void* p_out, void* r_out, void* x,
                  const double* alpha, const double* beta, double* sum_out, void* workspace,
                  size_t workspace_bytes, void* stream);

/* n Jacobi sweeps (jacobi.py:109-114, no tol branch) ping-ponging between xa (in: the start iterate) and
 * xb; in 2-D with scalar coefficients they are temporally blocked (2 or 4 sweeps per pass).
 * *result_in_b tells which buffer holds the result; the other one is scratch afterwards. */
int tvdk_jacobi_sweeps(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, void* xa,
                       const void* rhs, void* xb, int nsweeps, int* result_in_b, void* workspace,
                       size_t workspace_bytes, void* stream);

/* One call of skimage.restoration.denoise_tv_chambolle(image, weight, eps, max_num_iter) with every axis
 * spatial -- what darsia.tvd(img) runs by default (restoration/tvd.py:77-83).  Third-party algorithm,
 * PARITY UNPINNED (oracle/chambolle_oracle.py).  geom.nb must be 1 or a batch of independent images. */
typedef struct tvdk_chambolle {
  tvdk_geom_t geom;
  double weight, eps;
  int32_t max_num_iter;    /* >= 1 */
  int32_t check_every;     /* host polls the device "all done" flag every N iterations (0: 8) */
  const void* f;           /* device: image field */
  void* out;               /* device: result field */
  void* p0;                /* device: dim fields, zero on entry (the dual variable) */
  void* p1;                /* device: dim fields, scratch */
  int64_t p_comp_stride;
  void* workspace;         /* device, tvdk_op_workspace_bytes(&geom) bytes */
  size_t workspace_bytes;
  tvdk_stat_t* stats;      /* device, nb entries: iters, energy, done */
  double* history;         /* device or NULL: max_num_iter * nb energies */
} tvdk_chambolle_t;
int tvdk_chambolle(const tvdk_chambolle_t* c, void* stream, int* iters_launched);

/* Multigrid building blocks (darsia.MG, utils/linear_solvers/mg.py:94-242; the V-cycle itself is
 * host logic over these and tvdk_jacobi_sweep):
 *   tvdk_residual:        r = rhs - (omega x - ell Lap x)        (mg.py:212; pass ell / h^2 for spacing h; a
 *                         per-voxel ell multiplies the unit-coefficient Laplacian from outside, mg.py:104-106)
 *   tvdk_mg_restrict:     coarse = pairwise averages of fine     (mg.py:107-132; coarse extents = fine / 2)
 *   tvdk_mg_prolong_add:  fine += repeat-twice(coarse), edge-padded for odd extents (mg.py:226-239)
 * `fine` and `coarse` describe two grids of the same dim / dtype / nb. */
int tvdk_residual(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* x,
                  const void* rhs, void* r, void* workspace, size_t workspace_bytes, void* stream);
int tvdk_mg_restrict(const tvdk_geom_t* fine, const tvdk_geom_t* coarse, const void* x_fine, void* x_coarse,
                     void* stream);
int tvdk_mg_prolong_add(const tvdk_geom_t* fine, const tvdk_geom_t* coarse, const void* e_coarse, void* x_fine,
                        void* stream);

/* ---- slab-decomposed solve: ONE volume cut along axis 0 over several GPUs -----------------------
 *
 * The reference has no distributed code; this is split_bregman_tvd.py:14-177 (same arithmetic per voxel
 * as tvdk_solve) for a volume whose planes of axis 0 are distributed over `world` ranks, SURVEY.md 8(e).
 * No NCCL on the data path: every rank owns an ARENA (one device allocation holding a small header, the
 * fields f, u, t and the workspace) that all other ranks have mapped (CUDA peer access inside one process,
 * CUDA IPC between processes).  After each kernel a rank PUSHES the boundary planes its neighbours need
 * straight into their ghost planes and its partial sums into every rank's mailbox (peer stores over
 * NVLink), then raises a flag; a one-warp kernel waits for the flags of all ranks, adds the partial sums
 * in rank order (bitwise identical on every rank) and applies the scalar update (alpha, beta, increment,
 * stopping flags) on the device -- halo exchange + all-reduce in two tiny launches, no host round trip.
 *
 * Layout rules shared by all ranks (so that a field has the same arena offset everywhere):
 *   - local arrays hold  TVDK_SLAB_GL ghost planes | owned planes | TVDK_SLAB_GH ghost planes (+ unused
 *     planes up to the common capacity): geom.n[0] = GL + max_r nown[r] + GH on every rank,
 *     geom.z_begin = GL, geom.z_end = GL + nown[rank], geom.g0_off = first owned global plane - GL;
 *   - problem.f / u / t / workspace / per-voxel coefficient maps all point into the rank's arena, at the
 *     same offsets on every rank; the first tvdk_slab_header_bytes() bytes of the arena belong to the
 *     library and must be zero before the first solve.
 * Unsupported here (TVDK_ERR_UNSUPPORTED): warm starts (x0_d / x0_b), the adaptive schedule, verbose energy.
 */
#define TVDK_MAX_SLABS 16
#define TVDK_SLAB_GL 1
#define TVDK_SLAB_GH 2

typedef struct tvdk_slab_comm {
  int32_t world;                 /* number of slabs (ranks) */
  int32_t rank;                  /* this slab */
  int32_t nown[TVDK_MAX_SLABS];  /* owned planes of every rank */
  void* arena[TVDK_MAX_SLABS];   /* base of every rank's arena as mapped for THIS rank's device (own included) */
  size_t arena_bytes;
  uint64_t* seq;                 /* host counter, advanced by the library; identical on all ranks (start at 0) */
  int32_t timeout_ms;            /* flag wait limit inside the kernels (0: 20000); on expiry the header records it */
} tvdk_slab_comm_t;

size_t tvdk_slab_header_bytes(void);
/* Workspace needed by tvdk_solve_slab for this problem (scratch + the solver's extra fields). */
size_t tvdk_slab_workspace_bytes(const tvdk_problem_t* p);

/* Drive `nlocal` ranks of one decomposition from this host thread in lock step (nlocal = 1 with one process
 * per GPU; nlocal = world with every slab in one process: tests, or one process driving all GPUs).  Local
 * ranks are ordered by event dependencies between their streams, remote ranks by the flags.  problems[i],
 * comms[i], streams[i], devices[i] (CUDA device ordinal made current for rank i's launches; -1: leave the
 * current device) describe local rank i.  All ranks must pass the same solver parameters.
 * With eps >= 0 the stop test is read back (blocking) every check_every outer iterations so that all ranks
 * take the same decision.  Returns after enqueueing. */
int tvdk_solve_slab(int nlocal, const tvdk_problem_t* const* problems, const tvdk_slab_comm_t* const* comms,
                    void* const* streams, const int* devices, int* iters_launched);

/* First sequence number whose flag wait timed out on this rank (0 = none); reads the arena header
 * (synchronises the stream). */
int tvdk_slab_error(const void* arena, void* stream, int64_t* seq_out);

/* One halo exchange of whole planes without a solve (static fields: the image, coefficient maps):
 * `fields[k]` (arena pointers) each send n_up last owned planes up and n_down first owned planes down. */
int tvdk_slab_exchange(int nlocal, const tvdk_geom_t* const* geoms, const tvdk_slab_comm_t* const* comms,
                       int nfields, const void* const* const* fields, int n_up, int n_down,
                       void* const* streams, const int* devices);

/* Launch mode of the compute kernels, process-wide: 0 = plain stream order, 1 = programmatic dependent launch (the
 * next kernel's thread blocks become resident while the previous kernel drains and are held by
 * `griddepcontrol.wait` until it has completed: same data flow, shorter gaps between launches).  Returns the
 * previous mode.  Initial value: environment variable TVDK_PDL, else the build's default. */
int tvdk_set_pdl(int on);

/* Diagnostics for measurements (bench.py's in-step kernel time): from now on every launch of a kernel whose C++ type
 * name contains `kernel_substring` (e.g. "KCgFused") is bracketed by a pair of CUDA events on the stream it is launched
 * on; NULL or "" disarms.  tvdk_profile_read waits for the recorded events, returns their number and summed duration,
 * frees them and disarms.  Disarmed -- the normal state -- a launch pays one flag load. */
int tvdk_profile_arm(const char* kernel_substring);
int tvdk_profile_read(int64_t* launches, double* total_ms);

/* Peer mapping helpers for one-process-per-GPU runs (CUDA IPC; thin wrappers so that callers need no CUDA
 * bindings): export a 64-byte handle + offset for a device pointer / open one exported by another process
 * of the same node (returns the pointer mapped in this process) / unmap it again (`ptr` and `offset` as
 * tvdk_ipc_open returned / took them; the exporter's memory stays allocated on ITS device until every
 * importer has closed the mapping) / enable peer access from the current device to `peer_device`. */
int tvdk_ipc_export(const void* device_ptr, unsigned char handle_out[64], int64_t* offset_out);
int tvdk_ipc_open(const unsigned char handle[64], int64_t offset, void** ptr_out);
int tvdk_ipc_close(void* ptr, int64_t offset);
int tvdk_enable_peer(int peer_device);

#ifdef __cplusplus
}
#endif
#endif /* TVDK_H */
