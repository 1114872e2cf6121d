// tvdk.cu -- sm_100a backend of libtvdk: the __global__ runner that wraps every kernel functor
// of tvdk_core.h, the deterministic two-stage reduction, and the CUDA plumbing the host driver
// (tvdk_host.inl) needs.  Build: see darsia_b200/csrc/Makefile (nvcc -gencode
// arch=compute_100a,code=sm_100a -lineinfo).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <type_traits>

#include "tvdk_core.h"

namespace tvdk {

constexpr int kMaxThreads = 512;
constexpr int kMaxWarps = kMaxThreads / 32;

// Block-wide sum with a fixed combination order: xor-shuffle tree inside each warp, then the
// warp results are added in warp order by thread 0.  Bitwise reproducible for a given launch
// shape (no atomics on floating-point data anywhere).
template <int NRED>
__device__ __forceinline__ void block_sum(double (&val)[NRED], double (*wsum)[NRED], int tid, int nthreads) {
  const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
#pragma unroll
  for (int r = 0; r < NRED; ++r) {
    double v = val[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) wsum[warp][r] = v;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < NRED; ++r) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += wsum[w][r];
      val[r] = s;
    }
  }
  __syncthreads();
}

template <class K, typename = void>
struct has_unroll2 : std::false_type {};
template <class K>
struct has_unroll2<K, std::void_t<decltype(K::UNROLL2)>> : std::bool_constant<K::UNROLL2> {};

// kernels may cap their block size (K::MAXT) to get more than 128 registers per thread
template <class K, typename = void>
struct max_threads : std::integral_constant<int, kMaxThreads> {};
template <class K>
struct max_threads<K, std::void_t<decltype(K::MAXT)>> : std::integral_constant<int, K::MAXT> {};

// (no minBlocksPerMultiprocessor argument: even "1" makes ptxas spend ~20 more registers per thread on
//  every kernel, which costs occupancy -- measured -15 % on the shrink kernel, profiles/r1_ab_launch_bounds.txt)
template <class K>
__global__ void __launch_bounds__(max_threads<K>::value) tvdk_kernel(const typename K::Args a, double* partials,
                                                            unsigned* counters) {
  extern __shared__ __align__(16) unsigned char smem[];
  LaunchCtx lc;
  lc.tx = threadIdx.x; lc.ty = threadIdx.y; lc.ntx = blockDim.x; lc.nty = blockDim.y;
  lc.bx = blockIdx.x; lc.by = blockIdx.y; lc.bz = blockIdx.z; lc.nbx = gridDim.x; lc.nby = gridDim.y;
  K k;
  if (!k.init(a, lc, smem)) return;  // block-uniform (whole problems are skipped)
  if constexpr (has_unroll2<K>::value) {
    // march loop unrolled by two with a compile-time phase (see KFusedJ1)
    while (k.more()) {
      k.template pre_p<0>(a, smem);
      __syncthreads();
      k.template post_p<0>(a, smem);
      if (!k.more()) break;
      k.template pre_p<1>(a, smem);
      __syncthreads();
      k.template post_p<1>(a, smem);
    }
  } else {
    while (k.more()) {
      k.pre(a, smem);
      __syncthreads();
      k.post(a, smem);
    }
  }
  if constexpr (K::NRED > 0) {
    constexpr int NR = K::NRED;
    __shared__ double wsum[kMaxWarps][NR];
    __shared__ bool is_last;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const unsigned nblk = gridDim.x * gridDim.y;
    const unsigned blk = blockIdx.y * gridDim.x + blockIdx.x;
    const int b = blockIdx.z;
    double v[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = k.red[r];
    block_sum<NR>(v, wsum, tid, nthreads);
    // partials: [problem][sum][block] -- the second stage reads them coalesced
    double* pb = partials + (size_t)b * nblk * NR;
    if (tid == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) pb[(size_t)r * nblk + blk] = v[r];
      __threadfence();
      const unsigned prev = atomicAdd(&counters[b], 1u);
      is_last = (prev == nblk - 1);
    }
    __syncthreads();
    if (is_last) {  // last block of this problem: ordered second stage + scalar updates
      __threadfence();
      double s[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) s[r] = 0.0;
      // fixed order per thread (i = tid, tid + nthreads, ...), four independent loads in flight per sum
      unsigned i = tid;
      for (; i + 3u * nthreads < nblk; i += 4u * nthreads) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const double* q = pb + (size_t)r * nblk + i;
          const double x0 = __ldcg(q), x1 = __ldcg(q + nthreads), x2 = __ldcg(q + 2 * nthreads), x3 = __ldcg(q + 3 * nthreads);
          s[r] += x0;
          s[r] += x1;
          s[r] += x2;
          s[r] += x3;
        }
      }
      for (; i < nblk; i += nthreads) {
#pragma unroll
        for (int r = 0; r < NR; ++r) s[r] += __ldcg(pb + (size_t)r * nblk + i);
      }
      block_sum<NR>(s, wsum, tid, nthreads);
      if (tid == 0) {
        K::finalize(a, b, s);
        counters[b] = 0;
      }
    }
  }
}

namespace backend {

static int cuda_fail(cudaError_t e, const char* what);

template <class K>
int launch(const typename K::Args& a, int gx, int gy, int gz, int ntx, int nty, size_t smem, void* stream,
           double* partials, unsigned* counters) {
  static thread_local size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(tvdk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    smem_set = smem;
  }
  dim3 grid(gx, gy, gz), block(ntx, nty, 1);
  tvdk_kernel<K><<<grid, block, smem, (cudaStream_t)stream>>>(a, partials, counters);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  return 0;
}

static int memset_async(void* p, int value, size_t bytes, void* stream) {
  cudaError_t e = cudaMemsetAsync(p, value, bytes, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaMemsetAsync");
}

// two pinned ints + events per host thread for the convergence polling
struct PollState {
  int* host = nullptr;
  cudaEvent_t ev[2];
  bool ready = false;
};
static thread_local PollState g_poll;

static int poll_init() {
  if (g_poll.ready) return 0;
  cudaError_t e = cudaHostAlloc((void**)&g_poll.host, 2 * sizeof(int), cudaHostAllocDefault);
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc");
  for (int i = 0; i < 2; ++i) {
    e = cudaEventCreateWithFlags(&g_poll.ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
  }
  g_poll.ready = true;
  return 0;
}

static int poll_begin(int slot, const int* dev_flag, void* stream) {
  int rc = poll_init();
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(&g_poll.host[slot], dev_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync");
  e = cudaEventRecord(g_poll.ev[slot], (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaEventRecord");
}

static int poll_end(int slot, int* value) {
  cudaError_t e = cudaEventSynchronize(g_poll.ev[slot]);
  if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
  *value = g_poll.host[slot];
  return 0;
}

static int device_ok(char* msg, size_t n) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    snprintf(msg, n, "no CUDA device: %s", cudaGetErrorString(e));
    return 0;
  }
  // (cudaDeviceGetAttribute, not cudaGetDeviceProperties: the latter takes milliseconds per call)
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) {
    snprintf(msg, n, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return 0;
  }
  if (major != 10) {
    snprintf(msg, n, "libtvdk is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return 0;
  }
  return 1;
}

}  // namespace backend
}  // namespace tvdk

#include "tvdk_host.inl"

namespace tvdk {
namespace backend {
static int cuda_fail(cudaError_t e, const char* what) {
  return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}
}  // namespace backend
}  // namespace tvdk
