// tvdk.cu -- sm_100a backend of libtvdk: the __global__ runner that wraps every kernel functor
// of tvdk_core.h, the deterministic two-stage reduction, and the CUDA plumbing the host driver
// (tvdk_host.inl) needs.  Build: see darsia_b200/csrc/Makefile (nvcc -gencode
// arch=compute_100a,code=sm_100a -lineinfo).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <type_traits>
#include <typeinfo>
#include <utility>
#include <vector>

#include "tvdk_core.h"

#ifndef TVDK_PDL_DEFAULT
#define TVDK_PDL_DEFAULT 1
#endif

// NVTX ranges around the solves (header-only NVTX3: a no-op unless a profiler is attached)
#include <nvtx3/nvToolsExt.h>
namespace tvdk {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace tvdk
#define TVDK_RANGE(name) ::tvdk::NvtxRange tvdk_nvtx_range_(name)

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

// Which axis-0 chunk this thread block works on.  Blocks are dispatched in ascending blockIdx order; with g.rev the
// chunks are taken from the far end first, so that a launch starts on the planes its predecessor wrote LAST -- for
// images of a few L2 sizes (BASELINE config 1: 7 fields x 45 MB vs 126 MB of L2) those are still resident.
__device__ __forceinline__ int chunk_index(const Geom& g) {
  return g.rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
}

template <class K, typename = void>
struct has_unroll2 : std::false_type {};
template <class K>
struct has_unroll2<K, std::void_t<decltype(K::UNROLL2)>> : std::bool_constant<K::UNROLL2> {};

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may become RESIDENT while its predecessor in the stream drains; `griddepcontrol.wait` holds it until the predecessor
// has completed and its writes are visible (a no-op for a plain launch), `launch_dependents` then lets the successor
// do the same behind this grid.  Nothing is read or written before the wait, so the data flow is that of plain
// stream order; what is saved is the launch latency and the block-scheduling ramp between two launches.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

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
  pdl_prologue();
  LaunchCtx lc;
  lc.tx = threadIdx.x; lc.ty = threadIdx.y; lc.ntx = blockDim.x; lc.nty = blockDim.y;
  lc.bx = blockIdx.x; lc.by = chunk_index(a.g); lc.bz = blockIdx.z; lc.nbx = gridDim.x; lc.nby = gridDim.y;
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
    const unsigned blk = lc.by * gridDim.x + blockIdx.x;  // (logical chunk: the order of the sums does not depend on g.rev)
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

// Runner of the TMA-staged kernels: the argument block is __grid_constant__ (the tensor maps are read from
// kernel-parameter space by the TMA unit), the mbarriers are initialised before anybody may wait on them.
template <class K>
__global__ void __launch_bounds__(max_threads<K>::value) tvdk_kernel_tma(const __grid_constant__ typename K::Args a,
                                                                         double* partials, unsigned* counters) {
  extern __shared__ __align__(128) unsigned char smem_tma[];
  pdl_prologue();
  LaunchCtx lc;
  lc.tx = threadIdx.x; lc.ty = threadIdx.y; lc.ntx = blockDim.x; lc.nty = blockDim.y;
  lc.bx = blockIdx.x; lc.by = chunk_index(a.g); lc.bz = blockIdx.z; lc.nbx = gridDim.x; lc.nby = gridDim.y;
  if (K::skipped(a, lc)) return;  // block-uniform (whole problems are skipped)
  K::bar_init(a, lc, smem_tma);
  __syncthreads();
  K k;
  k.init(a, lc, smem_tma);
  if constexpr (K::SYNC_AFTER_INIT) __syncthreads();  // (what init published in shared memory is read by the first step)
  while (k.more()) {
    k.pre(a, smem_tma);
    __syncthreads();
    k.post(a, smem_tma);
  }
  if constexpr (K::NRED > 0) {
    constexpr int NR = K::NRED;
    __shared__ double wsum[kMaxWarps][NR];
    __shared__ bool is_last;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const unsigned nblk = gridDim.x * gridDim.y;
    const unsigned blk = lc.by * gridDim.x + blockIdx.x;  // (logical chunk: the order of the sums does not depend on g.rev)
    const int b = blockIdx.z;
    double v[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = k.red[r];
    block_sum<NR>(v, wsum, tid, nthreads);
    double* pb = partials + (size_t)b * nblk * NR;
    if (tid == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) pb[(size_t)r * nblk + blk] = v[r];
      __threadfence();
      const unsigned prev = atomicAdd(&counters[b], 1u);
      is_last = (prev == nblk - 1);
    }
    __syncthreads();
    if (is_last) {  // last block of this problem: ordered second stage + scalar updates (as in tvdk_kernel)
      __threadfence();
      double s[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) s[r] = 0.0;
      for (unsigned i = tid; i < nblk; i += nthreads) {
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

template <class K, typename = void>
struct is_tma_kernel : std::false_type {};
template <class K>
struct is_tma_kernel<K, std::void_t<decltype(K::TMA)>> : std::bool_constant<K::TMA> {};

constexpr int kMaxDevices = 64;
static int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

// ---- optional timing of one kernel family inside a real solve (tvdk_profile_arm / tvdk_profile_read) ----------------
// While armed, every launch whose kernel type name contains the pattern is bracketed by two CUDA events on its own
// stream.  Disarmed (the normal state) this costs one load of a flag per launch.
struct ProfState {
  std::mutex mu;
  std::string pattern;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
};
static std::atomic<bool> g_prof_on{false};
static ProfState& prof_state() {
  static ProfState s;
  return s;
}
template <class K>
struct ProfScope {
  cudaEvent_t stop = nullptr;
  cudaStream_t stream;
  explicit ProfScope(void* st) : stream((cudaStream_t)st) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfState& ps = prof_state();
    std::lock_guard<std::mutex> lk(ps.mu);
    if (ps.pattern.empty() || !strstr(typeid(K).name(), ps.pattern.c_str()) || ps.events.size() >= (1u << 20)) return;
    cudaEvent_t start = nullptr;
    if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) {
      stop = nullptr;
      return;
    }
    cudaEventRecord(start, stream);
    ps.events.emplace_back(start, stop);
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, stream);
  }
};

// launch mode of the march / TMA kernels: plain stream order, or programmatic dependent launch (tvdk_set_pdl, TVDK_PDL)
static bool pdl_default() {
  const char* e = getenv("TVDK_PDL");
  return e ? atoi(e) != 0 : TVDK_PDL_DEFAULT != 0;
}
static std::atomic<bool> g_pdl{pdl_default()};
static void pdl_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, dim3 grid, dim3 block, size_t smem, void* stream) {
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
}

template <class K>
int launch(const typename K::Args& a, int gx, int gy, int gz, int ntx, int nty, size_t smem, void* stream,
           double* partials, unsigned* counters) {
  ProfScope<K> prof(stream);
  // the opt-in above 48 KB is a per-DEVICE function attribute: cache it per (thread, kernel, device)
  static thread_local size_t smem_set[kMaxDevices] = {};
  if (smem > 48 * 1024) {
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices || smem > smem_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(tvdk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
      if (dev >= 0 && dev < kMaxDevices) smem_set[dev] = smem;
    }
  }
  dim3 grid(gx, gy, gz), block(ntx, nty, 1);
  if (g_pdl.load(std::memory_order_relaxed)) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    pdl_config(cfg, at, grid, block, smem, stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tvdk_kernel<K>, a, partials, counters);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch (programmatic dependent launch)");
    return 0;
  }
  tvdk_kernel<K><<<grid, block, smem, (cudaStream_t)stream>>>(a, partials, counters);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  return 0;
}

template <class K>
int launch_tma(const typename K::Args& a, int gx, int gy, int gz, int ntx, int nty, size_t smem, void* stream,
               double* partials, unsigned* counters) {
  ProfScope<K> prof(stream);
  static thread_local size_t smem_set[kMaxDevices] = {};
  if (smem > 48 * 1024) {
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices || smem > smem_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(tvdk_kernel_tma<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
      if (dev >= 0 && dev < kMaxDevices) smem_set[dev] = smem;
    }
  }
  dim3 grid(gx, gy, gz), block(ntx, nty, 1);
  if (g_pdl.load(std::memory_order_relaxed)) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    pdl_config(cfg, at, grid, block, smem, stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, tvdk_kernel_tma<K>, a, partials, counters);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch (TMA, programmatic dependent launch)");
    return 0;
  }
  tvdk_kernel_tma<K><<<grid, block, smem, (cudaStream_t)stream>>>(a, partials, counters);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch (TMA)");
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda): a rank-3 tiled map of one field,
// no swizzle, no interleave, zero fill outside the tensor
static int tma_encode(TmaMap* m, const void* base, int esize, long long nx, long long ny, long long nz, long long row_bytes,
                      long long plane_bytes, int bx, int by);

// resident thread blocks per SM of a kernel at a launch shape (cached: the planner asks on every launch)
template <class K>
int occupancy(int threads, size_t smem) {
  struct Entry {
    int threads;
    size_t smem;
    int blocks;
  };
  static thread_local Entry cache[8] = {};
  static thread_local int used = 0;
  for (int i = 0; i < used; ++i)
    if (cache[i].threads == threads && cache[i].smem == smem) return cache[i].blocks;
  int blocks = 0;
  cudaError_t e;
  if (smem > 48 * 1024) {  // the opt-in has to be in place for the query to count the launch as valid
    if constexpr (is_tma_kernel<K>::value)
      cudaFuncSetAttribute(tvdk_kernel_tma<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else
      cudaFuncSetAttribute(tvdk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if constexpr (is_tma_kernel<K>::value)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, tvdk_kernel_tma<K>, threads, smem);
  else
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, tvdk_kernel<K>, threads, smem);
  if (e != cudaSuccess) {
    cudaGetLastError();
    blocks = 4;
  }
  if (used < 8) cache[used++] = Entry{threads, smem, blocks};
  return blocks;
}

static int memset_async(void* p, int value, size_t bytes, void* stream) {
  cudaError_t e = cudaMemsetAsync(p, value, bytes, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaMemsetAsync");
}

// two pinned ints + events per (host thread, device) for the convergence polling: events belong to the
// device that was current when they were created
struct PollState {
  int* host = nullptr;
  cudaEvent_t ev[2];
  bool ready = false;
};
static thread_local PollState g_poll[kMaxDevices];

static PollState* poll_state() {
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) {
    cuda_fail(cudaErrorInvalidDevice, "convergence polling: no current device");
    return nullptr;
  }
  PollState& ps = g_poll[dev];
  if (ps.ready) return &ps;
  cudaError_t e = cudaHostAlloc((void**)&ps.host, 2 * sizeof(int), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cuda_fail(e, "cudaHostAlloc");
    return nullptr;
  }
  for (int i = 0; i < 2; ++i) {
    e = cudaEventCreateWithFlags(&ps.ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      cuda_fail(e, "cudaEventCreate");
      return nullptr;
    }
  }
  ps.ready = true;
  return &ps;
}

static int poll_begin(int slot, const int* dev_flag, void* stream) {
  PollState* ps = poll_state();
  if (!ps) return TVDK_ERR_INVALID;
  cudaError_t e = cudaMemcpyAsync(&ps->host[slot], dev_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync");
  e = cudaEventRecord(ps->ev[slot], (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaEventRecord");
}

static int poll_end(int slot, int* value) {
  PollState* ps = poll_state();
  if (!ps) return TVDK_ERR_INVALID;
  cudaError_t e = cudaEventSynchronize(ps->ev[slot]);
  if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
  *value = ps->host[slot];
  return 0;
}

// ---- slab decomposition: halo + partial-sum push over peer memory, flag wait + ordered sum --------------
// Push: grid-stride 128-bit copy of the boundary planes into the neighbours' ghost planes (peer stores over
// NVLink; plain stores when the neighbour lives on the same device), then the last block publishes the partial
// sums and raises this rank's flag in every rank's mailbox.
__global__ void __launch_bounds__(256) comm_push_kernel(const CommPush a) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (int k = 0; k < a.nseg; ++k) {
    const uint4* src = reinterpret_cast<const uint4*>(a.seg[k].src);
    uint4* dst = reinterpret_cast<uint4*>(a.seg[k].dst);
    const long long n = a.seg[k].bytes >> 4;
    long long i = tid;
    for (; i + 3 * nth < n; i += 4 * nth) {  // four independent 128-bit loads in flight per thread
      const uint4 v0 = src[i], v1 = src[i + nth], v2 = src[i + 2 * nth], v3 = src[i + 3 * nth];
      dst[i] = v0;
      dst[i + nth] = v1;
      dst[i + 2 * nth] = v2;
      dst[i + 3 * nth] = v3;
    }
    for (; i < n; i += nth) dst[i] = src[i];
  }
  __threadfence_system();  // this thread's peer stores are visible system-wide before the block is counted
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if ((int)threadIdx.x < a.world) {
    MailSlot* m = a.mail[threadIdx.x];
    for (int r = 0; r < a.nred; ++r) m->v[r] = __ldcg(a.sums + r);
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&m->flag) = a.flag;
  }
  if (threadIdx.x == 0) *a.counter = 0;
}

// Finalize: lane s waits for rank s's flag (bounded), then one thread adds the partial sums in rank order
// and applies the scalar update (finalize_fin) -- identical bits on every rank.
__global__ void __launch_bounds__(32) comm_finalize_kernel(const CommFin a) {
  const int lane = threadIdx.x;
  bool ok = true;
  if (lane < a.world) {
    const volatile unsigned long long* fl = &a.mail[lane].flag;
    if (*fl != a.flag) {
      unsigned long long t0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      unsigned ns = 32;
      while (*fl != a.flag) {
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if ((long long)(t1 - t0) > a.timeout_ns) {
          ok = false;
          break;
        }
      }
    }
  }
  const bool all_ok = __all_sync(0xffffffffu, ok);
  __threadfence_system();  // order the reads of planes / sums after the flag reads
  if (lane == 0) {
    if (!all_ok) {
      if (*a.error == 0) *a.error = (long long)a.flag;
      return;
    }
    comm_finalize_body(a);
  }
}

static int comm_push(const CommPush& a, void* stream) {
  long long bytes = 0;
  for (int k = 0; k < a.nseg; ++k) bytes += a.seg[k].bytes;
  // enough CTAs to drive one NVLink direction (a few MB per step), few enough to leave the SMs to compute
  int grid = (int)((bytes / 16 + 256 * 8 - 1) / (256 * 8));
  grid = grid < 1 ? 1 : (grid > 64 ? 64 : grid);
  comm_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "comm_push launch");
}

static int comm_finalize(const CommFin& a, void* stream) {
  comm_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "comm_finalize launch");
}

static int set_device(int dev) {
  if (dev < 0) return 0;
  cudaError_t e = cudaSetDevice(dev);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaSetDevice");
}

// event edges between the streams of the ranks one host thread drives: after this call, work enqueued on any
// of the streams runs after everything enqueued so far on all of them
static int comm_sync_local(int n, void* const* streams, const int* devices) {
  static thread_local cudaEvent_t ev[TVDK_MAX_SLABS];
  static thread_local int ev_dev[TVDK_MAX_SLABS];
  static thread_local bool ev_ok[TVDK_MAX_SLABS] = {};
  bool same = true;
  for (int i = 1; i < n; ++i) same = same && streams[i] == streams[0] && devices[i] == devices[0];
  if (same) return 0;  // one stream: already ordered
  for (int i = 0; i < n; ++i) {
    int rc = set_device(devices[i]);
    if (rc) return rc;
    const int dev = current_device();
    if (!ev_ok[i] || ev_dev[i] != dev) {
      if (ev_ok[i]) cudaEventDestroy(ev[i]);
      cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
      ev_ok[i] = true;
      ev_dev[i] = dev;
    }
    cudaError_t e = cudaEventRecord(ev[i], (cudaStream_t)streams[i]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
  }
  for (int i = 0; i < n; ++i) {
    int rc = set_device(devices[i]);
    if (rc) return rc;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      cudaError_t e = cudaStreamWaitEvent((cudaStream_t)streams[i], ev[j], 0);
      if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent");
    }
  }
  return 0;
}

static int read_int(const int* dev_ptr, void* stream, int* value) {
  cudaError_t e = cudaMemcpyAsync(value, dev_ptr, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "read_int");
}
static int read_i64(const long long* dev_ptr, void* stream, long long* value) {
  cudaError_t e = cudaMemcpyAsync(value, dev_ptr, sizeof(long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(e, "read_i64");
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
#include "tvdk_slab.inl"

namespace tvdk {
namespace backend {
static int cuda_fail(cudaError_t e, const char* what) {
  return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}
}  // namespace backend
}  // namespace tvdk

// ---- peer mapping helpers (include/tvdk.h) ---------------------------------------------------------------
#include <cuda.h>

namespace tvdk {
namespace backend {
static int tma_encode(TmaMap* m, const void* base, int esize, long long nx, long long ny, long long nz, long long row_bytes,
                      long long plane_bytes, int bx, int by) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || !fn) return cuda_fail(e, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
    encode = (encode_fn)fn;
  }
  static_assert(sizeof(TmaMap) == sizeof(CUtensorMap), "TmaMap must be a CUtensorMap");
  if (bx > 256 || by > 256 || (bx * esize) % 16 != 0) return fail(TVDK_ERR_INVALID, "internal: bad TMA box %d x %d", bx, by);
  if (((uintptr_t)base & 15) || (row_bytes & 15) || (plane_bytes & 15))
    return fail(TVDK_ERR_INVALID, "TMA needs 16-byte aligned fields and row pitches");
  const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
  const cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)plane_bytes};
  const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(m), esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                            3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TVDK_ERR_INVALID, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}
}  // namespace backend
}  // namespace tvdk

int tvdk_set_pdl(int on) {
  const bool was = tvdk::backend::g_pdl.exchange(on != 0);
  return was ? 1 : 0;
}

int tvdk_profile_arm(const char* kernel_substring) {
  using namespace tvdk;
  backend::ProfState& ps = backend::prof_state();
  std::lock_guard<std::mutex> lk(ps.mu);
  for (auto& ev : ps.events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  ps.events.clear();
  ps.pattern = kernel_substring ? kernel_substring : "";
  backend::g_prof_on.store(!ps.pattern.empty());
  return 0;
}

int tvdk_profile_read(int64_t* launches, double* total_ms) {
  using namespace tvdk;
  if (!launches || !total_ms) return fail(TVDK_ERR_INVALID, "NULL pointer");
  backend::ProfState& ps = backend::prof_state();
  std::lock_guard<std::mutex> lk(ps.mu);
  backend::g_prof_on.store(false);
  double sum = 0.0;
  cudaError_t bad = cudaSuccess;
  for (auto& ev : ps.events) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(ev.second);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev.first, ev.second);
    if (e != cudaSuccess) bad = e;
    sum += ms;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  *launches = (int64_t)ps.events.size();
  *total_ms = sum;
  ps.events.clear();
  ps.pattern.clear();
  return bad == cudaSuccess ? 0 : backend::cuda_fail(bad, "tvdk_profile_read");
}

int tvdk_ipc_export(const void* device_ptr, unsigned char handle_out[64], int64_t* offset_out) {
  using namespace tvdk;
  if (!device_ptr || !handle_out || !offset_out) return fail(TVDK_ERR_INVALID, "NULL pointer");
  // the IPC handle names the whole cudaMalloc allocation: find its base (driver entry point, no -lcuda needed)
  typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static range_fn get_range = nullptr;
  if (!get_range) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || !fn) return backend::cuda_fail(e, "cudaGetDriverEntryPoint(cuMemGetAddressRange)");
    get_range = (range_fn)fn;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (CUdeviceptr)device_ptr) != CUDA_SUCCESS)
    return fail(TVDK_ERR_INVALID, "cuMemGetAddressRange failed: not a device allocation");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (void*)base);
  if (e != cudaSuccess) return backend::cuda_fail(e, "cudaIpcGetMemHandle");
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  *offset_out = (int64_t)((CUdeviceptr)device_ptr - base);
  return 0;
}

int tvdk_ipc_open(const unsigned char handle[64], int64_t offset, void** ptr_out) {
  using namespace tvdk;
  if (!handle || !ptr_out) return fail(TVDK_ERR_INVALID, "NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return backend::cuda_fail(e, "cudaIpcOpenMemHandle");
  *ptr_out = reinterpret_cast<unsigned char*>(base) + offset;
  return 0;
}

int tvdk_ipc_close(void* ptr, int64_t offset) {
  using namespace tvdk;
  if (!ptr) return fail(TVDK_ERR_INVALID, "NULL pointer");
  // (the exporting device keeps the allocation alive for as long as any process still maps it)
  cudaError_t e = cudaIpcCloseMemHandle(reinterpret_cast<unsigned char*>(ptr) - offset);
  return e == cudaSuccess ? 0 : backend::cuda_fail(e, "cudaIpcCloseMemHandle");
}

int tvdk_enable_peer(int peer_device) {
  using namespace tvdk;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return backend::cuda_fail(e, "cudaGetDevice");
  if (dev == peer_device) return 0;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) return backend::cuda_fail(e, "cudaDeviceCanAccessPeer");
  if (!can) return fail(TVDK_ERR_UNSUPPORTED, "device %d cannot access device %d", dev, peer_device);
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  return e == cudaSuccess ? 0 : backend::cuda_fail(e, "cudaDeviceEnablePeerAccess");
}
