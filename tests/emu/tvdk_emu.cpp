// tvdk_emu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A CPU "thread stepper" for the kernel functors of darsia_b200/csrc/tvdk_core.h: the very same
// kernel source is compiled with g++ and every thread of a block is advanced phase by phase
// (init | pre | barrier | post ...), blocks one after another.  It exposes the same C ABI as
// libtvdk.so with host pointers in place of device pointers, so the CPU test-suite can check
// the kernels' index/halo/boundary logic, the host driver and the multi-rank (gloo) paths
// against the oracle without a GPU.  The product (darsia_b200) never loads this library: its
// loader only accepts darsia_b200/csrc/libtvdk.so and a CUDA device.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "../../darsia_b200/csrc/tvdk_core.h"
#define TVDK_RANGE(name) (void)0

namespace tvdk {
namespace backend {

template <class K, typename = void>
struct has_tma : std::false_type {};
template <class K>
struct has_tma<K, std::void_t<decltype(K::TMA)>> : std::bool_constant<K::TMA> {};

template <class K>
int launch(const typename K::Args& a, int gx, int gy, int gz, int ntx, int nty, size_t smem_bytes, void*,
           double*, unsigned*) {
  const int nthreads = ntx * nty;
  std::vector<K> th(nthreads);
  std::vector<unsigned char> smem_store(smem_bytes + 256, 0);
  // 128-byte aligned base, like the device's dynamic shared memory
  unsigned char* smem = smem_store.data();
  smem += (128 - (reinterpret_cast<uintptr_t>(smem) & 127)) & 127;
  constexpr int NR = K::NRED > 0 ? K::NRED : 1;
  for (int bz = 0; bz < gz; ++bz) {
    double tot[NR];
    for (int r = 0; r < NR; ++r) tot[r] = 0.0;
    bool any = false;
    for (int by = 0; by < gy; ++by)
      for (int bx = 0; bx < gx; ++bx) {
        int alive = 0;
        if constexpr (has_tma<K>::value) {
          LaunchCtx l0;
          l0.tx = l0.ty = 0; l0.ntx = ntx; l0.nty = nty; l0.bx = bx; l0.by = by; l0.bz = bz; l0.nbx = gx; l0.nby = gy;
          if (K::skipped(a, l0)) continue;
          K::bar_init(a, l0, smem);
        }
        for (int t = 0; t < nthreads; ++t) {
          LaunchCtx lc;
          lc.tx = t % ntx; lc.ty = t / ntx; lc.ntx = ntx; lc.nty = nty;
          lc.bx = bx; lc.by = a.g.rev ? gy - 1 - by : by; lc.bz = bz; lc.nbx = gx; lc.nby = gy;
          alive += th[t].init(a, lc, smem) ? 1 : 0;
        }
        if (alive == 0) continue;
        if (alive != nthreads) abort();  // early exit must be block-uniform
        any = true;
        while (th[0].more()) {
          for (int t = 0; t < nthreads; ++t) th[t].pre(a, smem);
          for (int t = 0; t < nthreads; ++t) th[t].post(a, smem);
        }
        if (K::NRED > 0) {
          double blk[NR];
          for (int r = 0; r < NR; ++r) blk[r] = 0.0;
          for (int t = 0; t < nthreads; ++t)
            for (int r = 0; r < K::NRED; ++r) blk[r] += th[t].red[r];
          for (int r = 0; r < NR; ++r) tot[r] += blk[r];
        }
      }
    if (K::NRED > 0 && any) K::finalize(a, bz, tot);
  }
  return 0;
}

// TMA-staged kernels: same stepping; the "TMA unit" is a synchronous copy inside the issuing thread
template <class K>
int launch_tma(const typename K::Args& a, int gx, int gy, int gz, int ntx, int nty, size_t smem_bytes, void* stream,
               double* partials, unsigned* counters) {
  return launch<K>(a, gx, gy, gz, ntx, nty, smem_bytes, stream, partials, counters);
}
static int tma_encode(TmaMap* m, const void* base, int esize, long long nx, long long ny, long long nz, long long row_bytes,
                      long long plane_bytes, int bx, int by) {
  m->base = reinterpret_cast<const unsigned char*>(base);
  m->dim[0] = nx; m->dim[1] = ny; m->dim[2] = nz;
  m->stride_b[0] = row_bytes; m->stride_b[1] = plane_bytes;
  m->box[0] = bx; m->box[1] = by; m->box[2] = 1;
  m->esize = esize;
  return 0;
}

template <class K>
int occupancy(int, size_t) {
  return 8;
}

static int memset_async(void* p, int value, size_t bytes, void*) {
  memset(p, value, bytes);
  return 0;
}
static int g_poll_val[2];
static int poll_begin(int slot, const int* flag, void*) {
  g_poll_val[slot] = *flag;
  return 0;
}
static int poll_end(int slot, int* value) {
  *value = g_poll_val[slot];
  return 0;
}
static int device_ok(char*, size_t) { return 1; }

// slab decomposition: "peer memory" is host memory of the same process, kernels run synchronously in the
// calling thread; ranks driven in lock step see every flag already raised, ranks driven from different
// threads (one per rank, like one process per GPU) really wait on each other's flags
static int comm_push(const CommPush& a, void*) {
  for (int k = 0; k < a.nseg; ++k) memcpy(a.seg[k].dst, a.seg[k].src, (size_t)a.seg[k].bytes);
  for (int s = 0; s < a.world; ++s) {
    for (int r = 0; r < a.nred; ++r) a.mail[s]->v[r] = a.sums[r];
    __atomic_store_n(&a.mail[s]->flag, a.flag, __ATOMIC_RELEASE);
  }
  return 0;
}
static int comm_finalize(const CommFin& a, void*) {
  for (int s = 0; s < a.world; ++s) {
    long long spins = 0;
    while (__atomic_load_n(&a.mail[s].flag, __ATOMIC_ACQUIRE) != a.flag) {
      if (++spins > 2000000000LL) {  // a lock-step group can never get here; threads give up after ~seconds
        if (*a.error == 0) *a.error = (long long)a.flag;
        return 0;
      }
    }
  }
  comm_finalize_body(a);
  return 0;
}
static int comm_sync_local(int, void* const*, const int*) { return 0; }
static int set_device(int) { return 0; }
static int read_int(const int* p, void*, int* v) {
  *v = *p;
  return 0;
}
static int read_i64(const long long* p, void*, long long* v) {
  *v = *p;
  return 0;
}

}  // namespace backend
}  // namespace tvdk

#include "../../darsia_b200/csrc/tvdk_host.inl"
#include "../../darsia_b200/csrc/tvdk_slab.inl"

// (peer mapping helpers: one address space here)
int tvdk_ipc_export(const void* p, unsigned char h[64], int64_t* off) {
  memset(h, 0, 64);
  memcpy(h, &p, sizeof p);
  *off = 0;
  return 0;
}
int tvdk_ipc_open(const unsigned char h[64], int64_t off, void** out) {
  void* p = nullptr;
  memcpy(&p, h, sizeof p);
  *out = reinterpret_cast<unsigned char*>(p) + off;
  return 0;
}
int tvdk_ipc_close(void*, int64_t) { return 0; }
int tvdk_set_pdl(int) { return 0; }  // (no launches to overlap)
// (no device clock here: nothing is ever recorded)
int tvdk_profile_arm(const char*) { return 0; }
int tvdk_profile_read(int64_t* launches, double* total_ms) {
  if (!launches || !total_ms) return tvdk::fail(TVDK_ERR_INVALID, "NULL pointer");
  *launches = 0;
  *total_ms = 0.0;
  return 0;
}
int tvdk_enable_peer(int) { return 0; }

extern "C" int tvdk_is_emulator(void) { return 1; }
