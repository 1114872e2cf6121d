// tvdk_host.inl -- host side of libtvdk: argument checking, launch geometry, the outer
// split-Bregman loop, and the extern "C" entry points declared in include/tvdk.h.
//
// Included by exactly one translation unit per backend, after the backend has defined
//   namespace tvdk::backend {
//     template <class K> int launch(const typename K::Args&, int gx, int gy, int gz, int ntx,
//                                   int nty, size_t smem, void* stream, double* partials,
//                                   unsigned* counters);
//     int memset_async(void* p, int value, size_t bytes, void* stream);
//     int poll_begin(int slot, const int* dev_flag, void* stream);   // async read of a device int
//     int poll_end(int slot, int* value);                            // wait for it
//     int device_ok(char* msg, size_t n);
//   }
// The sm_100a backend is tvdk.cu; tests/emu/tvdk_emu.cpp is a CPU thread-stepper used only by
// the CPU test-suite.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <type_traits>

namespace tvdk {

static thread_local std::string g_last_error;
static thread_local long long g_launches = 0;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define TVDK_TRY(expr)            \
  do {                            \
    int rc_ = (expr);             \
    if (rc_ != 0) return rc_;     \
  } while (0)

// ---------------------------------------------------------------------------------------
// geometry / coefficient conversion
// ---------------------------------------------------------------------------------------
static int check_geom(const tvdk_geom_t* g) {
  if (!g) return fail(TVDK_ERR_INVALID, "geom is NULL");
  if (g->dim != 2 && g->dim != 3) return fail(TVDK_ERR_INVALID, "dim must be 2 or 3, got %d", g->dim);
  if (g->dtype != TVDK_F64 && g->dtype != TVDK_F32) return fail(TVDK_ERR_INVALID, "bad dtype %d", g->dtype);
  for (int a = 0; a < g->dim; ++a)
    if (g->n[a] < 1) return fail(TVDK_ERR_INVALID, "extent n[%d] = %d must be >= 1", a, g->n[a]);
  const int vec = g->dtype == TVDK_F64 ? 2 : 4;
  const int nlast = g->n[g->dim - 1];
  if (g->pitch < nlast || g->pitch % vec != 0)
    return fail(TVDK_ERR_INVALID, "pitch %d must be >= %d and a multiple of %d", g->pitch, nlast, vec);
  if (g->nb < 1 || g->nb > 65535) return fail(TVDK_ERR_INVALID, "nb = %d out of range [1, 65535]", g->nb);
  if (g->z_begin < 0 || g->z_end > g->n[0] || g->z_begin >= g->z_end)
    return fail(TVDK_ERR_INVALID, "bad owned plane range [%d, %d) of %d", g->z_begin, g->z_end, g->n[0]);
  // g0_off may be slightly negative: the (unused) lower ghost planes of the first slab
  if (g->g0_off + g->z_begin < 0 || g->g0_off + g->z_end > g->g0_n)
    return fail(TVDK_ERR_INVALID, "owned planes [%d, %d) + offset %d exceed the global extent %d", g->z_begin,
                g->z_end, g->g0_off, g->g0_n);
  return 0;
}

// L2 prefetch distance of the march kernels in planes: per kernel (K::PF2D / K::PF3D, measured);
// the environment variables TVDK_PF (planes, 0 = off) and TVDK_PFSUB (1 = one thread per
// 128-byte line issues the hint) override it for tuning runs.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int prefetch_override() {
  static const int v = std::min(env_int("TVDK_PF", -1), 64);
  return v;
}
static int prefetch_sub() {
  static const int v = env_int("TVDK_PFSUB", 0) != 0;
  return v;
}
// Jacobi sweeps per pass (K4xS), measured in profiles/r1_sweepn.txt: 2-D 4000x7000x3 fp64: 0.332 ms per
// sweep alone, 0.171 two per pass, 0.153 four per pass; 1788x3180: 0.033 / 0.0185 / 0.0192; 3-D 512^3:
// 0.56 alone vs 0.62 two per pass (two halo rows on each side of y cost more than the bytes saved).
static int sweeps_per_pass(int dim, long long voxels, int remaining, bool hetero, bool force3d = false) {
  if (hetero || remaining < 2) return 1;
  if (dim != 2) return force3d ? 2 : 1;
  return (remaining >= 4 && voxels >= (1LL << 23)) ? 4 : 2;
}

template <class K, typename = void>
struct KernelMaxT {
  static constexpr int value = 512;
};
template <class K>
struct KernelMaxT<K, std::void_t<decltype(K::MAXT)>> {
  static constexpr int value = K::MAXT;
};
template <class K, typename = void>
struct KernelZc {
  static int get(int) { return 32; }
};
template <class K>
struct KernelZc<K, std::void_t<decltype(K::ZC2D)>> {
  static int get(int dim) { return dim == 2 ? K::ZC2D : 32; }
};
template <class K, typename = void>
struct KernelZc3 {
  static constexpr int value = 32;
};
template <class K>
struct KernelZc3<K, std::void_t<decltype(K::ZC3D)>> {
  static constexpr int value = K::ZC3D;
};
template <class K, typename = void>
struct KernelNtx3 {
  static constexpr int value = 32;
};
template <class K>
struct KernelNtx3<K, std::void_t<decltype(K::NTX3D)>> {
  static constexpr int value = K::NTX3D;
};
template <class K, typename = void>
struct KernelNty3 {
  static int get(bool big_tile) { return big_tile ? 16 : 8; }
};
template <class K>
struct KernelNty3<K, std::void_t<decltype(K::NTY3D)>> {
  static int get(bool) { return K::NTY3D; }
};
template <class K, typename = void>
struct KernelNtx2 : std::integral_constant<int, 128> {};
template <class K>
struct KernelNtx2<K, std::void_t<decltype(K::NTX2D)>> : std::integral_constant<int, K::NTX2D> {};
template <class K, typename = void>
struct KernelNtx2Max : std::integral_constant<int, 1024> {};
template <class K>
struct KernelNtx2Max<K, std::void_t<decltype(K::NTX2D_MAX)>> : std::integral_constant<int, K::NTX2D_MAX> {};
template <class K, typename = void>
struct KernelIsTma : std::false_type {};
template <class K>
struct KernelIsTma<K, std::void_t<decltype(K::TMA)>> : std::bool_constant<K::TMA> {};
template <class K, typename = void>
struct KernelPf {
  static int get(int) { return 0; }
};
template <class K>
struct KernelPf<K, std::void_t<decltype(K::PF2D)>> {
  static int get(int dim) { return dim == 2 ? K::PF2D : K::PF3D; }
};

static Geom make_geom(const tvdk_geom_t& g) {
  Geom r;
  r.n0 = g.n[0];
  r.ny = g.dim == 3 ? g.n[1] : 1;
  r.nx = g.n[g.dim - 1];
  r.pitch = g.pitch;
  r.s0 = (long long)r.ny * r.pitch;
  r.sb = (long long)r.n0 * r.s0;
  r.nb = g.nb;
  r.g0_off = g.g0_off;
  r.g0_n = g.g0_n;
  r.z_begin = g.z_begin;
  r.z_end = g.z_end;
  r.zchunk = g.z_end - g.z_begin;
  r.xtiles = r.ytiles = 1;
  r.pf = 0;
  r.pfsub = prefetch_sub();
  r.rev = 0;
  return r;
}

template <typename T>
static Coef<T> make_coef(const tvdk_coef_t& c) {
  Coef<T> r;
  r.p = reinterpret_cast<const T*>(c.device_ptr);
  r.val = (T)c.value;
  r.sb = c.batch_stride;
  return r;
}

// Scalars as the reference computes them with Python floats (jacobi.py:57-72,101-106;
// split_bregman_tvd.py:140,145), cast to T afterwards.
template <typename T>
static Coefs<T> make_coefs(const tvdk_coef_t* mu, const tvdk_coef_t* om, const tvdk_coef_t* el, int dim) {
  static const tvdk_coef_t one = {nullptr, 1.0, 0};
  if (!mu) mu = &one;
  if (!om) om = &one;
  if (!el) el = &one;
  Coefs<T> k;
  k.mu = make_coef<T>(*mu);
  k.omega = make_coef<T>(*om);
  k.ell = make_coef<T>(*el);
  Scalars<T>& s = k.sc;
  s.mu = (T)mu->value;
  s.omega = (T)om->value;
  s.ell = (T)el->value;
  s.kappa = (T)(mu->value / el->value);
  const double diag = om->value + el->value * 2 * dim / 1.0;
  s.inv_diag = (T)(1.0 / diag);
  s.c_nb = (T)(1.0 / (diag / el->value));
  s.dim_t = (T)dim;
  s.hetero = (mu->device_ptr || om->device_ptr || el->device_ptr) ? 1 : 0;
  return k;
}

// ---------------------------------------------------------------------------------------
// launch geometry
// ---------------------------------------------------------------------------------------
struct Tune {
  int tile_x, tile_y, zchunk;  // 0 = default
};

struct LaunchCfg {
  int ntx, nty, gx, gy, gz;
  size_t smem;
  Geom g;
};

static thread_local Tune g_tune = {0, 0, 0};
// Inside a solve consecutive march launches alternate the direction in which their thread blocks walk the axis-0 chunks
// (Geom::rev): every kernel reads what its predecessor wrote, and starts where that one ended -- on lines still in L2.
// TVDK_ALTERNATE=0 switches it off (A/B runs).  The results do not depend on it (same per-voxel arithmetic, reductions
// summed by logical chunk index).
static thread_local int g_alternate = 0;  // 0: off ; 1 / 2: next launch ascending / descending
static thread_local int g_force_variants = 0;  // tvdk_problem_t.fused == 2: take every implemented kernel variant
static thread_local int g_force_tma2d = 0;     // tvdk_problem_t.fused == 4: the 2-D TMA-staged fused kernel whatever TVDK_TMA2D says
static bool alternate_enabled() {
  static const int v = env_int("TVDK_ALTERNATE", 1);
  return v != 0;
}
static int next_direction() {
  if (!g_alternate) return 0;
  const int rev = g_alternate == 2 ? 1 : 0;
  g_alternate = 3 - g_alternate;
  return rev;
}


static Tune merge_tune(int tx, int ty, int zc) {
  Tune t;
  t.tile_x = tx > 0 ? tx : g_tune.tile_x;
  t.tile_y = ty > 0 ? ty : g_tune.tile_y;
  t.zchunk = zc > 0 ? zc : g_tune.zchunk;
  return t;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }
static bool wave_aware_enabled() {
  static const int v = env_int("TVDK_WAVES", 1);
  return v != 0;
}

template <class K, typename T, int D>
static LaunchCfg plan_launch(Geom g, const Tune& tune, bool big_tile) {
  typedef typename K::C C;
  constexpr int V = Vec<T>::N;
  LaunchCfg L;
  const int halo_x = C::LHX + C::RHX;
  const int need_x = (g.nx + V - 1) / V + halo_x;  // vectors to cover one row incl. halo
  if (D == 2) {
    int ntx = tune.tile_x > 0 ? round_up(tune.tile_x, 32) : KernelNtx2<K>::value;
    ntx = std::min(ntx, round_up(need_x, 32));
    ntx = std::min(ntx, KernelMaxT<K>::value);
    ntx = std::min(ntx, KernelNtx2Max<K>::value);  // (TMA-staged kernels: a box holds at most 256 elements per axis)
    ntx = std::max(ntx, 32);
    L.ntx = ntx;
    L.nty = 1;
  } else {
    int ntx = tune.tile_x > 0 ? tune.tile_x : KernelNtx3<K>::value;
    ntx = std::min(ntx, round_up(need_x, 8));
    ntx = std::max(ntx, 8);
    int nty = tune.tile_y > 0 ? tune.tile_y : KernelNty3<K>::get(big_tile);
    const int halo_y = K::LH + K::RH;
    nty = std::min(nty, round_up(g.ny + halo_y, 4));
    nty = std::max(nty, halo_y + 1);
    while ((ntx * nty) % 32 != 0) ++nty;
    while (ntx * nty > 512 && nty > halo_y + 1) --nty;
    while ((ntx * nty) % 32 != 0 && nty > halo_y + 1) --nty;
    L.ntx = ntx;
    L.nty = nty;
  }
  const int iw = C::iw(L.ntx), ih = C::ih(L.nty);
  g.xtiles = (g.nx + iw - 1) / iw;
  g.ytiles = D == 3 ? (g.ny + ih - 1) / ih : 1;
  const int nz = g.z_end - g.z_begin;
  int zc = tune.zchunk;
  if (zc <= 0) {
    // Many more CTAs than resident slots (148 SMs x 4..8 CTAs): several waves, so the tail wave
    // is a small fraction, while a chunk stays long enough to amortise its 1-2 plane prologue.
    // Measured on B200 (profiles/r1_tune_2d_f64.txt): 32 planes per CTA is the sweet spot in 2-D.
    // (per problem, NOT times nb: the launch shape -- hence the summation order of every
    // reduction -- must not depend on how many problems share a launch, so that a frame gives
    // bitwise the same result alone, in a batch, or on another rank)
    const long long tiles = (long long)g.xtiles * g.ytiles;
    const long long target = 148LL * 32;
    const long long want = std::max(1LL, (target + tiles - 1) / tiles);
    zc = (int)((nz + want - 1) / want);
    zc = std::max(zc, 8);
    zc = std::min(zc, 32);
    // kernels with a deep chunk prologue (K7 re-reads 6 planes per chunk) prefer longer chunks as long
    // as a problem alone still fills the GPU several times over
    const int pref = D == 2 ? KernelZc<K>::get(D) : KernelZc3<K>::value;
    if (pref > zc && tiles * ((nz + pref - 1) / pref) >= 148LL * 8) zc = pref;
    // Small problems (BASELINE config 1: a 1788 x 3180 image is 13 tiles x 224 chunks of 8 rows = 2.5 waves of resident
    // thread blocks): pick the chunk length whose block count fills WHOLE waves.  Cost model: waves x (planes per block
    // + 3 planes of prologue); measured on that image (profiles/r2_cfg1_chunks.txt): 20 rows per block = one full wave,
    // 0.068 ms per launch against 0.074 with 8.  Still a function of the problem alone, not of nb.
    if (wave_aware_enabled() && tiles * ((nz + 7) / 8) <= 6LL * 148 * 8) {
      size_t smem_k = 0;
      if constexpr (KernelIsTma<K>::value)
        smem_k = K::smem_bytes(L.ntx, L.nty);
      else
        smem_k = (size_t)K::NROWS * C::row_elems(L.ntx, L.nty) * sizeof(T);
      const long long slots = 148LL * std::max(1, backend::occupancy<K>(L.ntx * L.nty, smem_k));
      long long best_cost = -1;
      for (int cand = 8; cand <= std::min(64, std::max(8, nz)); ++cand) {
        const long long blocks = tiles * ((nz + cand - 1) / cand);
        const long long cost = ((blocks + slots - 1) / slots) * (cand + 3);
        if (best_cost < 0 || cost < best_cost) {
          best_cost = cost;
          zc = cand;
        }
      }
    }
  }
  zc = std::min(zc, nz);
  g.zchunk = zc;
  L.gx = g.xtiles * g.ytiles;
  L.gy = (nz + zc - 1) / zc;
  L.gz = g.nb;
  if constexpr (KernelIsTma<K>::value)
    L.smem = K::smem_bytes(L.ntx, L.nty);
  else
    L.smem = (size_t)K::NROWS * C::row_elems(L.ntx, L.nty) * sizeof(T);
  L.g = g;
  return L;
}

// device scratch carved from the caller's workspace
struct Scratch {
  double* partials;
  unsigned* counters;
  int* ndone;
  unsigned char* fields;  // start of the field area
  size_t partial_cap;     // doubles
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// upper bound of (blocks per launch) * NRED over every kernel configuration we use
template <typename T, int D>
static size_t partial_capacity(const Geom& g, const Tune& tune) {
  // the fused kernel has the widest halo (fewest interior columns per block) => most blocks
  LaunchCfg a = plan_launch<KFusedJ1<T, D, false, false>, T, D>(g, tune, true);
  LaunchCfg b = plan_launch<KSweep<T, D, false>, T, D>(g, tune, false);
  LaunchCfg c = plan_launch<KShrink<T, D, false, false>, T, D>(g, tune, false);
  LaunchCfg d = plan_launch<KDot<T, D, 0>, T, D>(g, tune, false);
  LaunchCfg e = plan_launch<KCgFused<T, D, false>, T, D>(g, tune, true);
  LaunchCfg ch = plan_launch<KChamb<T, D, 0>, T, D>(g, tune, true);
  size_t m = (size_t)ch.gx * ch.gy * ch.gz * 2;
  if constexpr (D == 3) {
    LaunchCfg tm = plan_launch<KFusedJ1T<T, false>, T, D>(g, tune, true);
    m = std::max(m, (size_t)tm.gx * tm.gy * tm.gz * 2);
    LaunchCfg ts = plan_launch<KStencilT<T, TS_PQ>, T, D>(g, tune, true);
    m = std::max(m, (size_t)ts.gx * ts.gy * ts.gz * 2);
  }
  if constexpr (D == 2) {
    LaunchCfg x2 = plan_launch<KFusedJ1x2<T, false, false>, T, D>(g, tune, true);
    m = std::max(m, (size_t)x2.gx * x2.gy * x2.gz * 2);
    LaunchCfg tm = plan_launch<KFusedJ1T<T, false, 2>, T, D>(g, tune, true);
    m = std::max(m, (size_t)tm.gx * tm.gy * tm.gz * 2);
  }
  for (const LaunchCfg* L : {&a, &b, &c, &d}) m = std::max(m, (size_t)L->gx * L->gy * L->gz * 2);
  m = std::max(m, (size_t)e.gx * e.gy * e.gz * 4);
  return m + 64;
}

static size_t scratch_bytes(size_t partial_cap, int nb) {
  return align_up(partial_cap * sizeof(double), 256) + align_up((size_t)nb * sizeof(unsigned), 256) + 256;
}

static Scratch carve_scratch(void* ws, size_t partial_cap, int nb) {
  Scratch s;
  unsigned char* p = reinterpret_cast<unsigned char*>(ws);
  s.partials = reinterpret_cast<double*>(p);
  p += align_up(partial_cap * sizeof(double), 256);
  s.counters = reinterpret_cast<unsigned*>(p);
  p += align_up((size_t)nb * sizeof(unsigned), 256);
  s.ndone = reinterpret_cast<int*>(p);
  p += 256;
  s.fields = p;
  s.partial_cap = partial_cap;
  return s;
}

template <class K, typename T, int D>
static int run(typename K::Args& a, const Geom& g, const Tune& tune, bool big_tile, const Scratch& sc,
               void* stream) {
  LaunchCfg L = plan_launch<K, T, D>(g, tune, big_tile);
  a.g = L.g;
  a.g.rev = next_direction();
  a.g.pf = prefetch_override() >= 0 ? prefetch_override() : KernelPf<K>::get(D);
  if (K::NRED > 0 && (size_t)L.gx * L.gy * L.gz * K::NRED > sc.partial_cap)
    return fail(TVDK_ERR_WORKSPACE, "internal: partial buffer too small");
  ++g_launches;
  return backend::launch<K>(a, L.gx, L.gy, L.gz, L.ntx, L.nty, L.smem, stream, sc.partials, sc.counters);
}

// TMA-staged kernels (3-D anisotropic fused Jacobi-1): on by default, TVDK_TMA=0 selects the march kernel (A/B runs)
static bool tma_enabled() {
  static const int v = env_int("TVDK_TMA", 1);
  return v != 0;
}

// one field as a 3-D tensor (x: nx of pitch, y: ny, z: nb * n0 planes) with a (bx, by, 1) box per plane
template <typename T>
static int field_map(TmaMap* m, const T* base, const Geom& g, int bx, int by) {
  return backend::tma_encode(m, base, (int)sizeof(T), g.nx, g.ny, (long long)g.nb * g.n0, (long long)g.pitch * (long long)sizeof(T),
                             g.s0 * (long long)sizeof(T), bx, by);
}

// ... and for 2-D images (anisotropic, scalar coefficients; BASELINE config 1): off by default until measured, TVDK_TMA2D=1
#ifndef TVDK_TMA2D_DEFAULT
#define TVDK_TMA2D_DEFAULT 0
#endif
static bool tma2d_enabled() {
  static const int v = env_int("TVDK_TMA2D", TVDK_TMA2D_DEFAULT);
  return v != 0;
}

// K1 for problems with scalar coefficients: the TMA-staged kernel (3-D; 2-D anisotropic)
template <typename T, bool ISO, int D = 3>
static int run_fj1_tma(const Geom& g, const Tune& tune, const Coefs<T>& k, const T* f, const T* u_in, T* u_out, const T* t_in,
                       T* t_out, long long tcs_in, long long tcs_out, const tvdk_stat_t* skip, const Fin& fin, const Scratch& sc,
                       void* stream) {
  typedef KFusedJ1T<T, ISO, D> K;
  LaunchCfg L = plan_launch<K, T, D>(g, tune, true);
  static const bool trace = env_int("TVDK_TRACE_ROUTE", 0) != 0;
  if (trace) fprintf(stderr, "[route] KFusedJ1T D=%d ntx=%d nty=%d grid=%d x %d x %d smem=%zu zchunk=%d\n", D, L.ntx, L.nty, L.gx, L.gy, L.gz, L.smem, L.g.zchunk);
  typename K::Args a;
  memset(&a, 0, sizeof a);
  a.g = L.g;
  a.g.rev = next_direction();
  a.k = k;
  a.u_out = u_out;
  a.t_out = t_out;
  a.tcs_out = tcs_out;
  a.skip = skip;
  a.fin = fin;
  TVDK_TRY(field_map(&a.tm_u, u_in, L.g, K::box_w(0, L.ntx), K::box_h(0, L.nty)));
  TVDK_TRY(field_map(&a.tm_f, f, L.g, K::box_w(1, L.ntx), K::box_h(1, L.nty)));
  TVDK_TRY(field_map(&a.tm_t0, t_in, L.g, K::box_w(2, L.ntx), K::box_h(2, L.nty)));
  // (2-D: there is no y component; its map stays a valid descriptor that is never used)
  TVDK_TRY(field_map(&a.tm_t1, t_in + (D == 3 ? tcs_in : 0), L.g, K::box_w(3, L.ntx), K::box_h(3, L.nty)));
  TVDK_TRY(field_map(&a.tm_t2, t_in + (D - 1) * tcs_in, L.g, K::box_w(4, L.ntx), K::box_h(4, L.nty)));
  if ((size_t)L.gx * L.gy * L.gz * K::NRED > sc.partial_cap) return fail(TVDK_ERR_WORKSPACE, "internal: partial buffer too small");
  ++g_launches;
  return backend::launch_tma<K>(a, L.gx, L.gy, L.gz, L.ntx, L.nty, L.smem, stream, sc.partials, sc.counters);
}

// the fused Jacobi(1) iteration, whichever kernel serves this problem class
template <typename T, int D>
static int run_fj1(bool iso, bool het, const Geom& g, const Tune& tune, const Coefs<T>& k, const T* f, const T* u_in, T* u_out,
                   const T* t_in, T* t_out, long long tcs_in, long long tcs_out, const tvdk_stat_t* skip, const Fin& fin,
                   const Scratch& sc, void* stream);

// run f(std::bool_constant<a>, std::bool_constant<b>) -- maps runtime flags onto template args
template <class F>
static int dispatch2(bool a, bool b, F&& f) {
  if (a) {
    if (b) return f(std::true_type{}, std::true_type{});
    return f(std::true_type{}, std::false_type{});
  }
  if (b) return f(std::false_type{}, std::true_type{});
  return f(std::false_type{}, std::false_type{});
}

static Fin no_fin() {
  Fin f;
  memset(&f, 0, sizeof f);
  f.kind = FIN_NONE;
  f.eps = -1.0;
  return f;
}

template <typename T, int D>
static int run_fj1(bool iso, bool het, const Geom& g, const Tune& tune, const Coefs<T>& k, const T* f, const T* u_in, T* u_out,
                   const T* t_in, T* t_out, long long tcs_in, long long tcs_out, const tvdk_stat_t* skip, const Fin& fin,
                   const Scratch& sc, void* stream) {
  if constexpr (D == 3) {
    // (a TMA box holds at most 256 elements per axis: wider forced tiles stay on the march kernel)
    const LaunchCfg probe = plan_launch<KFusedJ1T<T, false>, T, 3>(g, tune, true);
    if (!het && tma_enabled() && (probe.ntx + 2) * Vec<T>::N <= 256 && probe.nty + 2 <= 256) {
      // The isotropic variant measured SLOWER than the march kernel (profiles/r2_tma_iso.txt, 512^3: fp64 0.80 vs 0.83 of
      // the HBM peak, fp32 0.66 vs 0.78: forming phi one plane ahead costs more than the halo rows it removes), so it only
      // runs when every variant is forced (fused = 2: tests) or TVDK_TMA_ISO=1.
      static const int iso_env = env_int("TVDK_TMA_ISO", 0);
      if (iso && (g_force_variants || iso_env))
        return run_fj1_tma<T, true>(g, tune, k, f, u_in, u_out, t_in, t_out, tcs_in, tcs_out, skip, fin, sc, stream);
      if (!iso) return run_fj1_tma<T, false>(g, tune, k, f, u_in, u_out, t_in, t_out, tcs_in, tcs_out, skip, fin, sc, stream);
    }
  }
  if constexpr (D == 2) {
    if (!het && !iso && tma_enabled() && (tma2d_enabled() || g_force_variants || g_force_tma2d))
      return run_fj1_tma<T, false, 2>(g, tune, k, f, u_in, u_out, t_in, t_out, tcs_in, tcs_out, skip, fin, sc, stream);
  }
  return dispatch2(iso, het, [&](auto I, auto H) {
    typedef KFusedJ1<T, D, decltype(I)::value, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.f = f; a.u_in = u_in; a.u_out = u_out;
    a.t_in = t_in; a.t_out = t_out; a.tcs_in = tcs_in; a.tcs_out = tcs_out;
    a.skip = skip; a.fin = fin;
    return run<K, T, D>(a, g, tune, true, sc, stream);
  });
}

// K5 / K4 in 3-D with scalar coefficients: the TMA-staged single-field stencil kernel
template <typename T, int MODE>
static int run_stencil_tma(const Geom& g, const Tune& tune, const Coefs<T>& k, const T* a_in, const T* b_in, T* out, DRef coef,
                           IRef first, int first_all, const tvdk_stat_t* skip, const Fin& fin, const Scratch& sc, void* stream) {
  typedef KStencilT<T, MODE> K;
  LaunchCfg L = plan_launch<K, T, 3>(g, tune, true);
  typename K::Args a;
  memset(&a, 0, sizeof a);
  a.g = L.g;
  a.g.rev = next_direction();
  a.k = k;
  a.out = out;
  a.coef = coef;
  a.first = first;
  a.first_all = first_all;
  a.skip = skip;
  a.fin = fin;
  // (on the first CG spin the old p is never read; its map still has to be a valid descriptor)
  TVDK_TRY(field_map(&a.tm_a, a_in ? a_in : b_in, L.g, K::box_w(0, L.ntx), K::box_h(0, L.nty)));
  TVDK_TRY(field_map(&a.tm_b, b_in, L.g, K::box_w(1, L.ntx), K::box_h(1, L.nty)));
  if (K::NRED > 0 && (size_t)L.gx * L.gy * L.gz * K::NRED > sc.partial_cap)
    return fail(TVDK_ERR_WORKSPACE, "internal: partial buffer too small");
  ++g_launches;
  return backend::launch_tma<K>(a, L.gx, L.gy, L.gz, L.ntx, L.nty, L.smem, stream, sc.partials, sc.counters);
}

template <typename T>
static bool stencil_tma_ok(const Geom& g, const Tune& tune) {
  if (!tma_enabled()) return false;
  const LaunchCfg probe = plan_launch<KStencilT<T, TS_PQ>, T, 3>(g, tune, true);
  return (probe.ntx + 2) * Vec<T>::N <= 256 && probe.nty + 2 <= 256;
}

// K5: p_out = first ? r : beta p_in + r ; p_out . (A p_out)
template <typename T, int D>
static int run_cgpq(bool het, const Geom& g, const Tune& tune, const Coefs<T>& k, const T* p_in, const T* r_in, T* p_out, DRef beta,
                    IRef first, int first_all, const tvdk_stat_t* skip, const Fin& fin, const Scratch& sc, void* stream) {
  if constexpr (D == 3) {
    if (!het && stencil_tma_ok<T>(g, tune))
      return run_stencil_tma<T, TS_PQ>(g, tune, k, p_in, r_in, p_out, beta, first, first_all, skip, fin, sc, stream);
  }
  return dispatch2(het, false, [&](auto H, auto) {
    typedef KCg<T, D, CG_PQ, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.p_in = p_in; a.r_in = r_in; a.p_out = p_out;
    a.coef = beta; a.first = first; a.first_all = first_all;
    a.skip = skip; a.fin = fin;
    return run<K, T, D>(a, g, tune, false, sc, stream);
  });
}

// K4: one Jacobi sweep (the TMA kernel has no tol-branch sums: fin must be empty for it)
template <typename T, int D>
static int run_sweep(bool het, const Geom& g, const Tune& tune, const Coefs<T>& k, const T* x_in, const T* rhs, T* x_out,
                     const tvdk_stat_t* skip, const Fin& fin, const Scratch& sc, void* stream) {
  if constexpr (D == 3) {
    // (fp32 only: 0.92 vs 0.78 of the HBM peak at 512^3; in fp64 the march kernel's 0.95 beats the staged 0.93,
    //  profiles/r2_stencil_ab.txt)
    if (sizeof(T) == 4 && !het && fin.kind == FIN_NONE && !fin.sum_out && stencil_tma_ok<T>(g, tune)) {
      DRef none;
      none.p = nullptr;
      none.stride = 0;
      IRef nof;
      nof.p = nullptr;
      nof.stride = 0;
      return run_stencil_tma<T, TS_SWEEP>(g, tune, k, x_in, rhs, x_out, none, nof, 0, skip, fin, sc, stream);
    }
  }
  return dispatch2(het, false, [&](auto H, auto) {
    typedef KSweep<T, D, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.x_in = x_in; a.rhs = rhs; a.x_out = x_out; a.skip = skip; a.fin = fin;
    return run<K, T, D>(a, g, tune, true, sc, stream);  // 3-D: 16-row tiles (profiles/r1_tile3d.txt)
  });
}

// ---------------------------------------------------------------------------------------
// the outer loop
// ---------------------------------------------------------------------------------------
struct FieldPlan {
  // number of extra fields taken from the workspace
  bool fusedj1, cg, cgfused, jac, expl, adapt;
  int n_fields;
};

static FieldPlan field_plan(const tvdk_problem_t* p) {
  FieldPlan fp;
  fp.expl = p->x0_b != nullptr || p->x0_d != nullptr;
  fp.cg = p->solver_kind == TVDK_SOLVER_CG;
  fp.adapt = false;
  if (p->adaptive_mask)
    for (int i = 0; i < p->max_num_iter; ++i) fp.adapt = fp.adapt || p->adaptive_mask[i] != 0;
  // the adaptive schedule runs on the building-block route: Jacobi keeps the diagonal of the
  // initial ell (jacobi.py:101-106 caches it), rhs and shrink use the current one
  fp.fusedj1 = p->fused && !fp.cg && p->solver_maxiter == 1 && p->solver_tol < 0.0 && !fp.adapt;
  // single-pass CG iterations (K7): fixed iteration count only -- with a tolerance the stopping
  // test would act on a predicted instead of a summed residual norm
  // ... and 2-D only: in 3-D its two halo rows on each side of y cost more (0.5 of the HBM peak
  // measured, profiles/r1_pf_sweep2.txt) than the 2 words per iteration it saves
  // ... and not with the adaptive ell schedule: ell = 1/|grad u| makes the systems so ill-conditioned that the
  // predicted rho costs 3-4 more digits than the directly summed one (2e-7 vs 1e-12 against the oracle)
  fp.cgfused = fp.cg && p->fused && p->solver_tol == 0.0 && (p->geom.dim == 2 || p->fused >= 2) && !fp.adapt;
  fp.jac = !fp.cg;
  const int D = p->geom.dim;
  int n = 0;
  if (fp.adapt) n += 1 + 2 * D;  // ell field, explicit d and b after an update of ell
  if (fp.fusedj1) {
    n += 1 + D;           // U2, T2
    if (fp.expl) n += 2;  // rhs + xa for the explicit first iteration
  } else if (fp.cg) {
    n += fp.cgfused ? 5 : 4;  // X, R, P0, P1 (+ R1: the single-pass route ping-pongs r as well)
  } else {
    n += (p->solver_maxiter >= 2) ? 3 : 2;  // rhs, XA, (XB)
  }
  fp.n_fields = n;
  return fp;
}

template <typename T, int D>
static size_t workspace_bytes_impl(const tvdk_problem_t* p) {
  Geom g = make_geom(p->geom);
  Tune tune = merge_tune(p->tile_x, p->tile_y, p->zchunk);
  const size_t cap = partial_capacity<T, D>(g, tune);
  const FieldPlan fp = field_plan(p);
  const size_t field_bytes = align_up((size_t)g.sb * g.nb * sizeof(T), 256);
  return scratch_bytes(cap, g.nb) + fp.n_fields * field_bytes + 256;
}

template <typename T, int D>
static int solve_impl(const tvdk_problem_t* p, void* stream, int* iters_launched) {
  const Geom g = make_geom(p->geom);
  const Tune tune = merge_tune(p->tile_x, p->tile_y, p->zchunk);
  const bool iso = p->isotropic != 0;
  const FieldPlan fp = field_plan(p);
  const size_t need = workspace_bytes_impl<T, D>(p);
  if (!p->workspace || p->workspace_bytes < need)
    return fail(TVDK_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", need, p->workspace_bytes);
  if (!p->f || !p->u || !p->t || !p->stats) return fail(TVDK_ERR_INVALID, "f, u, t and stats must be non-NULL");
  if (fp.expl && (!p->x0_b || !p->x0_d)) return fail(TVDK_ERR_INVALID, "warm start needs both x0_d and x0_b");
  if (p->solver_maxiter < 1) return fail(TVDK_ERR_INVALID, "solver maxiter must be >= 1");
  if (fp.cg && p->solver_tol < 0.0)
    return fail(TVDK_ERR_INVALID, "CG needs a numeric tol (SciPy rejects rtol=None); use 0.0 for a fixed count");
  if (p->geom.z_begin != 0 || p->geom.z_end != p->geom.n[0] || p->geom.g0_off != 0 || p->geom.g0_n != p->geom.n[0])
    return fail(TVDK_ERR_INVALID, "tvdk_solve works on whole volumes; slabs use the building-block entry points");

  const size_t cap = partial_capacity<T, D>(g, tune);
  const Scratch sc = carve_scratch(p->workspace, cap, g.nb);
  const size_t fe = align_up((size_t)g.sb * g.nb * sizeof(T), 256) / sizeof(T);  // elements per field slot
  T* wsf = reinterpret_cast<T*>(sc.fields);
  int next_field = 0;
  auto take = [&](int n) {
    T* r = wsf + (size_t)next_field * fe;
    next_field += n;
    return r;
  };

  Coefs<T> k = make_coefs<T>(&p->mu, &p->omega, &p->ell, D);  // current coefficients (ell may adapt)
  const Coefs<T> k0 = k;  // Jacobi sweeps keep the initial ell's diagonal (jacobi.py:101-106)
  const T* f = reinterpret_cast<const T*>(p->f);
  T* ubuf[2] = {reinterpret_cast<T*>(p->u), nullptr};
  T* tbuf[2] = {reinterpret_cast<T*>(p->t), nullptr};
  long long tcs[2] = {p->t_comp_stride, (long long)fe};
  T *RB = nullptr, *XA = nullptr, *XB = nullptr, *X = nullptr, *P[2] = {nullptr, nullptr};
  T *ELL = nullptr, *XD = nullptr, *XBB = nullptr, *R1 = nullptr;
  if (fp.adapt) {
    ELL = take(1);
    XD = take(D);
    XBB = take(D);
  }
  // explicit d, b of the next iteration: the caller's warm start, or the expansion made after
  // an update of ell (the state is compressed with the kappa it was shrunk with)
  const T* xd_cur = reinterpret_cast<const T*>(p->x0_d);
  const T* xb_cur = reinterpret_cast<const T*>(p->x0_b);
  long long xcs_cur = p->x0_comp_stride;
  if (fp.fusedj1) {
    ubuf[1] = take(1);
    tbuf[1] = take(D);
    if (fp.expl) {
      RB = take(1);
      XA = take(1);
    }
  } else if (fp.cg) {
    X = take(1);
    RB = take(1);
    P[0] = take(1);
    P[1] = take(1);
    if (fp.cgfused) R1 = take(1);
  } else {
    RB = take(1);
    XA = take(1);
    if (p->solver_maxiter >= 2) XB = take(1);
  }

  tvdk_stat_t* st = p->stats;
  TVDK_TRY(backend::memset_async(st, 0, sizeof(tvdk_stat_t) * (size_t)g.nb, stream));
  TVDK_TRY(backend::memset_async(sc.counters, 0, sizeof(unsigned) * (size_t)g.nb, stream));
  TVDK_TRY(backend::memset_async(sc.ndone, 0, sizeof(int), stream));

  {  // ||img|| (split_bregman_tvd.py:58)
    typename KDot<T, D, 0>::Args a;
    memset(&a, 0, sizeof a);
    a.x = f;
    a.y = f;
    a.fin = no_fin();
    a.fin.kind = FIN_NRMF;
    a.fin.stats = st;
    TVDK_TRY((run<KDot<T, D, 0>, T, D>(a, g, tune, false, sc, stream)));
  }

  auto outer_fin = [&](int it, int buf_id) {
    Fin fn = no_fin();
    fn.kind = FIN_OUTER;
    fn.stats = st;
    fn.eps = p->eps;
    fn.history = p->history;
    fn.iter = it;
    fn.nb = g.nb;
    fn.ndone = sc.ndone;
    fn.buf_id = buf_id;
    return fn;
  };

  bool het = k.sc.hetero != 0;
  // the Jacobi sweeps only see omega and ell (initial values, k0): a per-voxel mu alone -- the usual
  // "weight = mask * mu" case, examples/regularization.py:13-22 -- leaves them on the scalar kernels
  const bool het0 = k0.omega.p != nullptr || k0.ell.p != nullptr;

  // shrink + Bregman update + in-place u update, shared by the unfused routes
  const int stat_dstride = (int)(sizeof(tvdk_stat_t) / sizeof(double));
  auto launch_shrink = [&](int it, int cur, const T* u0, const T* u1, const T* u2, int use_sel, bool expl_now,
                           const T* axpy_p = nullptr) {
    const Fin fn = outer_fin(it, cur);
    return dispatch2(iso, het, [&](auto I, auto H) {
      typedef KShrink<T, D, decltype(I)::value, decltype(H)::value> K;
      typename K::Args a;
      memset(&a, 0, sizeof a);
      a.k = k; a.unew[0] = u0; a.unew[1] = u1; a.unew[2] = u2; a.use_sel = use_sel;
      a.axpy_p = axpy_p; a.axpy_alpha.p = &st[0].alpha; a.axpy_alpha.stride = stat_dstride;
      a.cg_mode = fp.cg ? 1 : 0;
      a.u_old = ubuf[cur]; a.u_out = ubuf[cur];
      a.t_in = expl_now ? nullptr : tbuf[cur]; a.t_out = tbuf[cur]; a.tcs = tcs[cur];
      a.xb = xb_cur; a.xcs = xcs_cur;
      a.skip = st; a.fin = fn;
      return run<K, T, D>(a, g, tune, false, sc, stream);
    });
  };

#define TVDK_RHS_ARGS(a, cur, expl_now)                                   \
  memset(&a, 0, sizeof a);                                                \
  a.k = k; a.f = f;                                                       \
  a.t = (expl_now) ? nullptr : tbuf[cur]; a.tcs = tcs[cur];               \
  a.xd = xd_cur; a.xb = xb_cur; a.xcs = xcs_cur;                          \
  a.u = ubuf[cur]; a.skip = st; a.fin = no_fin();

  auto launch_rhs_sweep = [&](int cur, bool expl_now, T* rhs, T* x, const Fin& fn) {
    return dispatch2(iso, het, [&](auto I, auto H) {
      typedef KRhs<T, D, decltype(I)::value, RHS_SWEEP, decltype(H)::value> K;
      typename K::Args a;
      TVDK_RHS_ARGS(a, cur, expl_now)
      a.rhs = rhs; a.x = x; a.fin = fn;
      return run<K, T, D>(a, g, tune, false, sc, stream);
    });
  };
  auto launch_rhs_cginit = [&](int cur, bool expl_now, T* x, T* r, const Fin& fn) {
    return dispatch2(iso, het, [&](auto I, auto H) {
      typedef KRhs<T, D, decltype(I)::value, RHS_CGINIT, decltype(H)::value> K;
      typename K::Args a;
      TVDK_RHS_ARGS(a, cur, expl_now)
      a.x = x; a.r = r; a.fin = fn;
      return run<K, T, D>(a, g, tune, false, sc, stream);
    });
  };

  auto launch_rhs_only = [&](int cur, bool expl_now, T* rhs) {
    return dispatch2(iso, het, [&](auto I, auto H) {
      typedef KRhs<T, D, decltype(I)::value, RHS_ONLY, decltype(H)::value> K;
      typename K::Args a;
      TVDK_RHS_ARGS(a, cur, expl_now)
      a.rhs = rhs;
      return run<K, T, D>(a, g, tune, false, sc, stream);
    });
  };

  const bool tolmode = !fp.cg && p->solver_tol >= 0.0;
  bool expl_next = fp.expl;
  const int check_every = p->check_every > 0 ? p->check_every : 8;
  int cur = 0, launched = 0, slot = 0;
  bool pending[2] = {false, false};
  bool all_done = false;

  for (int it = 0; it < p->max_num_iter && !all_done; ++it) {
    const bool expl_now = expl_next;
    expl_next = false;
    if (p->verbose_every > 0 && it % p->verbose_every == 0 && p->energy_history) {
      // energy of the current iterate (split_bregman_tvd.py:162-167 prints _functional(img_iter))
      typename KEnergy<T, D>::Args a;
      memset(&a, 0, sizeof a);
      a.k = k; a.x = ubuf[cur]; a.f = f; a.skip = st;
      a.fin = no_fin();
      a.fin.kind = FIN_ENERGY; a.fin.stats = st; a.fin.history = p->energy_history; a.fin.iter = it; a.fin.nb = g.nb;
      TVDK_TRY((run<KEnergy<T, D>, T, D>(a, g, tune, false, sc, stream)));
    }
    bool pair = false;
    if constexpr (D == 2) {
      // two outer iterations per pass (K1x2) wherever nothing has to happen between them
      // Measured (profiles/r1_x2.txt, 4000x7000x3): 0.74 vs 0.82 ms per iteration for anisotropic fp64 with
      // scalar coefficients; slower than K1 for isotropic (sqrt + divide in two of the four stages), per-voxel
      // maps and fp32 (the pipeline holds > 200 registers per thread) and for small images => only the
      // winning case takes it by default, fused = 2 forces it everywhere (tests).
      const bool wins = std::is_same<T, double>::value && !iso && !het && (long long)g.n0 * g.nx >= (1LL << 23);
      pair = fp.fusedj1 && !expl_now && p->eps < 0.0 && p->verbose_every <= 0 && it + 1 < p->max_num_iter &&
             !(fp.adapt) && p->fused != 3 && (wins || p->fused == 2);
      if (pair) {
        Fin fn = outer_fin(it, 1 - cur);
        fn.kind = FIN_OUTER2;
        TVDK_TRY(dispatch2(iso, het, [&](auto I, auto H) {
          typedef KFusedJ1x2<T, decltype(I)::value, decltype(H)::value> K;
          typename K::Args a;
          memset(&a, 0, sizeof a);
          a.k = k; a.f = f; a.u_in = ubuf[cur]; a.u_out = ubuf[1 - cur];
          a.t_in = tbuf[cur]; a.t_out = tbuf[1 - cur]; a.tcs_in = tcs[cur]; a.tcs_out = tcs[1 - cur];
          a.skip = st; a.fin = fn;
          return run<K, T, D>(a, g, tune, true, sc, stream);
        }));
        cur = 1 - cur;
        ++it;  // this launch was iterations it and it + 1
        ++launched;
      }
    }
    if (pair) {
      // (both iterations done above)
    } else if (fp.fusedj1 && !expl_now) {
      const Fin fn = outer_fin(it, 1 - cur);
      TVDK_TRY((run_fj1<T, D>(iso, het, g, tune, k, f, ubuf[cur], ubuf[1 - cur], tbuf[cur], tbuf[1 - cur], tcs[cur], tcs[1 - cur], st,
                              fn, sc, stream)));
      cur = 1 - cur;
    } else if (fp.cg) {
      Fin fn = no_fin();
      fn.kind = FIN_CGINIT; fn.stats = st; fn.tol = p->solver_tol;
      TVDK_TRY(launch_rhs_cginit(cur, expl_now, X, RB, fn));
      if (fp.cgfused) {
        // iteration 0: p_0 = r_0 is never written; alpha_0 and beta_1 come from r.(A r), (A r).(A r)
        TVDK_TRY(dispatch2(het, false, [&](auto H, auto) {
          typedef KCg<T, D, CG_FIRST, decltype(H)::value> K;
          typename K::Args a;
          memset(&a, 0, sizeof a);
          a.k = k; a.p_in = RB; a.skip = st; a.fin = no_fin(); a.fin.kind = FIN_CGFIRST; a.fin.stats = st;
          return run<K, T, D>(a, g, tune, false, sc, stream);
        }));
        T* R[2] = {RB, R1};
        const T* p_last = RB;
        for (int i = 1; i < p->solver_maxiter; ++i) {
          TVDK_TRY(dispatch2(het, false, [&](auto H, auto) {
            typedef KCgFused<T, D, decltype(H)::value> K;
            typename K::Args a;
            memset(&a, 0, sizeof a);
            a.k = k; a.p_in = p_last; a.r_in = R[(i - 1) & 1]; a.p_out = P[i & 1]; a.r_out = R[i & 1]; a.x = X;
            a.alpha.p = &st[0].alpha; a.alpha.stride = stat_dstride;
            a.beta.p = &st[0].beta; a.beta.stride = stat_dstride;
            a.skip = st; a.fin = no_fin(); a.fin.kind = FIN_CGFUSED; a.fin.stats = st;
            return run<K, T, D>(a, g, tune, true, sc, stream);
          }));
          p_last = P[i & 1];
        }
        // x_m = x_{m-1} + alpha_{m-1} p_{m-1} is formed inside the shrink kernel
        TVDK_TRY(launch_shrink(it, cur, X, nullptr, nullptr, 0, expl_now, p_last));
      } else {
      for (int i = 0; i < p->solver_maxiter; ++i) {
        {
          DRef beta;
          beta.p = &st[0].beta; beta.stride = (int)(sizeof(tvdk_stat_t) / sizeof(double));
          IRef frst;
          frst.p = &st[0].cg_first; frst.stride = (int)(sizeof(tvdk_stat_t) / sizeof(int));
          Fin fq = no_fin();
          fq.kind = FIN_CGPQ; fq.stats = st;
          TVDK_TRY((run_cgpq<T, D>(het, g, tune, k, P[(i & 1) ^ 1], RB, P[i & 1], beta, frst, 0, st, fq, sc, stream)));
        }
        TVDK_TRY(dispatch2(het, false, [&](auto H, auto) {
          typedef KCg<T, D, CG_UPDATE, decltype(H)::value> K;
          typename K::Args b;
          memset(&b, 0, sizeof b);
          b.k = k; b.p_in = P[i & 1]; b.x = X; b.r = RB;
          b.coef.p = &st[0].alpha; b.coef.stride = (int)(sizeof(tvdk_stat_t) / sizeof(double));
          b.skip = st; b.fin = no_fin(); b.fin.kind = FIN_CGRR; b.fin.stats = st;
          return run<K, T, D>(b, g, tune, false, sc, stream);
        }));
      }
      TVDK_TRY(launch_shrink(it, cur, X, nullptr, nullptr, 0, expl_now));
      }
    } else {
      // Jacobi: sweep 1 fused with the rhs, then ping-pong XA/XB (jacobi.py:109-131)
      Fin fn = no_fin();
      if (tolmode) {
        fn.kind = FIN_JTOL; fn.stats = st; fn.tol = p->solver_tol; fn.buf_id = 1; fn.first_sweep = 1;
      }
      T* xin = XA;
      T* xout = XB;
      int out_id = 2;
      int s_first = 2;
      if (fp.adapt) {  // rhs with the current ell; every sweep with the initial ell's diagonal
        TVDK_TRY(launch_rhs_only(cur, expl_now, RB));
        xin = ubuf[cur];
        xout = XA;
        out_id = 1;
        s_first = 1;
      } else {
        TVDK_TRY(launch_rhs_sweep(cur, expl_now, RB, XA, fn));
      }
      for (int s = s_first; s <= p->solver_maxiter; ++s) {
        // temporal blocking: S sweeps per pass wherever the plain recurrence allows it (scalar
        // coefficients, no tol branch): 4 at a time in 2-D, 2 at a time in 3-D
        if (!tolmode && !het0 && p->fused && s > 1) {
          const int rem = p->solver_maxiter - s + 1;
          const int S = sweeps_per_pass(D, (long long)g.n0 * g.ny * g.nx, rem, false, p->fused >= 2);
          if (S > 1) {
            auto go = [&](auto SS) {
              typedef KSweepN<T, D, decltype(SS)::value> K;
              typename K::Args a;
              memset(&a, 0, sizeof a);
              a.k = k0; a.x_in = xin; a.rhs = RB; a.x_out = xout; a.skip = st; a.fin = no_fin();
              return run<K, T, D>(a, g, tune, true, sc, stream);
            };
            if (S == 4)
              TVDK_TRY(go(std::integral_constant<int, 4>{}));
            else
              TVDK_TRY(go(std::integral_constant<int, 2>{}));
            std::swap(xin, xout);  // (s > 1: both are XA / XB here)
            out_id = 3 - out_id;
            s += S - 1;
            continue;
          }
        }
        {
          Fin fs = no_fin();
          if (tolmode) {
            fs.kind = FIN_JTOL; fs.stats = st; fs.tol = p->solver_tol; fs.buf_id = out_id;
            fs.first_sweep = (s == 1);
          }
          TVDK_TRY((run_sweep<T, D>(het0, g, tune, k0, xin, RB, xout, st, fs, sc, stream)));
        }
        if (s == 1) {
          xin = XA;
          xout = XB;
          out_id = 2;
        } else {
          std::swap(xin, xout);
          out_id = 3 - out_id;
        }
      }
      if (tolmode)
        TVDK_TRY(launch_shrink(it, cur, ubuf[cur], XA, XB ? XB : XA, 1, expl_now));
      else
        TVDK_TRY(launch_shrink(it, cur, xin, nullptr, nullptr, 0, expl_now));
    }
    if (fp.adapt && p->adaptive_mask[it] && it + 1 < p->max_num_iter) {
      // split_bregman_tvd.py:149-158.  d and b stay what the shrinkage with the OLD kappa made
      // them: expand them from the compressed state before ell changes, the next iteration
      // reads them explicitly.
      TVDK_TRY(dispatch2(iso, false, [&](auto I, auto) {
        typedef KExpand<T, D, decltype(I)::value> K;
        typename K::Args a;
        memset(&a, 0, sizeof a);
        a.k = k; a.t = tbuf[cur]; a.tcs = tcs[cur]; a.d = XD; a.b = XBB; a.ocs = (long long)fe; a.fin = no_fin();
        return run<K, T, D>(a, g, tune, false, sc, stream);
      }));
      {
        typename KAdapt<T, D>::Args a;
        memset(&a, 0, sizeof a);
        a.x = ubuf[cur]; a.ell = ELL; a.skip = st; a.fin = no_fin();
        TVDK_TRY((run<KAdapt<T, D>, T, D>(a, g, tune, false, sc, stream)));
      }
      k.ell.p = ELL;
      k.ell.val = (T)0;
      k.ell.sb = g.sb;
      k.sc.hetero = 1;
      het = true;
      xd_cur = XD;
      xb_cur = XBB;
      xcs_cur = (long long)fe;
      expl_next = true;
    }
    ++launched;
    // convergence polling (split_bregman_tvd.py:173-175) without stalling the queue: the flag
    // read after chunk k is inspected before chunk k+2 is enqueued
    if (p->eps >= 0.0 && (launched % check_every) == 0 && it + 1 < p->max_num_iter) {
      const int other = slot ^ 1;
      TVDK_TRY(backend::poll_begin(slot, sc.ndone, stream));
      pending[slot] = true;
      if (pending[other]) {
        int v = 0;
        TVDK_TRY(backend::poll_end(other, &v));
        pending[other] = false;
        if (v >= g.nb) all_done = true;
      }
      slot = other;
    }
  }
  for (int s2 = 0; s2 < 2; ++s2)
    if (pending[s2]) {
      int v = 0;
      TVDK_TRY(backend::poll_end(s2, &v));
    }
  if (fp.fusedj1) {  // bring every problem's iterate back into p->u
    typename KSelectCopy<T, D>::Args a;
    memset(&a, 0, sizeof a);
    a.src = ubuf[1]; a.dst = ubuf[0]; a.stats = st; a.want = 1; a.fin = no_fin();
    TVDK_TRY((run<KSelectCopy<T, D>, T, D>(a, g, tune, false, sc, stream)));
    for (int d = 0; d < D; ++d) {
      typename KSelectCopy<T, D>::Args b;
      memset(&b, 0, sizeof b);
      b.src = tbuf[1] + d * tcs[1]; b.dst = tbuf[0] + d * tcs[0]; b.stats = st; b.want = 1; b.fin = no_fin();
      TVDK_TRY((run<KSelectCopy<T, D>, T, D>(b, g, tune, false, sc, stream)));
    }
  }
  if (iters_launched) *iters_launched = launched;
  return 0;
}

// dispatch helpers ---------------------------------------------------------------------
#define TVDK_DISPATCH(geom, CALL)                                   \
  do {                                                              \
    if ((geom).dtype == TVDK_F64) {                                 \
      if ((geom).dim == 2) { typedef double T; constexpr int D = 2; CALL; } \
      else { typedef double T; constexpr int D = 3; CALL; }         \
    } else {                                                        \
      if ((geom).dim == 2) { typedef float T; constexpr int D = 2; CALL; }  \
      else { typedef float T; constexpr int D = 3; CALL; }          \
    }                                                               \
  } while (0)

template <typename T, int D>
static size_t op_ws_impl(const tvdk_geom_t* gg) {
  Geom g = make_geom(*gg);
  Tune tune = merge_tune(0, 0, 0);
  return scratch_bytes(partial_capacity<T, D>(g, tune), g.nb) + 256;
}

template <typename T, int D>
static int op_scratch(const tvdk_geom_t* gg, void* ws, size_t ws_bytes, Scratch* sc, void* stream) {
  Geom g = make_geom(*gg);
  Tune tune = merge_tune(0, 0, 0);
  const size_t cap = partial_capacity<T, D>(g, tune);
  if (!ws || ws_bytes < scratch_bytes(cap, g.nb))
    return fail(TVDK_ERR_WORKSPACE, "operator workspace of %zu bytes needed, %zu given", scratch_bytes(cap, g.nb), ws_bytes);
  *sc = carve_scratch(ws, cap, g.nb);
  return backend::memset_async(sc->counters, 0, sizeof(unsigned) * (size_t)g.nb, stream);
}

static Fin sum_fin(double* out) {
  Fin f = no_fin();
  f.sum_out = out;
  return f;
}

}  // namespace tvdk

// ---------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------
using namespace tvdk;

// (definitions pick up C linkage from the declarations in include/tvdk.h)

int tvdk_version(void) { return TVDK_VERSION; }
const char* tvdk_last_error(void) { return g_last_error.c_str(); }

int tvdk_device_ok(void) {
  char msg[256] = {0};
  const int ok = backend::device_ok(msg, sizeof msg);
  if (!ok) g_last_error = msg;
  return ok;
}

int64_t tvdk_launch_count(int reset) {
  const long long n = g_launches;
  if (reset) g_launches = 0;
  return n;
}

int tvdk_set_tune(int tile_x, int tile_y, int zchunk) {
  if (tile_x < 0 || tile_y < 0 || zchunk < 0 || (tile_x % 8) != 0) return fail(TVDK_ERR_INVALID, "bad tune values");
  g_tune.tile_x = tile_x;
  g_tune.tile_y = tile_y;
  g_tune.zchunk = zchunk;
  return 0;
}

int tvdk_pitch(int dtype, int n_last) {
  const int a = dtype == TVDK_F64 ? 16 : 32;  // 128 bytes
  return (n_last + a - 1) / a * a;
}

int64_t tvdk_field_elems(const tvdk_geom_t* g) {
  if (check_geom(g)) return -1;
  return make_geom(*g).sb;
}

size_t tvdk_workspace_bytes(const tvdk_problem_t* p) {
  if (!p || check_geom(&p->geom)) return 0;
  size_t r = 0;
  TVDK_DISPATCH(p->geom, r = (workspace_bytes_impl<T, D>(p)));
  return r;
}

size_t tvdk_op_workspace_bytes(const tvdk_geom_t* g) {
  if (check_geom(g)) return 0;
  size_t r = 0;
  TVDK_DISPATCH(*g, r = (op_ws_impl<T, D>(g)));
  return r;
}

int tvdk_solve(const tvdk_problem_t* p, void* stream, int* iters_launched) {
  if (!p) return fail(TVDK_ERR_INVALID, "problem is NULL");
  TVDK_TRY(check_geom(&p->geom));
  if (p->max_num_iter < 0) return fail(TVDK_ERR_INVALID, "max_num_iter must be >= 0");
  if (p->solver_kind != TVDK_SOLVER_JACOBI && p->solver_kind != TVDK_SOLVER_CG)
    return fail(TVDK_ERR_UNSUPPORTED, "unknown solver kind %d (Jacobi and CG only)", p->solver_kind);
  int rc = 0;
  TVDK_RANGE("tvdk_solve");
  g_alternate = alternate_enabled() ? 1 : 0;
  g_force_variants = p->fused == 2 ? 1 : 0;
  g_force_tma2d = p->fused == 4 ? 1 : 0;
  TVDK_DISPATCH(p->geom, rc = (solve_impl<T, D>(p, stream, iters_launched)));
  g_alternate = 0;
  g_force_variants = 0;
  g_force_tma2d = 0;
  return rc;
}

int tvdk_pack(const tvdk_geom_t* gg, int src_dtype, const void* src, void* dst, void* stream) {
  TVDK_TRY(check_geom(gg));
  if (!src || !dst) return fail(TVDK_ERR_INVALID, "NULL pointer");
  Geom g = make_geom(*gg);
  const long long per_row = (long long)g.nx * g.nb;
  const int ntx = 256;
  const int gy = (int)std::min<long long>((per_row + ntx - 1) / ntx, 64);
  const long long rows = (long long)g.n0 * g.ny;
#define TVDK_PACK_CALL(TI, TF)                                                          \
  {                                                                                     \
    typename KPack<TI, TF, 1>::Args a;                                                  \
    memset(&a, 0, sizeof a);                                                            \
    a.g = g; a.src = src; a.dst = dst; a.fin = no_fin();                                \
    ++g_launches;                                                                       \
    return backend::launch<KPack<TI, TF, 1>>(a, (int)rows, gy, 1, ntx, 1, 0, stream, nullptr, nullptr); \
  }
  if (src_dtype == TVDK_F64 && gg->dtype == TVDK_F64) TVDK_PACK_CALL(double, double)
  if (src_dtype == TVDK_F64 && gg->dtype == TVDK_F32) TVDK_PACK_CALL(double, float)
  if (src_dtype == TVDK_F32 && gg->dtype == TVDK_F64) TVDK_PACK_CALL(float, double)
  if (src_dtype == TVDK_F32 && gg->dtype == TVDK_F32) TVDK_PACK_CALL(float, float)
#undef TVDK_PACK_CALL
  return fail(TVDK_ERR_INVALID, "bad dtype");
}

int tvdk_unpack(const tvdk_geom_t* gg, int dst_dtype, const void* src, void* dst, void* stream) {
  TVDK_TRY(check_geom(gg));
  if (!src || !dst) return fail(TVDK_ERR_INVALID, "NULL pointer");
  Geom g = make_geom(*gg);
  const long long per_row = (long long)g.nx * g.nb;
  const int ntx = 256;
  const int gy = (int)std::min<long long>((per_row + ntx - 1) / ntx, 64);
  const long long rows = (long long)g.n0 * g.ny;
#define TVDK_UNPACK_CALL(TI, TF)                                                        \
  {                                                                                     \
    typename KPack<TI, TF, 0>::Args a;                                                  \
    memset(&a, 0, sizeof a);                                                            \
    a.g = g; a.src = src; a.dst = dst; a.fin = no_fin();                                \
    ++g_launches;                                                                       \
    return backend::launch<KPack<TI, TF, 0>>(a, (int)rows, gy, 1, ntx, 1, 0, stream, nullptr, nullptr); \
  }
  if (dst_dtype == TVDK_F64 && gg->dtype == TVDK_F64) TVDK_UNPACK_CALL(double, double)
  if (dst_dtype == TVDK_F64 && gg->dtype == TVDK_F32) TVDK_UNPACK_CALL(double, float)
  if (dst_dtype == TVDK_F32 && gg->dtype == TVDK_F64) TVDK_UNPACK_CALL(float, double)
  if (dst_dtype == TVDK_F32 && gg->dtype == TVDK_F32) TVDK_UNPACK_CALL(float, float)
#undef TVDK_UNPACK_CALL
  return fail(TVDK_ERR_INVALID, "bad dtype");
}

// ---- operators ---------------------------------------------------------------------------
// Every operator: check, carve the scratch, build Args, dispatch (isotropic, heterogeneous) onto
// the template instantiation, launch.
namespace tvdk {

template <typename T, int D>
struct OpCtx {
  Geom g;
  Tune tune;
  Scratch sc;
  int init(const tvdk_geom_t* gg, void* ws, size_t wsb, void* stream) {
    g = make_geom(*gg);
    tune = merge_tune(0, 0, 0);
    return op_scratch<T, D>(gg, ws, wsb, &sc, stream);
  }
};

template <typename T, int D>
static int rhs_impl(const tvdk_geom_t* gg, int iso, const tvdk_coef_t* mu, const tvdk_coef_t* om,
                    const tvdk_coef_t* el, const void* f, const void* t, int64_t tcs, void* rhs, void* ws,
                    size_t wsb, void* stream, const void* xd = nullptr, const void* xb = nullptr, int64_t xcs = 0) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(mu, om, el, D);
  return dispatch2(iso != 0, k.sc.hetero != 0, [&](auto I, auto H) {
    typedef KRhs<T, D, decltype(I)::value, RHS_ONLY, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.f = (const T*)f; a.t = (const T*)t; a.tcs = tcs; a.rhs = (T*)rhs; a.fin = no_fin();
    a.xd = (const T*)xd; a.xb = (const T*)xb; a.xcs = xcs;  // (t == nullptr: the explicit d, b of a warm start)
    return run<K, T, D>(a, o.g, o.tune, false, o.sc, stream);
  });
}

template <typename T, int D>
static int sweep_impl(const tvdk_geom_t* gg, const tvdk_coef_t* om, const tvdk_coef_t* el, const void* x_in,
                      const void* rhs, void* x_out, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(nullptr, om, el, D);
  return run_sweep<T, D>(k.sc.hetero != 0, o.g, o.tune, k, (const T*)x_in, (const T*)rhs, (T*)x_out, nullptr, no_fin(), o.sc, stream);
}

template <typename T, int D>
static int matvec_impl(const tvdk_geom_t* gg, const tvdk_coef_t* om, const tvdk_coef_t* el, const void* x, void* y,
                       void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(nullptr, om, el, D);
  return dispatch2(k.sc.hetero != 0, false, [&](auto H, auto) {
    typedef KCg<T, D, CG_MATVEC, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.p_in = (const T*)x; a.p_out = (T*)y; a.fin = no_fin();
    return run<K, T, D>(a, o.g, o.tune, false, o.sc, stream);
  });
}

template <typename T, int D>
static int shrink_impl(const tvdk_geom_t* gg, int iso, const tvdk_coef_t* mu, const tvdk_coef_t* el,
                       const void* u_new, const void* u_old, void* u_out, const void* t_in, void* t_out,
                       int64_t tcs, double* sum_out, void* ws, size_t wsb, void* stream, const void* xb = nullptr,
                       int64_t xcs = 0) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(mu, nullptr, el, D);
  return dispatch2(iso != 0, k.sc.hetero != 0, [&](auto I, auto H) {
    typedef KShrink<T, D, decltype(I)::value, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.unew[0] = (const T*)u_new; a.u_old = (const T*)u_old; a.u_out = (T*)u_out;
    a.t_in = (const T*)t_in; a.t_out = (T*)t_out; a.tcs = tcs; a.fin = sum_fin(sum_out);
    a.xb = (const T*)xb; a.xcs = xcs;  // (t_in == nullptr: the explicit b of a warm start)
    return run<K, T, D>(a, o.g, o.tune, false, o.sc, stream);
  });
}

template <typename T, int D>
static int expand_impl(const tvdk_geom_t* gg, int iso, const tvdk_coef_t* mu, const tvdk_coef_t* el, const void* t,
                       int64_t tcs, void* d, void* b, int64_t ocs, void* stream) {
  const Geom g = make_geom(*gg);
  const Tune tune = merge_tune(0, 0, 0);
  Scratch sc;
  memset(&sc, 0, sizeof sc);
  sc.partial_cap = (size_t)-1;
  const Coefs<T> k = make_coefs<T>(mu, nullptr, el, D);
  return dispatch2(iso != 0, false, [&](auto I, auto) {
    typedef KExpand<T, D, decltype(I)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.t = (const T*)t; a.tcs = tcs; a.d = (T*)d; a.b = (T*)b; a.ocs = ocs; a.fin = no_fin();
    return run<K, T, D>(a, g, tune, false, sc, stream);
  });
}

template <typename T, int D, int MODE>
static int dot_impl(const tvdk_geom_t* gg, const void* x, const void* y, double* out, void* ws, size_t wsb,
                    void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  typename KDot<T, D, MODE>::Args a;
  memset(&a, 0, sizeof a);
  a.x = (const T*)x; a.y = (const T*)y; a.fin = sum_fin(out);
  return run<KDot<T, D, MODE>, T, D>(a, o.g, o.tune, false, o.sc, stream);
}

template <typename T, int D>
static int energy_impl(const tvdk_geom_t* gg, const tvdk_coef_t* mu, const tvdk_coef_t* om, const void* x,
                       const void* f, double* out, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  typename KEnergy<T, D>::Args a;
  memset(&a, 0, sizeof a);
  a.k = make_coefs<T>(mu, om, nullptr, D);
  a.x = (const T*)x; a.f = (const T*)f; a.fin = sum_fin(out);
  return run<KEnergy<T, D>, T, D>(a, o.g, o.tune, false, o.sc, stream);
}

template <typename T, int D>
static int fj1_impl(const tvdk_geom_t* gg, int iso, const tvdk_coef_t* mu, const tvdk_coef_t* om,
                    const tvdk_coef_t* el, const void* f, const void* u_in, void* u_out, const void* t_in,
                    void* t_out, int64_t tcs, double* sum_out, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(mu, om, el, D);
  return run_fj1<T, D>(iso != 0, k.sc.hetero != 0, o.g, o.tune, k, (const T*)f, (const T*)u_in, (T*)u_out, (const T*)t_in, (T*)t_out,
                       tcs, tcs, nullptr, sum_fin(sum_out), o.sc, stream);
}

template <typename T, int D>
static int fj1x2_impl(const tvdk_geom_t* gg, int iso, const tvdk_coef_t* mu, const tvdk_coef_t* om,
                      const tvdk_coef_t* el, const void* f, const void* u_in, void* u_out, const void* t_in,
                      void* t_out, int64_t tcs, double* sum_out, void* ws, size_t wsb, void* stream) {
  if constexpr (D != 2) {
    return fail(TVDK_ERR_UNSUPPORTED, "the two-iteration fused kernel is 2-D only");
  } else {
    OpCtx<T, D> o;
    TVDK_TRY(o.init(gg, ws, wsb, stream));
    const Coefs<T> k = make_coefs<T>(mu, om, el, D);
    return dispatch2(iso != 0, k.sc.hetero != 0, [&](auto I, auto H) {
      typedef KFusedJ1x2<T, decltype(I)::value, decltype(H)::value> K;
      typename K::Args a;
      memset(&a, 0, sizeof a);
      a.k = k; a.f = (const T*)f; a.u_in = (const T*)u_in; a.u_out = (T*)u_out;
      a.t_in = (const T*)t_in; a.t_out = (T*)t_out; a.tcs_in = tcs; a.tcs_out = tcs; a.fin = sum_fin(sum_out);
      return run<K, T, D>(a, o.g, o.tune, true, o.sc, stream);
    });
  }
}

template <typename T, int D>
static int cginit_impl(const tvdk_geom_t* gg, int iso, const tvdk_coef_t* mu, const tvdk_coef_t* om,
                       const tvdk_coef_t* el, const void* f, const void* t, int64_t tcs, const void* u, void* x,
                       void* r, double* sum_out, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(mu, om, el, D);
  return dispatch2(iso != 0, k.sc.hetero != 0, [&](auto I, auto H) {
    typedef KRhs<T, D, decltype(I)::value, RHS_CGINIT, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.f = (const T*)f; a.t = (const T*)t; a.tcs = tcs; a.u = (const T*)u; a.x = (T*)x; a.r = (T*)r;
    a.fin = sum_fin(sum_out);
    return run<K, T, D>(a, o.g, o.tune, false, o.sc, stream);
  });
}

template <typename T, int D>
static int cgpq_impl(const tvdk_geom_t* gg, const tvdk_coef_t* om, const tvdk_coef_t* el, const void* p_in,
                     const void* r, void* p_out, const double* beta, int first, double* sum_out, void* ws,
                     size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(nullptr, om, el, D);
  DRef b;
  b.p = beta; b.stride = 1;
  IRef nof;
  nof.p = nullptr; nof.stride = 0;
  return run_cgpq<T, D>(k.sc.hetero != 0, o.g, o.tune, k, (const T*)p_in, (const T*)r, (T*)p_out, b, nof, first, nullptr, sum_fin(sum_out),
                        o.sc, stream);
}

template <typename T, int D>
static int cgupd_impl(const tvdk_geom_t* gg, const tvdk_coef_t* om, const tvdk_coef_t* el, const void* pp, void* x,
                      void* r, const double* alpha, double* sum_out, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(nullptr, om, el, D);
  return dispatch2(k.sc.hetero != 0, false, [&](auto H, auto) {
    typedef KCg<T, D, CG_UPDATE, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.p_in = (const T*)pp; a.x = (T*)x; a.r = (T*)r;
    a.coef.p = alpha; a.coef.stride = 1; a.fin = sum_fin(sum_out);
    return run<K, T, D>(a, o.g, o.tune, false, o.sc, stream);
  });
}

template <typename T, int D>
static int cgfirst_impl(const tvdk_geom_t* gg, const tvdk_coef_t* om, const tvdk_coef_t* el, const void* r,
                        double* sum_out, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(nullptr, om, el, D);
  return dispatch2(k.sc.hetero != 0, false, [&](auto H, auto) {
    typedef KCg<T, D, CG_FIRST, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.p_in = (const T*)r; a.fin = sum_fin(sum_out);
    return run<K, T, D>(a, o.g, o.tune, false, o.sc, stream);
  });
}

template <typename T, int D>
static int cgfused_impl(const tvdk_geom_t* gg, const tvdk_coef_t* om, const tvdk_coef_t* el, const void* p_in,
                        const void* r_in, void* p_out, void* r_out, void* x, const double* alpha,
                        const double* beta, double* sum_out, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(nullptr, om, el, D);
  return dispatch2(k.sc.hetero != 0, false, [&](auto H, auto) {
    typedef KCgFused<T, D, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.p_in = (const T*)p_in; a.r_in = (const T*)r_in; a.p_out = (T*)p_out; a.r_out = (T*)r_out;
    a.x = (T*)x; a.alpha.p = alpha; a.alpha.stride = 1; a.beta.p = beta; a.beta.stride = 1;
    a.fin = sum_fin(sum_out);
    return run<K, T, D>(a, o.g, o.tune, true, o.sc, stream);
  });
}

template <typename T, int D>
static int chambolle_impl(const tvdk_chambolle_t* c, void* stream, int* iters_launched) {
  const Geom g = make_geom(c->geom);
  const Tune tune = merge_tune(0, 0, 0);
  Scratch sc;
  TVDK_TRY((op_scratch<T, D>(&c->geom, c->workspace, c->workspace_bytes, &sc, stream)));
  tvdk_stat_t* st = c->stats;
  TVDK_TRY(backend::memset_async(st, 0, sizeof(tvdk_stat_t) * (size_t)g.nb, stream));
  TVDK_TRY(backend::memset_async(sc.ndone, 0, sizeof(int), stream));
  T* P[2] = {reinterpret_cast<T*>(c->p0), reinterpret_cast<T*>(c->p1)};
  const double tau = 1.0 / (2.0 * D);
  const int check_every = c->check_every > 0 ? c->check_every : 8;
  int launched = 0, slot = 0;
  bool pending[2] = {false, false}, all_done = false;
  for (int it = 0; it < c->max_num_iter && !all_done; ++it) {
    typedef KChamb<T, D, 0> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.f = reinterpret_cast<const T*>(c->f);
    a.p_in = P[it & 1]; a.p_out = P[(it & 1) ^ 1]; a.pcs = c->p_comp_stride;
    a.tau = (T)tau; a.tau_over_w = (T)(tau / c->weight);
    a.skip = st;
    a.fin = no_fin();
    a.fin.kind = FIN_CHAMB; a.fin.stats = st; a.fin.eps = c->eps; a.fin.history = c->history; a.fin.iter = it;
    a.fin.nb = g.nb; a.fin.ndone = sc.ndone; a.fin.buf_id = it & 1; a.fin.weight = c->weight;
    a.fin.inv_size = 1.0 / ((double)g.n0 * g.ny * g.nx);
    TVDK_TRY((run<K, T, D>(a, g, tune, true, sc, stream)));
    ++launched;
    if ((launched % check_every) == 0 && it + 1 < c->max_num_iter) {  // stop enqueueing once every image is done
      const int other = slot ^ 1;
      TVDK_TRY(backend::poll_begin(slot, sc.ndone, stream));
      pending[slot] = true;
      if (pending[other]) {
        int v = 0;
        TVDK_TRY(backend::poll_end(other, &v));
        pending[other] = false;
        if (v >= g.nb) all_done = true;
      }
      slot = other;
    }
  }
  for (int s2 = 0; s2 < 2; ++s2)
    if (pending[s2]) {
      int v = 0;
      TVDK_TRY(backend::poll_end(s2, &v));
    }
  {  // the result: image + div of the p the last executed iteration read (stats[b].cur_buf)
    typedef KChamb<T, D, 1> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.f = reinterpret_cast<const T*>(c->f);
    a.p_in = P[0]; a.p_in1 = P[1]; a.pcs = c->p_comp_stride; a.out = reinterpret_cast<T*>(c->out);
    a.skip = st; a.fin = no_fin();
    TVDK_TRY((run<K, T, D>(a, g, tune, true, sc, stream)));
  }
  if (iters_launched) *iters_launched = launched;
  return 0;
}

template <typename T, int D>
static int sweeps_impl(const tvdk_geom_t* gg, const tvdk_coef_t* om, const tvdk_coef_t* el, void* xa, const void* rhs,
                       void* xb, int n, int* result_in_b, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(nullptr, om, el, D);
  const bool het = k.sc.hetero != 0;
  T* src = (T*)xa;
  T* dst = (T*)xb;
  for (int rem = n; rem > 0;) {
    const int S = sweeps_per_pass(D, (long long)o.g.n0 * o.g.ny * o.g.nx, rem, het);
    if (S == 1) {
      TVDK_TRY(dispatch2(het, false, [&](auto H, auto) {
        typedef KSweep<T, D, decltype(H)::value> K;
        typename K::Args a;
        memset(&a, 0, sizeof a);
        a.k = k; a.x_in = src; a.rhs = (const T*)rhs; a.x_out = dst; a.fin = no_fin();
        return run<K, T, D>(a, o.g, o.tune, true, o.sc, stream);
      }));
    } else {
      auto go = [&](auto SS) {
        typedef KSweepN<T, D, decltype(SS)::value> K;
        typename K::Args a;
        memset(&a, 0, sizeof a);
        a.k = k; a.x_in = src; a.rhs = (const T*)rhs; a.x_out = dst; a.fin = no_fin();
        return run<K, T, D>(a, o.g, o.tune, true, o.sc, stream);
      };
      if (S == 4)
        TVDK_TRY(go(std::integral_constant<int, 4>{}));
      else
        TVDK_TRY(go(std::integral_constant<int, 2>{}));
    }
    std::swap(src, dst);
    rem -= S;
  }
  if (result_in_b) *result_in_b = (src == (T*)xb) ? 1 : 0;
  return 0;
}

template <typename T, int D>
static int resid_impl(const tvdk_geom_t* gg, const tvdk_coef_t* om, const tvdk_coef_t* el, const void* x,
                      const void* rhs, void* r, void* ws, size_t wsb, void* stream) {
  OpCtx<T, D> o;
  TVDK_TRY(o.init(gg, ws, wsb, stream));
  const Coefs<T> k = make_coefs<T>(nullptr, om, el, D);
  return dispatch2(k.sc.hetero != 0, false, [&](auto H, auto) {
    typedef KCg<T, D, CG_RESID, decltype(H)::value> K;
    typename K::Args a;
    memset(&a, 0, sizeof a);
    a.k = k; a.p_in = (const T*)x; a.r_in = (const T*)rhs; a.p_out = (T*)r; a.fin = no_fin();
    return run<K, T, D>(a, o.g, o.tune, false, o.sc, stream);
  });
}

template <typename T, int D, int MODE>
static int mgxfer_impl(const tvdk_geom_t* target, const tvdk_geom_t* other, const void* src, void* dst,
                       void* stream) {
  typedef KMgXfer<T, D, MODE> K;
  typename K::Args a;
  memset(&a, 0, sizeof a);
  a.g = make_geom(*target);
  a.go = make_geom(*other);
  a.src = (const T*)src; a.dst = (T*)dst; a.fin = no_fin();
  const int ntx = 128;
  const int gy = std::max(1, std::min((a.g.nx + ntx - 1) / ntx, 64));
  const long long rows = (long long)a.g.n0 * a.g.ny;
  ++g_launches;
  return backend::launch<K>(a, (int)rows, gy, a.g.nb, ntx, 1, 0, stream, nullptr, nullptr);
}

static int check_mg_pair(const tvdk_geom_t* f, const tvdk_geom_t* c) {
  TVDK_TRY(check_geom(f));
  TVDK_TRY(check_geom(c));
  if (f->dim != c->dim || f->dtype != c->dtype || f->nb != c->nb)
    return fail(TVDK_ERR_INVALID, "fine and coarse grids must agree in dim, dtype and nb");
  for (int a = 0; a < f->dim; ++a)
    if (c->n[a] != f->n[a] / 2 || c->n[a] < 1)
      return fail(TVDK_ERR_INVALID, "coarse extent n[%d] = %d must be fine / 2 = %d (>= 1)", a, c->n[a], f->n[a] / 2);
  return 0;
}

}  // namespace tvdk

#define TVDK_C_ENTRY(g, CALL)        \
  TVDK_TRY(check_geom(g));           \
  int rc = 0;                        \
  TVDK_DISPATCH(*g, rc = (CALL));    \
  return rc;

int tvdk_rhs(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* omega,
             const tvdk_coef_t* ell, const void* f, const void* t, int64_t t_comp_stride, void* rhs,
             void* workspace, size_t workspace_bytes, void* stream) {
  TVDK_C_ENTRY(g, (rhs_impl<T, D>(g, isotropic, mu, omega, ell, f, t, t_comp_stride, rhs, workspace,
                                  workspace_bytes, stream)))
}

int tvdk_rhs_explicit(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* f,
                      const void* d, const void* b, int64_t db_comp_stride, void* rhs, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (!d || !b) return fail(TVDK_ERR_INVALID, "tvdk_rhs_explicit needs both d and b");
  TVDK_C_ENTRY(g, (rhs_impl<T, D>(g, 0, nullptr, omega, ell, f, nullptr, 0, rhs, workspace, workspace_bytes, stream,
                                  d, b, db_comp_stride)))
}

int tvdk_jacobi_sweep(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* x_in,
                      const void* rhs, void* x_out, void* workspace, size_t workspace_bytes, void* stream) {
  TVDK_C_ENTRY(g, (sweep_impl<T, D>(g, omega, ell, x_in, rhs, x_out, workspace, workspace_bytes, stream)))
}

int tvdk_cg_matvec(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* x, void* y,
                   void* workspace, size_t workspace_bytes, void* stream) {
  TVDK_C_ENTRY(g, (matvec_impl<T, D>(g, omega, ell, x, y, workspace, workspace_bytes, stream)))
}

int tvdk_shrink(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* ell,
                const void* u_new, const void* u_old, void* u_out, const void* t_in, void* t_out,
                int64_t t_comp_stride, double* sum_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (u_out && u_out == u_new) return fail(TVDK_ERR_INVALID, "u_out must not alias u_new (it may alias u_old)");
  TVDK_C_ENTRY(g, (shrink_impl<T, D>(g, isotropic, mu, ell, u_new, u_old, u_out, t_in, t_out, t_comp_stride,
                                     sum_out, workspace, workspace_bytes, stream)))
}

int tvdk_shrink_explicit(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* ell,
                         const void* u_new, const void* u_old, void* u_out, const void* b_in, int64_t b_comp_stride,
                         void* t_out, int64_t t_comp_stride, double* sum_out, void* workspace, size_t workspace_bytes,
                         void* stream) {
  if (!b_in) return fail(TVDK_ERR_INVALID, "tvdk_shrink_explicit needs b_in");
  if (u_out && u_out == u_new) return fail(TVDK_ERR_INVALID, "u_out must not alias u_new (it may alias u_old)");
  TVDK_C_ENTRY(g, (shrink_impl<T, D>(g, isotropic, mu, ell, u_new, u_old, u_out, nullptr, t_out, t_comp_stride,
                                     sum_out, workspace, workspace_bytes, stream, b_in, b_comp_stride)))
}

int tvdk_expand_state(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* ell,
                      const void* t, int64_t t_comp_stride, void* d, void* b, int64_t db_comp_stride,
                      void* stream) {
  TVDK_C_ENTRY(g, (expand_impl<T, D>(g, isotropic, mu, ell, t, t_comp_stride, d, b, db_comp_stride, stream)))
}

int tvdk_dot(const tvdk_geom_t* g, const void* x, const void* y, double* out, void* workspace,
             size_t workspace_bytes, void* stream) {
  TVDK_C_ENTRY(g, (dot_impl<T, D, 0>(g, x, y, out, workspace, workspace_bytes, stream)))
}

int tvdk_norm_diff2(const tvdk_geom_t* g, const void* x, const void* y, double* out, void* workspace,
                    size_t workspace_bytes, void* stream) {
  TVDK_C_ENTRY(g, (dot_impl<T, D, 1>(g, x, y, out, workspace, workspace_bytes, stream)))
}

int tvdk_energy(const tvdk_geom_t* g, const tvdk_coef_t* mu, const tvdk_coef_t* omega, const void* x,
                const void* f, double* out, void* workspace, size_t workspace_bytes, void* stream) {
  TVDK_C_ENTRY(g, (energy_impl<T, D>(g, mu, omega, x, f, out, workspace, workspace_bytes, stream)))
}

int tvdk_fused_jacobi1(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* omega,
                       const tvdk_coef_t* ell, const void* f, const void* u_in, void* u_out, const void* t_in,
                       void* t_out, int64_t t_comp_stride, double* sum_out, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (u_in == u_out || t_in == t_out) return fail(TVDK_ERR_INVALID, "fused kernel is out-of-place");
  TVDK_C_ENTRY(g, (fj1_impl<T, D>(g, isotropic, mu, omega, ell, f, u_in, u_out, t_in, t_out, t_comp_stride,
                                  sum_out, workspace, workspace_bytes, stream)))
}

int tvdk_cg_init(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* omega,
                 const tvdk_coef_t* ell, const void* f, const void* t, int64_t t_comp_stride, const void* u, void* x,
                 void* r, double* sum_out, void* workspace, size_t workspace_bytes, void* stream) {
  TVDK_C_ENTRY(g, (cginit_impl<T, D>(g, isotropic, mu, omega, ell, f, t, t_comp_stride, u, x, r, sum_out,
                                     workspace, workspace_bytes, stream)))
}

int tvdk_cg_pq(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* p_in,
               const void* r, void* p_out, const double* beta, int first, double* sum_out, void* workspace,
               size_t workspace_bytes, void* stream) {
  if (p_in == p_out) return fail(TVDK_ERR_INVALID, "cg_pq is out-of-place in p");
  TVDK_C_ENTRY(g, (cgpq_impl<T, D>(g, omega, ell, p_in, r, p_out, beta, first, sum_out, workspace,
                                   workspace_bytes, stream)))
}

int tvdk_cg_update(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* p, void* x,
                   void* r, const double* alpha, double* sum_out, void* workspace, size_t workspace_bytes,
                   void* stream) {
  TVDK_C_ENTRY(g, (cgupd_impl<T, D>(g, omega, ell, p, x, r, alpha, sum_out, workspace, workspace_bytes, stream)))
}

int tvdk_cg_first(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* r,
                  double* sum_out, void* workspace, size_t workspace_bytes, void* stream) {
  TVDK_C_ENTRY(g, (cgfirst_impl<T, D>(g, omega, ell, r, sum_out, workspace, workspace_bytes, stream)))
}

int tvdk_cg_fused(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* p_in,
                  const void* r_in, void* p_out, void* r_out, void* x, const double* alpha, const double* beta,
                  double* sum_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (p_out == p_in || p_out == r_in || r_out == r_in || r_out == p_in)
    return fail(TVDK_ERR_INVALID, "cg_fused is out-of-place in p and r");
  TVDK_C_ENTRY(g, (cgfused_impl<T, D>(g, omega, ell, p_in, r_in, p_out, r_out, x, alpha, beta, sum_out, workspace,
                                      workspace_bytes, stream)))
}

int tvdk_residual(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, const void* x,
                  const void* rhs, void* r, void* workspace, size_t workspace_bytes, void* stream) {
  if (r == x) return fail(TVDK_ERR_INVALID, "residual is out-of-place in x");
  TVDK_C_ENTRY(g, (resid_impl<T, D>(g, omega, ell, x, rhs, r, workspace, workspace_bytes, stream)))
}

int tvdk_mg_restrict(const tvdk_geom_t* fine, const tvdk_geom_t* coarse, const void* x_fine, void* x_coarse,
                     void* stream) {
  TVDK_TRY(check_mg_pair(fine, coarse));
  if (!x_fine || !x_coarse) return fail(TVDK_ERR_INVALID, "NULL pointer");
  int rc = 0;
  TVDK_DISPATCH(*fine, rc = (mgxfer_impl<T, D, 0>(coarse, fine, x_fine, x_coarse, stream)));
  return rc;
}

int tvdk_mg_prolong_add(const tvdk_geom_t* fine, const tvdk_geom_t* coarse, const void* e_coarse, void* x_fine,
                        void* stream) {
  TVDK_TRY(check_mg_pair(fine, coarse));
  if (!e_coarse || !x_fine) return fail(TVDK_ERR_INVALID, "NULL pointer");
  int rc = 0;
  TVDK_DISPATCH(*fine, rc = (mgxfer_impl<T, D, 1>(fine, coarse, e_coarse, x_fine, stream)));
  return rc;
}

int tvdk_jacobi_sweeps(const tvdk_geom_t* g, const tvdk_coef_t* omega, const tvdk_coef_t* ell, void* xa,
                       const void* rhs, void* xb, int nsweeps, int* result_in_b, void* workspace,
                       size_t workspace_bytes, void* stream) {
  if (!xa || !xb || !rhs || xa == xb) return fail(TVDK_ERR_INVALID, "jacobi_sweeps needs two distinct iterate buffers");
  if (nsweeps < 0) return fail(TVDK_ERR_INVALID, "nsweeps must be >= 0");
  if (g && (g->z_begin != 0 || g->z_end != g->n[0] || g->g0_off != 0 || g->g0_n != g->n[0]))
    return fail(TVDK_ERR_INVALID, "jacobi_sweeps works on whole volumes (slabs sweep one at a time)");
  TVDK_C_ENTRY(g, (sweeps_impl<T, D>(g, omega, ell, xa, rhs, xb, nsweeps, result_in_b, workspace, workspace_bytes,
                                     stream)))
}

int tvdk_fused_jacobi1_x2(const tvdk_geom_t* g, int isotropic, const tvdk_coef_t* mu, const tvdk_coef_t* omega,
                          const tvdk_coef_t* ell, const void* f, const void* u_in, void* u_out, const void* t_in,
                          void* t_out, int64_t t_comp_stride, double* sum_out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (u_in == u_out || t_in == t_out) return fail(TVDK_ERR_INVALID, "fused kernel is out-of-place");
  if (g && (g->z_begin != 0 || g->z_end != g->n[0] || g->g0_off != 0 || g->g0_n != g->n[0]))
    return fail(TVDK_ERR_INVALID, "fused_jacobi1_x2 works on whole images");
  TVDK_C_ENTRY(g, (fj1x2_impl<T, D>(g, isotropic, mu, omega, ell, f, u_in, u_out, t_in, t_out, t_comp_stride,
                                    sum_out, workspace, workspace_bytes, stream)))
}

int tvdk_chambolle(const tvdk_chambolle_t* c, void* stream, int* iters_launched) {
  if (!c) return fail(TVDK_ERR_INVALID, "problem is NULL");
  TVDK_TRY(check_geom(&c->geom));
  if (c->max_num_iter < 1) return fail(TVDK_ERR_INVALID, "max_num_iter must be >= 1");
  if (!(c->weight > 0.0)) return fail(TVDK_ERR_INVALID, "weight must be positive");
  if (!c->f || !c->out || !c->p0 || !c->p1 || !c->stats) return fail(TVDK_ERR_INVALID, "NULL pointer");
  if (c->geom.z_begin != 0 || c->geom.z_end != c->geom.n[0] || c->geom.g0_off != 0 || c->geom.g0_n != c->geom.n[0])
    return fail(TVDK_ERR_INVALID, "tvdk_chambolle works on whole images");
  int rc = 0;
  TVDK_RANGE("tvdk_chambolle");
  TVDK_DISPATCH(c->geom, rc = (chambolle_impl<T, D>(c, stream, iters_launched)));
  return rc;
}
