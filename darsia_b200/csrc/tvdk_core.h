// tvdk_core.h -- the kernels of the split-Bregman TVD hot path, written once as per-thread
// "functor" state machines.
//
// Every kernel is a struct K with
//     bool init(const Args&, const LaunchCtx&, unsigned char* smem)   block-uniform early exit
//     bool more() const                                               block-uniform loop test
//     void pre (const Args&, unsigned char* smem)                     phase before the barrier
//     void post(const Args&, unsigned char* smem)                     phase after the barrier
//     double red[NRED]                                                per-thread partial sums
//     static void finalize(const Args&, int b, const double* totals)  run once per problem
// The sm_100a build (tvdk.cu) wraps K in a __global__ runner with one __syncthreads() between
// pre and post; tests/emu builds the SAME code with g++ and steps the threads of a block
// phase by phase on the CPU (test infrastructure only -- the product never loads it).
//
// Common scheme ("march"): a thread owns one 128-bit vector of the contiguous axis (2 fp64 /
// 4 fp32 values) at a fixed in-plane position and marches along axis 0, keeping the axis-0
// neighbours in registers; in-plane neighbours are exchanged through double-buffered shared
// memory rows (one barrier per plane).  Thread blocks overlap by a halo of threads that only
// feed values (and recompute u+ where the fused kernel needs it), so every global access of
// the interior is an aligned, coalesced 128-bit access and each input plane is read once per
// chunk.  Boundary rules follow SURVEY.md Appendix A (reference lines cited per kernel).
#pragma once

#include <stdint.h>

#include "../../include/tvdk.h"

#if defined(__CUDACC__)
#define TVDK_DEV __device__ __forceinline__
#define TVDK_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define TVDK_DEV inline
#define TVDK_HD inline
#endif

namespace tvdk {

struct LaunchCtx {
  int tx, ty, ntx, nty;  // thread coordinates / block shape
  int bx, by, bz;        // block coordinates
  int nbx, nby;          // grid shape (x, y)
};

// ------------------------------------------------------------------------------------------
// 128-bit vectors
// ------------------------------------------------------------------------------------------
template <typename T>
struct alignas(16) Vec {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

template <typename T>
TVDK_DEV Vec<T> vsplat(T x) {
  Vec<T> r;
#pragma unroll
  for (int i = 0; i < Vec<T>::N; ++i) r.v[i] = x;
  return r;
}

#if defined(__CUDACC__)
// read-only 128-bit load (LDG.E.128 through the non-coherent path)
TVDK_DEV Vec<double> vload_ro(const double* p) {
  double2 d = __ldg(reinterpret_cast<const double2*>(p));
  Vec<double> r;
  r.v[0] = d.x;
  r.v[1] = d.y;
  return r;
}
TVDK_DEV Vec<float> vload_ro(const float* p) {
  float4 d = __ldg(reinterpret_cast<const float4*>(p));
  Vec<float> r;
  r.v[0] = d.x;
  r.v[1] = d.y;
  r.v[2] = d.z;
  r.v[3] = d.w;
  return r;
}
template <typename T>
TVDK_DEV T sload_ro(const T* p) {
  return __ldg(p);
}
#else
template <typename T>
TVDK_DEV Vec<T> vload_ro(const T* p) {
  return *reinterpret_cast<const Vec<T>*>(p);
}
template <typename T>
TVDK_DEV T sload_ro(const T* p) {
  return *p;
}
#endif
template <typename T>
TVDK_DEV Vec<T> vload(const T* p) {
  return *reinterpret_cast<const Vec<T>*>(p);
}
template <typename T>
TVDK_DEV void vstore(T* p, const Vec<T>& x) {
  *reinterpret_cast<Vec<T>*>(p) = x;
}

#if defined(__CUDACC__)
// Branch-free fp64 reciprocal / division / square root: hardware seed (MUFU.RCP64H / RSQ64H,
// >= 20 good bits) + Newton steps in DFMA.  Operands on this path are positive and far from the
// subnormal / overflow range (s + 1e-18, omega + 2*dim*ell, ell), so the IEEE routines' slow
// paths -- a dozen extra instructions and two branches per call -- are dead weight here.
// Results agree with correctly rounded sqrt / division to <= 1 ulp.
TVDK_DEV double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
TVDK_DEV double tdiv(double a, double x) {
  const double r = fast_rcp(x);
  const double q = a * r;
  return fma(fma(-x, q, a), r, q);
}
TVDK_DEV double tsqrt(double q) {  // q >= 0
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
  const double h = 0.5 * q;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  double s = q * y;
  s = fma(fma(-s, s, q), 0.5 * y, s);
  return q > 0.0 ? s : 0.0;
}
TVDK_DEV float tdiv(float a, float x) { return __fdividef(a, x); }
TVDK_DEV float tsqrt(float x) { return sqrtf(x); }
TVDK_DEV double tcopysign(double m, double s) { return copysign(m, s); }
TVDK_DEV float tcopysign(float m, float s) { return copysignf(m, s); }
#else
TVDK_DEV double tdiv(double a, double x) { return a / x; }
TVDK_DEV float tdiv(float a, float x) { return a / x; }
TVDK_DEV double tsqrt(double x) { return sqrt(x); }
TVDK_DEV float tsqrt(float x) { return sqrtf(x); }
TVDK_DEV double tcopysign(double m, double s) { return std::copysign(m, s); }
TVDK_DEV float tcopysign(float m, float s) { return std::copysign(m, s); }
#endif
TVDK_DEV double tabs(double x) { return fabs(x); }
TVDK_DEV float tabs(float x) { return fabsf(x); }
TVDK_DEV double tmax(double a, double b) { return a > b ? a : b; }  // NaN-free inputs assumed
TVDK_DEV float tmax(float a, float b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------
// Geometry and coefficients (device-side mirrors of tvdk_geom_t / tvdk_coef_t)
// ------------------------------------------------------------------------------------------
struct Geom {
  int n0, ny, nx;      // local extents; 2-D: ny == 1
  int pitch;           // elements per row
  long long s0;        // stride of axis 0 (elements) = ny * pitch
  long long sb;        // batch stride (elements) = n0 * s0
  int nb;
  int g0_off, g0_n;    // slab decomposition along axis 0 (global plane = local + g0_off)
  int z_begin, z_end;  // owned planes (local indices)
  int zchunk;          // planes per CTA
  int xtiles, ytiles;  // in-plane tiling (depends on the kernel's halo widths)
  int pf;              // L2 prefetch distance in planes (0 = off)
  int pfsub;           // 1: one thread per 128-byte line issues the prefetch, 0: every thread
  int rev;             // 1: thread blocks take the axis-0 chunks in DESCENDING order (see tvdk_host.inl: run)
};

template <typename T>
struct Coef {
  const T* p;  // nullptr -> scalar
  T val;
  long long sb;
};

// All scalars are computed on the host in double precision exactly like the reference does
// with Python floats, then cast to T (NumPy treats Python scalars as weak-typed).
template <typename T>
struct Scalars {
  T mu, omega, ell;
  T kappa;     // mu / ell                      (split_bregman_tvd.py:140,145)
  T inv_diag;  // 1 / (omega + ell*2*dim)       (jacobi.py:57-72, :107)
  T c_nb;      // 1 / (diag / ell)              (jacobi.py:103-106, :112-114)
  T dim_t;     // dim as T
  int hetero;  // any per-voxel coefficient
};

template <typename T>
struct Coefs {
  Coef<T> mu, omega, ell;
  Scalars<T> sc;
};

// reference to a per-problem double scalar living either in tvdk_stat_t or in a plain array
struct DRef {
  const double* p;
  int stride;  // in doubles
  TVDK_DEV double at(int b) const { return p[(long long)b * stride]; }
};
struct IRef {
  const int* p;  // nullptr -> never set
  int stride;    // in ints
  TVDK_DEV int at(int b) const { return p ? p[(long long)b * stride] : 0; }
};

// ------------------------------------------------------------------------------------------
// Finalisation of per-problem reductions (runs in the last block of each problem)
// ------------------------------------------------------------------------------------------
enum FinKind {
  FIN_NONE = 0,
  FIN_NRMF,    // totals[0] = sum f^2
  FIN_OUTER,   // totals[0] = sum (u+ - u)^2
  FIN_OUTER2,  // totals = {sum (u' - u)^2, sum (u'' - u')^2}: two outer iterations in one launch (K1x2)
  FIN_CGINIT,  // totals = {sum rhs^2, sum r^2}
  FIN_CGPQ,    // totals[0] = p.q
  FIN_CGRR,    // totals[0] = r.r
  FIN_CGFIRST, // totals = {r.q, q.q}, q = A r_0          (single-pass CG route, K7)
  FIN_CGFUSED, // totals = {r.r, p.q, r.q, q.q}
  FIN_JTOL,    // totals = {sum (x_new - x)^2, sum x_in^2}
  FIN_ENERGY,  // totals = {sum (omega (x - f))^2, sum |mu G x|}
  FIN_CHAMB    // totals = {sum d^2, sum |grad out|}: Chambolle energy and stopping rule
};

struct Fin {
  int kind;
  tvdk_stat_t* stats;  // device, may be nullptr when kind == FIN_NONE
  double* sum_out;     // optional raw totals, [b * NRED + r]
  double eps;          // FIN_OUTER; < 0 = None
  double tol;          // CG / Jacobi tolerance
  double* history;     // optional, [iter * nb + b]
  int iter, nb;
  int* ndone;          // number of converged problems (host polls it)
  double weight;       // FIN_CHAMB: TV weight
  double inv_size;     // FIN_CHAMB: 1 / number of voxels
  int buf_id;          // FIN_OUTER: buffer now holding u ; FIN_JTOL: buffer the sweep wrote ; FIN_CHAMB: buffer read
  int first_sweep;     // FIN_JTOL: this sweep started from x0
};

TVDK_DEV void atomic_inc(int* p) {
#if defined(__CUDACC__)
  atomicAdd(p, 1);
#else
  *p += 1;
#endif
}

TVDK_DEV void finalize_fin(const Fin& f, int nred, int b, const double* tot) {
  if (f.sum_out) {
    for (int r = 0; r < nred; ++r) f.sum_out[b * nred + r] = tot[r];
  }
  if (f.kind == FIN_NONE || !f.stats) return;
  tvdk_stat_t& s = f.stats[b];
  switch (f.kind) {
    case FIN_NRMF:  // split_bregman_tvd.py:58
      s.nrm_f = sqrt(tot[0]);
      break;
    case FIN_OUTER: {  // split_bregman_tvd.py:161, :173-175
      double inc = sqrt(tot[0]) / s.nrm_f;
      s.inc = inc;
      s.iters += 1;
      s.cur_buf = f.buf_id;
      s.jac_stop = 0;
      s.jac_sel = 0;
      s.cg_stop = 0;
      if (f.history) f.history[(long long)f.iter * f.nb + b] = inc;
      const bool bad = !(inc <= 1.7976931348623157e308);  // NaN or +Inf
      if (bad) s.nonfinite = 1;
      if ((f.eps >= 0.0 && inc < f.eps) || bad) {
        s.done = 1;
        if (f.ndone) atomic_inc(f.ndone);
      }
    } break;
    case FIN_OUTER2: {  // two iterations of :161 (no eps: the host only fuses pairs when eps is None)
      const double i1 = sqrt(tot[0]) / s.nrm_f, i2 = sqrt(tot[1]) / s.nrm_f;
      if (!(i1 <= 1.7976931348623157e308) || !(i2 <= 1.7976931348623157e308)) {
        s.nonfinite = 1;
        s.done = 1;
        if (f.ndone) atomic_inc(f.ndone);
      }
      s.inc = i2;
      s.iters += 2;
      s.cur_buf = f.buf_id;
      if (f.history) {
        f.history[(long long)f.iter * f.nb + b] = i1;
        f.history[(long long)(f.iter + 1) * f.nb + b] = i2;
      }
    } break;
    case FIN_CGINIT: {  // scipy cg: bnrm2, atol, early return, r = b - A x
      double bn = sqrt(tot[0]);
      s.bnrm = bn;
      double at = f.tol * bn;
      s.atol = at > 0.0 ? at : 0.0;
      s.rho = tot[1];
      s.rho_prev = 0.0;
      s.cg_first = 1;
      s.cg_iters = 0;
      s.cg_stop = (bn == 0.0) || (sqrt(tot[1]) < s.atol) ? 1 : 0;
    } break;
    case FIN_CGPQ:  // alpha = rho / (p.q)
      s.pq = tot[0];
      s.alpha = s.rho / tot[0];
      break;
    case FIN_CGRR:  // rho_prev = rho; rho = r.r; beta; "if ||r|| < atol" of the next spin
      s.rho_prev = s.rho;
      s.rho = tot[0];
      s.beta = s.rho / s.rho_prev;
      s.cg_first = 0;
      s.cg_iters += 1;
      s.cg_stop = (sqrt(tot[0]) < s.atol) ? 1 : 0;
      break;
    case FIN_CGFIRST: {  // alpha_0 = rho_0 / (p_0.q_0), p_0 = r_0; beta_1 from the predicted rho_1
      const double rho = s.rho;
      const double al = rho / tot[0];
      const double rho_next = fma(al, fma(al, tot[1], -2.0 * tot[0]), rho);  // ||r - alpha q||^2
      s.pq = tot[0];
      s.alpha = al;
      s.rho_prev = rho;
      s.beta = rho_next / rho;
      s.rho_pred = rho_next;
      s.cg_first = 0;
      s.cg_iters = 1;
    } break;
    case FIN_CGFUSED: {  // rho_i summed directly; alpha_i = rho_i / (p_i.q_i); beta_{i+1} predicted
      const double rho = tot[0];
      const double al = rho / tot[1];
      const double rho_next = fma(al, fma(al, tot[3], -2.0 * tot[2]), rho);
      const double dev = fabs(s.rho_pred - rho) / rho;  // how well the previous launch predicted this rho
      if (dev > s.cg_pred_dev) s.cg_pred_dev = dev;
      s.rho_pred = rho_next;
      s.rho_prev = s.rho;
      s.rho = rho;
      s.pq = tot[1];
      s.alpha = al;
      s.beta = rho_next / rho;
      s.cg_iters += 1;
    } break;
    case FIN_JTOL: {  // jacobi.py:118-131 (old iterate kept on convergence, Q7)
      if (f.first_sweep) s.jac_x0nrm2 = tot[1];
      double err = sqrt(tot[0]) / sqrt(s.jac_x0nrm2);
      s.jac_err = err;
      if (err < f.tol)
        s.jac_stop = 1;
      else
        s.jac_sel = f.buf_id;
    } break;
    case FIN_CHAMB: {  // skimage _denoise_tv_chambolle_nd: E, E_init, E_previous, break test
      const double e = (tot[0] + f.weight * tot[1]) * f.inv_size;
      s.energy = e;
      s.iters += 1;
      s.cur_buf = f.buf_id;  // the p this iteration READ: the result is image + div of it
      if (f.history) f.history[(long long)f.iter * f.nb + b] = e;
      if (f.iter == 0) {
        s.e_init = e;
        s.e_prev = e;
      } else if (fabs(s.e_prev - e) < f.eps * s.e_init) {
        s.done = 1;
        if (f.ndone) atomic_inc(f.ndone);
      } else {
        s.e_prev = e;
      }
    } break;
    case FIN_ENERGY:  // split_bregman_tvd.py:71-77
      s.energy = 0.5 * tot[0] + tot[1];
      if (f.history) f.history[(long long)f.iter * f.nb + b] = s.energy;
      break;
    default:
      break;
  }
}

// ------------------------------------------------------------------------------------------
// Thread <-> column mapping of the march kernels
// ------------------------------------------------------------------------------------------
// Every thread loads UNCONDITIONALLY from a position clamped into the allocation (no branches
// around loads); what a thread outside the image loads is never consumed because the boundary
// rules select the other operand.  Blocks whose whole region (interior + halo) lies strictly
// inside the image take a fast path without any per-element boundary selects (`edge == false`).
// LH / RH: halo rows below / above along y (3-D); LHX / RHX: halo VECTORS left / right along x (a
// vector already holds >= 2 elements, so one vector covers an element halo of up to 2).
template <typename T, int D, int LH, int RH, int LHX_ = -1, int RHX_ = -1>
struct Col {
  static constexpr int V = Vec<T>::N;
  static constexpr int RHX = RHX_ < 0 ? (RH > 0 ? 1 : 0) : RHX_;
  static constexpr int LHX = LHX_ < 0 ? LH : LHX_;
  int X, Y, b;      // global coordinates of the first element of the thread's vector; problem
  int soff;         // shared-memory element offset of the vector inside one row array
  int spitch;       // shared-memory row pitch (elements)
  long long base;   // BYTE offset of (b, plane 0, Yc, Xc), clamped into the allocation
  long long cur;    // running BYTE offset of the current plane j
  long long cdel;   // BYTES: -b * sb (to rebase onto a coefficient map's own batch stride)
  long long s0b;    // BYTES between consecutive planes
  long long pfb;    // BYTES of the L2 prefetch distance (pf planes)
  int pfz;          // last plane index from which a prefetch still lands inside the allocation
  int z0, z1;       // planes [z0, z1) owned by this CTA
  bool st;          // interior thread whose results are stored
  bool edge;        // block touches the image boundary / is a partial tile (block-uniform)
  bool okc[V];      // element inside the image
  bool hl[V], hr[V];  // element has a left / right neighbour along x
  bool hylo, hyhi;    // row has a lower / upper neighbour along y (3-D)

  static TVDK_HD int iw(int ntx) { return (ntx - LHX - RHX) * V; }
  static TVDK_HD int ih(int nty) { return D == 3 ? nty - LH - RH : 1; }
  // (2-D: a single row with one vector of padding left and right -- no y neighbours to serve)
  static TVDK_HD int row_elems(int ntx, int nty) { return D == 3 ? (nty + 2) * (ntx + 2) * V : (ntx + 2) * V; }

  TVDK_DEV void setup(const Geom& g, const LaunchCtx& lc) {
    const int xt = lc.bx % g.xtiles, yt = lc.bx / g.xtiles;
    const int x0 = xt * iw(lc.ntx);
    X = x0 + (lc.tx - LHX) * V;
    bool rowok = true;
    Y = 0;
    bool edge_y = false;
    if (D == 3) {
      const int y0 = yt * ih(lc.nty);
      Y = y0 + (lc.ty - LH);
      rowok = (Y >= 0) && (Y < g.ny);
      edge_y = (y0 - LH < 0) || (y0 + ih(lc.nty) + RH > g.ny);
    }
    edge = edge_y || (x0 - LHX * V < 0) || (x0 + iw(lc.ntx) + RHX * V > g.nx);
    b = lc.bz;
    z0 = g.z_begin + lc.by * g.zchunk;
    z1 = z0 + g.zchunk < g.z_end ? z0 + g.zchunk : g.z_end;
    const bool inalloc = rowok && X >= 0 && X < g.pitch;
    const bool xin = lc.tx >= LHX && lc.tx < lc.ntx - RHX;
    const bool yin = (D == 3) ? (lc.ty >= LH && lc.ty < lc.nty - RH) : true;
    st = inalloc && xin && yin && X < g.nx;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      okc[v] = rowok && (X + v >= 0) && (X + v < g.nx);
      hl[v] = X + v > 0;
      hr[v] = X + v < g.nx - 1;
    }
    hylo = Y > 0;
    hyhi = Y < g.ny - 1;
    spitch = (lc.ntx + 2) * V;
    soff = (D == 3 ? (lc.ty + 1) * spitch : 0) + (lc.tx + 1) * V;
    const int Xc = X < 0 ? 0 : (X > g.pitch - V ? g.pitch - V : X);
    const int Yc = Y < 0 ? 0 : (Y > g.ny - 1 ? g.ny - 1 : Y);
    constexpr long long W = (long long)sizeof(T);
    cdel = -(long long)b * g.sb * W;
    s0b = g.s0 * W;
    base = ((long long)b * g.sb + (long long)Yc * g.pitch + Xc) * W;
    cur = base + (long long)z0 * s0b;
    pfb = (long long)g.pf * s0b;
    pfz = (g.pf > 0 && (!g.pfsub || (lc.tx & 7) == 0)) ? g.n0 - g.pf : -1;  // prefetch from plane j iff j < pfz
#if defined(__CUDACC__)
    // keep the per-thread invariants in registers instead of re-deriving them from %tid / the
    // kernel parameters on every plane (the compiler otherwise rematerialises them)
    asm volatile("" : "+r"(soff), "+l"(cur), "+l"(s0b));
#endif
  }
  TVDK_DEV long long at(const Geom&, int z) const { return base + (long long)z * s0b; }
  TVDK_DEV bool pf_ok(int z) const { return z < pfz; }
  // plane z exists in the local allocation and in the global domain
  TVDK_DEV bool plane_ok(const Geom& g, int z) const {
    const int gz = z + g.g0_off;
    return z >= 0 && z < g.n0 && gz >= 0 && gz < g.g0_n;
  }
  TVDK_DEV bool has_zlo(const Geom& g, int z) const { return z + g.g0_off > 0; }
  TVDK_DEV bool has_zhi(const Geom& g, int z) const { return z + g.g0_off < g.g0_n - 1; }
  // boundary predicates; compile-time true on the interior fast path
  template <bool E> TVDK_DEV bool HL(int v) const { return E ? hl[v] : true; }
  template <bool E> TVDK_DEV bool HR(int v) const { return E ? hr[v] : true; }
  template <bool E> TVDK_DEV bool OK(int v) const { return E ? okc[v] : true; }
  template <bool E> TVDK_DEV bool YLO() const { return E ? hylo : true; }
  template <bool E> TVDK_DEV bool YHI() const { return E ? hyhi : true; }
  template <bool E> TVDK_DEV bool ZLO(const Geom& g, int z) const { return E ? has_zlo(g, z) : true; }
  template <bool E> TVDK_DEV bool ZHI(const Geom& g, int z) const { return E ? has_zhi(g, z) : true; }
  // plane z touches the global z boundary (block-uniform)
  TVDK_DEV bool zedge(const Geom& g, int z) const { return !has_zlo(g, z) || !has_zhi(g, z); }
};

#if defined(__CUDACC__)
// L2 prefetch hint (no register is tied up, no scoreboard): pulls a line that this CTA will load
// a few planes later from HBM into L2, so that the march's one-plane-ahead register prefetch
// sees L2 latency instead of DRAM latency.
TVDK_DEV void pf_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
TVDK_DEV void pf_l2(const void*) {}
#endif

// all offsets `o` below are BYTE offsets
template <typename T>
TVDK_DEV const T* bptr(const T* base, long long o) {
  return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + o);
}
template <typename T>
TVDK_DEV T* bptr(T* base, long long o) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + o);
}
template <typename T>
TVDK_DEV Vec<T> ldo(const T* base, long long o) {
  return vload_ro(bptr(base, o));
}
template <typename T>
TVDK_DEV void pfo(const T* base, long long o) {
  pf_l2(bptr(base, o));
}
template <typename T>
TVDK_DEV Vec<T> ldw(const T* base, long long o) {  // plain load (data updated in place by this kernel)
  return vload(bptr(base, o));
}
template <typename T>
TVDK_DEV void stw(T* base, long long o, const Vec<T>& x) {
  vstore(bptr(base, o), x);
}
template <bool HET, typename T, typename C>
TVDK_DEV Vec<T> ldcoef(const C& c, const Coef<T>& k, long long o) {
  if (HET && k.p) return vload_ro(bptr(k.p + (long long)c.b * k.sb, o + c.cdel));
  return vsplat<T>(k.val);
}
template <typename T, typename C>
TVDK_DEV void sto(const C& c, T* base, long long o, const Vec<T>& x) {
  if (c.st) vstore(bptr(base, o), x);
}

// shared-memory row arrays: [nty + 2][(ntx + 2) * V], one vector of padding on every side
template <typename T>
TVDK_DEV T* srow(unsigned char* smem, int idx, int row_elems) {
  return reinterpret_cast<T*>(smem) + (long long)idx * row_elems;
}
template <typename T, typename C>
TVDK_DEV void sst(T* row, const C& c, const Vec<T>& x) {
  vstore(row + c.soff, x);
}
template <typename T, typename C>
TVDK_DEV T sxm(const T* row, const C& c) {
  return row[c.soff - 1];
}
template <typename T, typename C>
TVDK_DEV T sxp(const T* row, const C& c) {
  return row[c.soff + C::V];
}
template <typename T, typename C>
TVDK_DEV Vec<T> sym(const T* row, const C& c) {
  return vload(row + c.soff - c.spitch);
}
template <typename T, typename C>
TVDK_DEV Vec<T> syp(const T* row, const C& c) {
  return vload(row + c.soff + c.spitch);
}

// ------------------------------------------------------------------------------------------
// Pointwise pieces
// ------------------------------------------------------------------------------------------
// From t = grad(u) + b_old at one voxel: b (new Bregman variable) and w = ell * (b - d).
//   anisotropic: d = sign(t) max(|t| - kappa, 0); b = t - d   (split_bregman_tvd.py:80-81,142-146)
//   isotropic:   s = sqrt(sum t^2); phi = max(s - kappa, 0)/(s + 1e-18); d = t phi; b = t (1 - phi)
//                                                           (split_bregman_tvd.py:116-126)
// the isotropic shrinkage factor phi = max(s - kappa, 0) / (s + 1e-18), s = |t|_2   (split_bregman_tvd.py:117-126)
template <typename T, int D, bool ISO>
TVDK_DEV T iso_phi_of(const T (&t)[D], T kappa) {
  T s = t[0] * t[0];
#pragma unroll
  for (int a = 1; a < D; ++a) s = s + t[a] * t[a];
  s = tsqrt(s);
  return tdiv(tmax(s - kappa, (T)0), s + (T)1e-18);
}

template <typename T, int D, bool ISO>
TVDK_DEV void bw_from_t(const T (&t)[D], T kappa, T ell, T (&b)[D], T (&w)[D]) {
  if (ISO) {
    const T phi = iso_phi_of<T, D, true>(t, kappa);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const T d = t[a] * phi;
      b[a] = t[a] * ((T)1 - phi);
      w[a] = ell * (b[a] - d);
    }
  } else {
#pragma unroll
    for (int a = 0; a < D; ++a) {
      // sign(t) * max(|t| - kappa, 0); kappa >= 0 (validated by the host) makes the t = 0 case 0
      const T d = tcopysign(tmax(tabs(t[a]) - kappa, (T)0), t[a]);
      b[a] = t[a] - d;
      w[a] = ell * (b[a] - d);
    }
  }
}

// per-voxel scalars for the heterogeneous case
template <typename T>
struct VoxCoef {
  T kappa, ell, omega, inv_diag, c_nb;
};
template <bool HET, typename T>
TVDK_DEV VoxCoef<T> voxcoef(const Scalars<T>& sc, T mu, T om, T el) {
  VoxCoef<T> r;
  if (!HET) {
    r.kappa = sc.kappa;
    r.ell = sc.ell;
    r.omega = sc.omega;
    r.inv_diag = sc.inv_diag;
    r.c_nb = sc.c_nb;
  } else {
    r.kappa = tdiv(mu, el);
    r.ell = el;
    r.omega = om;
    const T diag = om + el * (T)2 * sc.dim_t;  // omega + ell*2*dim/h**2, h = 1 (jacobi.py:72)
    r.inv_diag = tdiv((T)1, diag);
    r.c_nb = el * r.inv_diag;  // N/(diag/ell) == N*ell/diag
  }
  return r;
}

// One axis of the reference Laplacian (derivatives.py:88-104) at one voxel:
//   1/2 [ G(l H x) + H(l G x) ],  G = backward_diff, H = forward_diff
template <typename T>
TVDK_DEV T lap_axis(T xm, T xc, T xp, bool has_lo, bool has_hi, T lm, T lc, T lp) {
  const T dp = has_hi ? xp - xc : (T)0;
  const T dm = has_lo ? xc - xm : (T)0;
  const T g = has_hi ? lp * dp - lc * dm : (T)0;
  const T h = has_lo ? lc * dp - lm * dm : (T)0;
  return (T)0.5 * (g + h);
}

// ------------------------------------------------------------------------------------------
// Neighbour helpers on a vector whose row has been published in shared memory
// ------------------------------------------------------------------------------------------
// N(x): sum of the 2*D edge-replicated neighbours in the reference's order (jacobi.py:40-55).
template <typename T, int D, bool E, typename C>
TVDK_DEV Vec<T> nsum(const C& c, const Vec<T>& lo, const Vec<T>& hi, const Vec<T>& xc, const T* row) {
  constexpr int V = Vec<T>::N;
  Vec<T> s;
#pragma unroll
  for (int v = 0; v < V; ++v) s.v[v] = lo.v[v] + hi.v[v];
  if (D == 3) {
    const Vec<T> ym = sym(row, c), yp = syp(row, c);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      s.v[v] += c.template YLO<E>() ? ym.v[v] : xc.v[v];
      s.v[v] += c.template YHI<E>() ? yp.v[v] : xc.v[v];
    }
  }
  const T xl = sxm(row, c), xr = sxp(row, c);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const T left = v > 0 ? xc.v[v > 0 ? v - 1 : 0] : xl;
    const T right = v < V - 1 ? xc.v[v < V - 1 ? v + 1 : 0] : xr;
    s.v[v] += c.template HL<E>(v) ? left : xc.v[v];
    s.v[v] += c.template HR<E>(v) ? right : xc.v[v];
  }
  return s;
}

// A x = omega x - L_ell x (cg.py:37-44) with x's z-neighbours in registers and its row in smem;
// `o` is the element offset of the vector (for the per-voxel-ell variant, which loads ell's
// neighbours straight from global memory: rare path, cg.py:43 / SURVEY Q18).
// OUT: a per-voxel ell multiplies the unit-coefficient Laplacian from OUTSIDE, `omega x - ell * laplace(x)` -- the
// multigrid residual operator (mg.py:94-105), unlike CG's `laplace(x, diffusion_coeff=ell)` (cg.py:37-44).
template <typename T, int D, bool E, bool HET, bool OUT = false, typename C>
TVDK_DEV Vec<T> apply_A(const C& c, const Geom& g, int z, long long o, const Vec<T>& lo, const Vec<T>& xc,
                        const Vec<T>& hi, const T* row, const Vec<T>& om, const Coef<T>& ell) {
  constexpr int V = Vec<T>::N;
  const bool zlo = c.template ZLO<E>(g, z), zhi = c.template ZHI<E>(g, z);
  Vec<T> lap;
  const T xl = sxm(row, c), xr = sxp(row, c);
  Vec<T> ym = xc, yp = xc;
  if (D == 3) {
    ym = sym(row, c);
    yp = syp(row, c);
  }
  const bool outside = OUT && HET && ell.p;
  if (!HET || !ell.p || outside) {
    const T l = outside ? (T)1 : ell.val;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T s = lap_axis(lo.v[v], xc.v[v], hi.v[v], zlo, zhi, l, l, l);
      if (D == 3) s = s + lap_axis(ym.v[v], xc.v[v], yp.v[v], c.template YLO<E>(), c.template YHI<E>(), l, l, l);
      const T left = v > 0 ? xc.v[v > 0 ? v - 1 : 0] : xl;
      const T right = v < V - 1 ? xc.v[v < V - 1 ? v + 1 : 0] : xr;
      s = s + lap_axis(left, xc.v[v], right, c.template HL<E>(v), c.template HR<E>(v), l, l, l);
      lap.v[v] = s;
    }
  } else {
    const T* lb = bptr(ell.p + (long long)c.b * ell.sb, o + c.cdel);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (!c.okc[v]) {
        lap.v[v] = (T)0;
        continue;
      }
      const T lc = sload_ro(lb + v);
      T s = lap_axis(lo.v[v], xc.v[v], hi.v[v], zlo, zhi, zlo ? sload_ro(lb + v - g.s0) : lc, lc,
                     zhi ? sload_ro(lb + v + g.s0) : lc);
      if (D == 3)
        s = s + lap_axis(ym.v[v], xc.v[v], yp.v[v], c.hylo, c.hyhi, c.hylo ? sload_ro(lb + v - g.pitch) : lc, lc,
                         c.hyhi ? sload_ro(lb + v + g.pitch) : lc);
      const T left = v > 0 ? xc.v[v > 0 ? v - 1 : 0] : xl;
      const T right = v < V - 1 ? xc.v[v < V - 1 ? v + 1 : 0] : xr;
      s = s + lap_axis(left, xc.v[v], right, c.hl[v], c.hr[v], c.hl[v] ? sload_ro(lb + v - 1) : lc, lc,
                       c.hr[v] ? sload_ro(lb + v + 1) : lc);
      lap.v[v] = s;
    }
  }
  Vec<T> y;
  if (outside) {
    const Vec<T> el = ldcoef<HET>(c, ell, o);
#pragma unroll
    for (int v = 0; v < V; ++v) y.v[v] = om.v[v] * xc.v[v] - el.v[v] * lap.v[v];
    return y;
  }
#pragma unroll
  for (int v = 0; v < V; ++v) y.v[v] = om.v[v] * xc.v[v] - lap.v[v];
  return y;
}

// raw planes of the compressed state (or of the explicit warm-start arrays) + the coefficients
// needed to turn them into b and w; loaded one plane ahead of their use
template <typename T, int D>
struct RawT {
  Vec<T> t[D], d[D], mu, el;
};
// (RO = false: plain loads, for state that the same kernel updates in place)
template <bool RO, bool HET, bool XOK, typename T, int D, typename C>
TVDK_DEV void load_rawt(RawT<T, D>& r, const C& c, const Coefs<T>& k, const T* t, long long tcs, const T* xb,
                        const T* xd, long long xcs, long long o) {
  if (!XOK || t) {  // XOK: the explicit warm-start arrays may be selected at run time
#pragma unroll
    for (int d = 0; d < D; ++d) r.t[d] = RO ? ldo(t + d * tcs, o) : ldw(t + d * tcs, o);
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      r.t[d] = ldo(xb + d * xcs, o);
      r.d[d] = ldo(xd + d * xcs, o);
    }
  }
  r.mu = ldcoef<HET>(c, k.mu, o);
  r.el = ldcoef<HET>(c, k.ell, o);
}
// b and w = ell (b - d) of one plane from its raw state
template <typename T, int D, bool ISO, bool HET>
TVDK_DEV void bw_plane(const RawT<T, D>& r, const Scalars<T>& sc, bool expl, Vec<T> (&b)[D], Vec<T> (&w)[D]) {
  constexpr int V = Vec<T>::N;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const VoxCoef<T> vc = voxcoef<HET>(sc, r.mu.v[v], (T)1, r.el.v[v]);
    if (!expl) {
      T t[D], bb[D], ww[D];
#pragma unroll
      for (int d = 0; d < D; ++d) t[d] = r.t[d].v[v];
      bw_from_t<T, D, ISO>(t, vc.kappa, vc.ell, bb, ww);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        b[d].v[v] = bb[d];
        w[d].v[v] = ww[d];
      }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        b[d].v[v] = r.t[d].v[v];
        w[d].v[v] = vc.ell * (r.t[d].v[v] - r.d[d].v[v]);
      }
    }
  }
}

// ==========================================================================================
// K1  fused outer iteration for Jacobi(maxiter=1):  rhs -> sweep -> shrink -> Bregman update
//     split_bregman_tvd.py:84-89 (rhs), jacobi.py:101-114 (sweep), :136-146 (shrink/update),
//     :161 (increment).  u+ is recomputed on a +1 halo (threads of the RH rows / right vector).
//     Shared rows: [slot*D + 0] = u(j), [slot*D + d] = w_d(j) (d >= 1), [2D + slot] = u+ ring.
// ==========================================================================================
template <typename T, int D, bool ISO, bool HET>
struct KFusedJ1 {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 2, NRED = 1, NROWS = 2 * D + 2;
  static constexpr int PF2D = ISO ? 3 : 2, PF3D = 3;  // L2 prefetch distance (profiles/r1_pf_sweep.txt)
  // 3-D launch shape (profiles/r1_tile3d_v2.txt): 64 planes per CTA (0.79 vs 0.76 of peak at 32); fp32 with
  // 16 threads = 64 columns per row like fp64 (0.82 vs 0.70 with 32 threads)
  static constexpr int ZC3D = 64;
  static constexpr int NTX3D = sizeof(T) == 4 ? 16 : 32;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* f;
    const T* u_in;
    T* u_out;
    const T* t_in;
    T* t_out;
    long long tcs_in, tcs_out;
    const tvdk_stat_t* skip;  // problems with skip[b].done are left untouched
    Fin fin;
  };
  C c;
  int j, rowe;
  Vec<T> u_lo, u_c, u_hi;     // u(j-1), u(j), u(j+1)
  Vec<T> w0_prev, w0_cur;     // ell (b_0 - d_0) at planes j-1, j
  Vec<T> wa_cur[D];           // w_a(j), a >= 1
  Vec<T> b_prev[D], b_cur[D];
  Vec<T> up_prev, om_f, invd, cnb;
  // prefetched raw data of plane j+1 (u of plane j+2)
  RawT<T, D> nx;
  Vec<T> nx_u, nx_f, nx_om;
  bool have_row;
  double red[NRED];

  TVDK_DEV void prefetch(const Args& a, int z, long long o) {  // raw data needed to process plane z (offset o)
    load_rawt<true, HET, false>(nx, c, a.k, a.t_in, a.tcs_in, (const T*)nullptr, (const T*)nullptr, 0, o);
    nx_f = ldo(a.f, o);
    nx_om = ldcoef<HET>(c, a.k.omega, o);
    if (c.has_zhi(a.g, z)) nx_u = ldo(a.u_in, o + c.s0b);
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = 0.0;
    j = c.z0;
    const long long o = c.at(a.g, j);
    u_c = ldo(a.u_in, o);
    u_lo = u_c;
    w0_prev = vsplat<T>((T)0);
    if (c.has_zlo(a.g, j)) {
      u_lo = ldo(a.u_in, o - c.s0b);
      RawT<T, D> r;
      load_rawt<true, HET, false>(r, c, a.k, a.t_in, a.tcs_in, (const T*)nullptr, (const T*)nullptr, 0, o - c.s0b);
      Vec<T> bb[D], ww[D];
      bw_plane<T, D, ISO, HET>(r, a.k.sc, false, bb, ww);
      w0_prev = ww[0];
    }
    up_prev = vsplat<T>((T)0);
#pragma unroll
    for (int d = 0; d < D; ++d) b_prev[d] = vsplat<T>((T)0);
    have_row = false;
    prefetch(a, j, o);
    return true;
  }
  TVDK_DEV bool more() const { return j <= c.z1; }

  // phase 1: consume the prefetched plane j (form b/w from t), publish u(j) and w_a(j) (a >= 1),
  //          then issue the loads of plane j+1
  // The march loop is unrolled by two in the sm_100a runner (UNROLL2): the phase PH = (j - z0) & 1
  // is then a compile-time constant, so the double-buffer slot of every shared row is an immediate
  // and the pipeline-register roll at the end of a step is renamed away across the two copies.
  // (measured, profiles/r1_ab_unroll2.txt: +1 % with scalar coefficients, -2 % with per-voxel maps,
  //  where the second copy of the loop body pushes the kernel over 128 registers)
  static constexpr bool UNROLL2 = !HET;
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    if ((j - c.z0) & 1) pre_p<1>(a, smem); else pre_p<0>(a, smem);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if ((j - c.z0) & 1) post_p<1>(a, smem); else post_p<0>(a, smem);
  }
  template <int PH>
  TVDK_DEV void pre_p(const Args& a, unsigned char* smem) {
    have_row = (j < c.z1) || c.plane_ok(a.g, j);
    if (!have_row) return;
    constexpr int slot = PH;
    u_hi = c.has_zhi(a.g, j) ? nx_u : u_c;
    Vec<T> ww[D];
    bw_plane<T, D, ISO, HET>(nx, a.k.sc, false, b_cur, ww);
    w0_cur = ww[0];
#pragma unroll
    for (int d = 1; d < D; ++d) wa_cur[d] = ww[d];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const VoxCoef<T> vc = voxcoef<HET>(a.k.sc, (T)0, nx_om.v[v], nx.el.v[v]);
      invd.v[v] = vc.inv_diag;
      cnb.v[v] = vc.c_nb;
      om_f.v[v] = vc.omega * nx_f.v[v];
    }
    sst(srow<T>(smem, slot * D + 0, rowe), c, u_c);
#pragma unroll
    for (int d = 1; d < D; ++d) sst(srow<T>(smem, slot * D + d, rowe), c, wa_cur[d]);
    // next plane is needed if another compute step follows
    if (j + 1 < c.z1 || (j + 1 == c.z1 && c.plane_ok(a.g, j + 1))) prefetch(a, j + 1, c.cur + c.s0b);
    if (c.pf_ok(j)) {
      const long long o = c.cur + c.pfb;
      pfo(a.u_in, o);
      pfo(a.f, o);
#pragma unroll
      for (int d = 0; d < D; ++d) pfo(a.t_in + d * a.tcs_in, o);
      if (HET) {
        if (a.k.mu.p) pfo(a.k.mu.p + (long long)c.b * a.k.mu.sb, o + c.cdel);
        if (a.k.omega.p) pfo(a.k.omega.p + (long long)c.b * a.k.omega.sb, o + c.cdel);
        if (a.k.ell.p) pfo(a.k.ell.p + (long long)c.b * a.k.ell.sb, o + c.cdel);
      }
    }
  }

  TVDK_DEV bool ZEDGE(const Args& a) const { return c.zedge(a.g, j) || c.zedge(a.g, j - 1); }
  template <bool E, int PH>
  TVDK_DEV void post_t(const Args& a, unsigned char* smem) {
    constexpr int slot = PH;
    Vec<T> up = vsplat<T>((T)0);
    if (have_row) {
      const T* urow = srow<T>(smem, slot * D + 0, rowe);
      // rhs = omega f + H_0 w_0 + H_1 w_1 (+ H_2 w_2), in the reference's order (:84-89)
      Vec<T> rhs = om_f;
      if (c.template ZLO<E>(a.g, j)) {
#pragma unroll
        for (int v = 0; v < V; ++v) rhs.v[v] += w0_cur.v[v] - w0_prev.v[v];
      }
      if (D == 3) {
        const Vec<T> wm = sym(srow<T>(smem, slot * D + 1, rowe), c);
        if (c.template YLO<E>()) {
#pragma unroll
          for (int v = 0; v < V; ++v) rhs.v[v] += wa_cur[1].v[v] - wm.v[v];
        }
      }
      {
        const T wl = sxm(srow<T>(smem, slot * D + (D - 1), rowe), c);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const T left = v > 0 ? wa_cur[D - 1].v[v > 0 ? v - 1 : 0] : wl;
          if (c.template HL<E>(v)) rhs.v[v] += wa_cur[D - 1].v[v] - left;
        }
      }
      const Vec<T> nb = nsum<T, D, E>(c, u_lo, u_hi, u_c, urow);
#pragma unroll
      for (int v = 0; v < V; ++v) up.v[v] = rhs.v[v] * invd.v[v] + nb.v[v] * cnb.v[v];
      sst(srow<T>(smem, 2 * D + slot, rowe), c, up);
    }
    if (j > c.z0) {
      // outputs of plane j-1: t_new = G u+ + b, u_out = u+, increment (:136-146, :161)
      const int zo = j - 1;
      const long long oo = c.cur - c.s0b;
      const T* prow = srow<T>(smem, 2 * D + (1 - PH), rowe);
      Vec<T> tn[D];
      const bool zhi = c.template ZHI<E>(a.g, zo);
#pragma unroll
      for (int v = 0; v < V; ++v) tn[0].v[v] = (zhi ? up.v[v] - up_prev.v[v] : (T)0) + b_prev[0].v[v];
      if (D == 3) {
        const Vec<T> yp = syp(prow, c);
#pragma unroll
        for (int v = 0; v < V; ++v)
          tn[1].v[v] = (c.template YHI<E>() ? yp.v[v] - up_prev.v[v] : (T)0) + b_prev[1].v[v];
      }
      {
        const T xr = sxp(prow, c);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const T right = v < V - 1 ? up_prev.v[v < V - 1 ? v + 1 : 0] : xr;
          tn[D - 1].v[v] = (c.template HR<E>(v) ? right - up_prev.v[v] : (T)0) + b_prev[D - 1].v[v];
        }
      }
      sto(c, a.u_out, oo, up_prev);
#pragma unroll
      for (int d = 0; d < D; ++d) sto(c, a.t_out + d * a.tcs_out, oo, tn[d]);
      if (c.st) {  // u_lo still holds u(j-1) here (rolled below)
#pragma unroll
        for (int v = 0; v < V; ++v) {
          if (c.template OK<E>(v)) {
            const double df = (double)up_prev.v[v] - (double)u_lo.v[v];
            red[0] += df * df;
          }
        }
      }
    }
    // roll the pipeline registers
    u_lo = u_c;
    u_c = u_hi;
    w0_prev = w0_cur;
    up_prev = up;
#pragma unroll
    for (int d = 0; d < D; ++d) b_prev[d] = b_cur[d];
    c.cur += c.s0b;
    ++j;
  }
  template <int PH>
  TVDK_DEV void post_p(const Args& a, unsigned char* smem) {
    if (c.edge || ZEDGE(a))
      post_t<true, PH>(a, smem);
    else
      post_t<false, PH>(a, smem);
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K1x2  TWO fused outer iterations for Jacobi(maxiter=1) in one pass (2-D, temporal blocking):
//     (u, t) -> (u', t') -> (u'', t'') with only u, f, t read and u'', t'' written: (3 + 2D)/2
//     words per voxel and outer iteration.  Four stages share the plane's single barrier, each one
//     plane behind the previous:   A1 plane j: u' ;  A2 plane j-1: t' (and its b', w') ;
//     B1 plane j-2: u'' ;  B2 plane j-3: t'' + stores.   Same arithmetic per voxel as K1
//     (split_bregman_tvd.py:84-89, jacobi.py:112-114, :136-146, :161 twice).
//     Shared rows (two slots each): 0 u(j) | 1 w1(j) | 2 u'(ring) | 3 u'(copy for B1) | 4 w1'(j-1) | 5 u''(ring).
// ==========================================================================================
template <typename T, bool ISO, bool HET>
struct KFusedJ1x2 {
  static constexpr int D = 2;
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 0, RH = 0, NRED = 2, NROWS = 12;
  static constexpr int RHXV = (3 + V - 1) / V;  // u'' at x+1 needs u' at x+2 needs u at x+3
  static constexpr int PF2D = 3, PF3D = 3;
  static constexpr int ZC2D = 64;  // 4 extra march steps per chunk
  static constexpr int MAXT = 128;  // block size cap: the four-stage pipeline needs > 128 registers per thread
  typedef Col<T, D, LH, RH, 1, RHXV> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* f;
    const T* u_in;
    T* u_out;
    const T* t_in;
    T* t_out;
    long long tcs_in, tcs_out;
    const tvdk_stat_t* skip;
    Fin fin;
  };
  C c;
  int j, rowe, js, jend;
  // stage A (iteration 1)
  Vec<T> u_lo, u_c, u_hi;          // u(j-1), u(j), u(j+1)
  Vec<T> wA0_prev, wA0_cur, wA1;   // w_0(j-1), w_0(j), w_1(j)
  Vec<T> bA_prev[D], bA_cur[D];    // b(j-1), b(j)
  Vec<T> upA1, upA2, upA3;         // u'(j-1), u'(j-2), u'(j-3)
  Vec<T> omfA, invdA, cnbA;        // per-voxel scalars of plane j
  Vec<T> mu0, el0, mu1, el1;       // mu, ell of planes j and j-1 (per-voxel maps only)
  // stage B (iteration 2)
  Vec<T> wB0_p2, wB0_p3, wB1_p2;   // w'_0(j-2), w'_0(j-3), w'_1(j-2)
  Vec<T> bB_p2[D], bB_p3[D];       // b'(j-2), b'(j-3)
  Vec<T> upB_p3;                   // u''(j-3)
  Vec<T> lf, lom, lel;             // f, omega, ell of plane j-2 (re-read: L1/L2 hits)
  // prefetched raw data of plane j+1 (u of plane j+2)
  RawT<T, D> nx;
  Vec<T> nx_u, nx_f, nx_om;
  double red[NRED];

  TVDK_DEV void prefetch(const Args& a, long long o, bool zhi) {
    load_rawt<true, HET, false>(nx, c, a.k, a.t_in, a.tcs_in, (const T*)nullptr, (const T*)nullptr, 0, o);
    nx_f = ldo(a.f, o);
    nx_om = ldcoef<HET>(c, a.k.omega, o);
    if (zhi) nx_u = ldo(a.u_in, o + c.s0b);
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = red[1] = 0.0;
    js = c.z0 - 1 > 0 ? c.z0 - 1 : 0;  // first plane of stage A1
    jend = c.z1 + 2;                   // stage B2 writes plane j - 3
    j = js;
    const long long o = c.at(a.g, j);
    c.cur = o;
    u_c = ldo(a.u_in, o);
    u_lo = u_c;
    const Vec<T> z = vsplat<T>((T)0);
    wA0_prev = z;
    if (j > 0) {
      u_lo = ldo(a.u_in, o - c.s0b);
      RawT<T, D> r;
      load_rawt<true, HET, false>(r, c, a.k, a.t_in, a.tcs_in, (const T*)nullptr, (const T*)nullptr, 0, o - c.s0b);
      Vec<T> bb[D], ww[D];
      bw_plane<T, D, ISO, HET>(r, a.k.sc, false, bb, ww);
      wA0_prev = ww[0];
    }
    upA1 = upA2 = upA3 = upB_p3 = z;
    wB0_p2 = wB0_p3 = wB1_p2 = z;
    mu0 = el0 = mu1 = el1 = lf = lom = lel = vsplat<T>((T)1);
#pragma unroll
    for (int d = 0; d < D; ++d) bA_prev[d] = bB_p2[d] = bB_p3[d] = z;
    prefetch(a, o, j + 1 < a.g.n0);
    return true;
  }
  TVDK_DEV bool more() const { return j <= jend; }

  static constexpr bool UNROLL2 = !HET;
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    if ((j - js) & 1) pre_p<1>(a, smem); else pre_p<0>(a, smem);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if ((j - js) & 1) post_p<1>(a, smem); else post_p<0>(a, smem);
  }

  // rows: index = 2 * kind + slot
  template <int KIND>
  TVDK_DEV T* row(unsigned char* smem, int slot) const { return srow<T>(smem, 2 * KIND + slot, rowe); }

  // pre: consume the prefetched plane j (b, w from t; per-voxel scalars), publish u(j) and w_1(j),
  //      issue the loads of plane j+1 and the re-reads of plane j-2
  template <int PH>
  TVDK_DEV void pre_p(const Args& a, unsigned char* smem) {
    if (j < a.g.n0) {
      u_hi = (j + 1 < a.g.n0) ? nx_u : u_c;
      Vec<T> ww[D];
      bw_plane<T, D, ISO, HET>(nx, a.k.sc, false, bA_cur, ww);
      wA0_cur = ww[0];
      wA1 = ww[1];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const VoxCoef<T> vc = voxcoef<HET>(a.k.sc, (T)0, nx_om.v[v], nx.el.v[v]);
        invdA.v[v] = vc.inv_diag;
        cnbA.v[v] = vc.c_nb;
        omfA.v[v] = vc.omega * nx_f.v[v];
      }
      sst(row<0>(smem, PH), c, u_c);
      sst(row<1>(smem, PH), c, wA1);
      mu0 = nx.mu;  // stage A2 of the NEXT step shrinks t'(j) with plane j's kappa
      el0 = nx.el;
      // loads of plane j+1 (u' is needed up to plane z1 + 1 only), issued before the barrier so that
      // the four stages after it hide their latency
      if (j + 1 < a.g.n0 && j + 1 <= c.z1 + 1) prefetch(a, c.cur + c.s0b, j + 2 < a.g.n0);
    }
    if (j - 2 >= 0 && j - 2 < a.g.n0) {
      const long long o2 = c.cur - 2 * c.s0b;
      lf = ldo(a.f, o2);
      lom = ldcoef<HET>(c, a.k.omega, o2);
      lel = ldcoef<HET>(c, a.k.ell, o2);
    }
    if (c.pf_ok(j)) {
      const long long o = c.cur + c.pfb;
      pfo(a.u_in, o);
      pfo(a.f, o);
#pragma unroll
      for (int d = 0; d < D; ++d) pfo(a.t_in + d * a.tcs_in, o);
      if (HET) {
        if (a.k.mu.p) pfo(a.k.mu.p + (long long)c.b * a.k.mu.sb, o + c.cdel);
        if (a.k.omega.p) pfo(a.k.omega.p + (long long)c.b * a.k.omega.sb, o + c.cdel);
        if (a.k.ell.p) pfo(a.k.ell.p + (long long)c.b * a.k.ell.sb, o + c.cdel);
      }
    }
  }

  // u+ of one plane: rhs (:84-89) then the sweep (jacobi.py:112-114)
  template <bool E>
  TVDK_DEV Vec<T> ustep(bool zlo, const Vec<T>& lo, const Vec<T>& cen, const Vec<T>& hi, const T* urow, const Vec<T>& w0c,
                        const Vec<T>& w0p, const Vec<T>& w1c, const T* w1row, const Vec<T>& omf, const Vec<T>& invd,
                        const Vec<T>& cnb) const {
    Vec<T> rhs = omf;
    if (zlo) {
#pragma unroll
      for (int v = 0; v < V; ++v) rhs.v[v] += w0c.v[v] - w0p.v[v];
    }
    const T wl = sxm(w1row, c);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const T left = v > 0 ? w1c.v[v > 0 ? v - 1 : 0] : wl;
      if (c.template HL<E>(v)) rhs.v[v] += w1c.v[v] - left;
    }
    const Vec<T> nb = nsum<T, D, E>(c, lo, hi, cen, urow);
    Vec<T> up;
#pragma unroll
    for (int v = 0; v < V; ++v) up.v[v] = rhs.v[v] * invd.v[v] + nb.v[v] * cnb.v[v];
    return up;
  }
  // t+ of one plane: G u+ + b (:136-146 on the compressed state)
  template <bool E>
  TVDK_DEV void tstep(bool zhi, const Vec<T>& up_hi, const Vec<T>& up_c, const T* uprow, const Vec<T> (&b)[D],
                      Vec<T> (&tn)[D]) const {
#pragma unroll
    for (int v = 0; v < V; ++v) tn[0].v[v] = (zhi ? up_hi.v[v] - up_c.v[v] : (T)0) + b[0].v[v];
    const T xr = sxp(uprow, c);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const T right = v < V - 1 ? up_c.v[v < V - 1 ? v + 1 : 0] : xr;
      tn[1].v[v] = (c.template HR<E>(v) ? right - up_c.v[v] : (T)0) + b[1].v[v];
    }
  }

  template <bool E, bool ZE, int PH>
  TVDK_DEV void post_t(const Args& a, unsigned char* smem) {
    const int n0 = a.g.n0;
    // ---- A1, plane j: u'(j)
    const bool exA1 = ZE ? (j < n0) : true;
    Vec<T> upA0 = upA1;
    if (exA1) {
      upA0 = ustep<E>(ZE ? j > 0 : true, ZE && j == 0 ? u_c : u_lo, u_c, u_hi, row<0>(smem, PH), wA0_cur, wA0_prev, wA1,
                      row<1>(smem, PH), omfA, invdA, cnbA);
      sst(row<2>(smem, PH), c, upA0);
    }
    // ---- A2, plane j-1: t'(j-1), its b', w'; increment of iteration 1
    const int p1 = j - 1;
    const bool exA2 = ZE ? (p1 >= js && p1 < n0) : true;
    Vec<T> bB_new[D], wB_new[D];
#pragma unroll
    for (int d = 0; d < D; ++d) bB_new[d] = wB_new[d] = upA1;
    if (exA2) {
      Vec<T> tB[D];
      tstep<E>(ZE ? p1 + 1 < n0 : true, upA0, upA1, row<2>(smem, PH ^ 1), bA_prev, tB);
      RawT<T, D> r;
#pragma unroll
      for (int d = 0; d < D; ++d) r.t[d] = tB[d];
      r.mu = mu1;
      r.el = el1;
      bw_plane<T, D, ISO, HET>(r, a.k.sc, false, bB_new, wB_new);
      sst(row<4>(smem, PH ^ 1), c, wB_new[1]);
      sst(row<3>(smem, PH ^ 1), c, upA1);
      if (c.st && p1 >= c.z0 && p1 < c.z1) {  // u_lo still holds u(j-1)
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) {
            const double df = (double)upA1.v[v] - (double)u_lo.v[v];
            red[0] += df * df;
          }
      }
    }
    // ---- B1, plane j-2: u''(j-2)
    const int p2 = j - 2;
    const bool exB1 = ZE ? (p2 >= js && p2 < n0) : true;
    Vec<T> upB_new = upB_p3;
    if (exB1) {
      Vec<T> omf, invd, cnb;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const VoxCoef<T> vc = voxcoef<HET>(a.k.sc, (T)0, lom.v[v], lel.v[v]);
        invd.v[v] = vc.inv_diag;
        cnb.v[v] = vc.c_nb;
        omf.v[v] = vc.omega * lf.v[v];
      }
      const bool zlo = ZE ? p2 > 0 : true, zhi = ZE ? p2 + 1 < n0 : true;
      upB_new = ustep<E>(zlo, zlo ? upA3 : upA2, upA2, zhi ? upA1 : upA2, row<3>(smem, PH), wB0_p2, wB0_p3, wB1_p2,
                         row<4>(smem, PH), omf, invd, cnb);
      sst(row<5>(smem, PH), c, upB_new);
    }
    // ---- B2, plane j-3: t''(j-3), stores, increment of iteration 2
    const int p3 = j - 3;
    if (p3 >= c.z0 && p3 < c.z1) {
      Vec<T> tC[D];
      tstep<E>(ZE ? p3 + 1 < n0 : true, upB_new, upB_p3, row<5>(smem, PH ^ 1), bB_p3, tC);
      const long long oo = c.cur - 3 * c.s0b;
      sto(c, a.u_out, oo, upB_p3);
#pragma unroll
      for (int d = 0; d < D; ++d) sto(c, a.t_out + d * a.tcs_out, oo, tC[d]);
      if (c.st) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) {
            const double df = (double)upB_p3.v[v] - (double)upA3.v[v];
            red[1] += df * df;
          }
      }
    }
    // ---- roll, then the loads of plane j+1
    u_lo = u_c;
    u_c = u_hi;
    wA0_prev = wA0_cur;
    mu1 = mu0;
    el1 = el0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      bA_prev[d] = bA_cur[d];
      bB_p3[d] = bB_p2[d];
      bB_p2[d] = bB_new[d];
    }
    upA3 = upA2;
    upA2 = upA1;
    upA1 = upA0;
    wB0_p3 = wB0_p2;
    wB0_p2 = wB_new[0];
    wB1_p2 = wB_new[1];
    upB_p3 = upB_new;
    c.cur += c.s0b;
    ++j;
  }
  template <int PH>
  TVDK_DEV void post_p(const Args& a, unsigned char* smem) {
    const bool ze = j - 3 <= js || j + 1 >= a.g.n0;  // some stage's plane is missing, first, or on the axis-0 boundary
    if (ze) {
      post_t<true, true, PH>(a, smem);
    } else if (c.edge) {
      post_t<true, false, PH>(a, smem);
    } else {
      post_t<false, false, PH>(a, smem);
    }
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K2  rhs from the compressed state, optionally fused with the first Jacobi sweep (MODE 1)
//     or with CG's r0 = rhs - A u, x = u (MODE 2).   split_bregman_tvd.py:84-89; jacobi.py:107-114;
//     cg.py:37-47 + scipy cg (r = b - A x).  With t == nullptr the first outer iteration of a
//     warm start reads the reference's explicit d and b (split_bregman_tvd.py:93-97).
//     Shared rows: [slot*D + 0] = u(j), [slot*D + d] = w_d(j) (d >= 1).
// ==========================================================================================
enum { RHS_ONLY = 0, RHS_SWEEP = 1, RHS_CGINIT = 2 };

template <typename T, int D, bool ISO, int MODE, bool HET>
struct KRhs {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 1, NRED = 2, NROWS = 2 * D;
  static constexpr int PF2D = 3, PF3D = 3;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* f;
    const T* t;        // compressed state, or nullptr when the warm-start arrays are used
    long long tcs;
    const T* xd;       // explicit d0, b0 (warm start), component stride xcs
    const T* xb;
    long long xcs;
    const T* u;        // MODE 1/2: current iterate
    T* rhs;            // MODE 0/1: rhs out
    T* x;              // MODE 1: first sweep out; MODE 2: copy of u
    T* r;              // MODE 2: r0
    const tvdk_stat_t* skip;
    Fin fin;
  };
  C c;
  int j, rowe;
  Vec<T> u_lo, u_c, u_hi, w0_prev, w0_cur, wa_cur[D], om_c, f_c, el_c;
  RawT<T, D> nx;
  Vec<T> nx_u, nx_f, nx_om;
  double red[NRED];

  TVDK_DEV void prefetch(const Args& a, int z, long long o) {
    load_rawt<true, HET, true>(nx, c, a.k, a.t, a.tcs, a.xb, a.xd, a.xcs, o);
    nx_f = ldo(a.f, o);
    nx_om = ldcoef<HET>(c, a.k.omega, o);
    if (MODE != RHS_ONLY && c.has_zhi(a.g, z)) nx_u = ldo(a.u, o + c.s0b);
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = red[1] = 0.0;
    j = c.z0;
    const long long o = c.at(a.g, j);
    w0_prev = vsplat<T>((T)0);
    u_c = vsplat<T>((T)0);
    if (MODE != RHS_ONLY) u_c = ldo(a.u, o);
    u_lo = u_c;
    if (c.has_zlo(a.g, j)) {
      RawT<T, D> r;
      load_rawt<true, HET, true>(r, c, a.k, a.t, a.tcs, a.xb, a.xd, a.xcs, o - c.s0b);
      Vec<T> bb[D], ww[D];
      bw_plane<T, D, ISO, HET>(r, a.k.sc, a.t == nullptr, bb, ww);
      w0_prev = ww[0];
      if (MODE != RHS_ONLY) u_lo = ldo(a.u, o - c.s0b);
    }
    prefetch(a, j, o);
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }

  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    const int slot = j & 1;
    Vec<T> bb[D], ww[D];
    bw_plane<T, D, ISO, HET>(nx, a.k.sc, a.t == nullptr, bb, ww);
    w0_cur = ww[0];
#pragma unroll
    for (int d = 1; d < D; ++d) wa_cur[d] = ww[d];
    f_c = nx_f;
    om_c = nx_om;
    el_c = nx.el;
    if (MODE != RHS_ONLY) {
      u_hi = c.has_zhi(a.g, j) ? nx_u : u_c;
      sst(srow<T>(smem, slot * D + 0, rowe), c, u_c);
    }
#pragma unroll
    for (int d = 1; d < D; ++d) sst(srow<T>(smem, slot * D + d, rowe), c, wa_cur[d]);
    if (j + 1 < c.z1) prefetch(a, j + 1, c.cur + c.s0b);
    if (c.pf_ok(j)) {
      const long long o = c.cur + c.pfb;
      pfo(a.f, o);
      if (a.t) {
#pragma unroll
        for (int d = 0; d < D; ++d) pfo(a.t + d * a.tcs, o);
      }
      if (MODE != RHS_ONLY) pfo(a.u, o);
    }
  }

  TVDK_DEV bool ZEDGE(const Args& a) const { return c.zedge(a.g, j); }
  template <bool E>
  TVDK_DEV void post_t(const Args& a, unsigned char* smem) {
    const int slot = j & 1;
    Vec<T> rhs;
#pragma unroll
    for (int v = 0; v < V; ++v) rhs.v[v] = om_c.v[v] * f_c.v[v];
    if (c.template ZLO<E>(a.g, j)) {
#pragma unroll
      for (int v = 0; v < V; ++v) rhs.v[v] += w0_cur.v[v] - w0_prev.v[v];
    }
    if (D == 3) {
      const Vec<T> wm = sym(srow<T>(smem, slot * D + 1, rowe), c);
      if (c.template YLO<E>()) {
#pragma unroll
        for (int v = 0; v < V; ++v) rhs.v[v] += wa_cur[1].v[v] - wm.v[v];
      }
    }
    {
      const T wl = sxm(srow<T>(smem, slot * D + (D - 1), rowe), c);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const T left = v > 0 ? wa_cur[D - 1].v[v > 0 ? v - 1 : 0] : wl;
        if (c.template HL<E>(v)) rhs.v[v] += wa_cur[D - 1].v[v] - left;
      }
    }
    if (MODE == RHS_ONLY || MODE == RHS_SWEEP) sto(c, a.rhs, c.cur, rhs);
    if (MODE == RHS_SWEEP) {
      const Vec<T> nb = nsum<T, D, E>(c, u_lo, u_hi, u_c, srow<T>(smem, slot * D + 0, rowe));
      Vec<T> xn;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const VoxCoef<T> vc = voxcoef<HET>(a.k.sc, (T)0, om_c.v[v], el_c.v[v]);
        xn.v[v] = rhs.v[v] * vc.inv_diag + nb.v[v] * vc.c_nb;
      }
      sto(c, a.x, c.cur, xn);
      if (c.st) {  // Jacobi tol branch, first sweep: ||x_new - x||^2 and ||x0||^2
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) {
            const double df = (double)xn.v[v] - (double)u_c.v[v];
            red[0] += df * df;
            red[1] += (double)u_c.v[v] * (double)u_c.v[v];
          }
      }
    }
    if (MODE == RHS_CGINIT) {
      const Vec<T> au = apply_A<T, D, E, HET>(c, a.g, j, c.cur, u_lo, u_c, u_hi, srow<T>(smem, slot * D + 0, rowe),
                                         om_c, a.k.ell);
      Vec<T> r0;
#pragma unroll
      for (int v = 0; v < V; ++v) r0.v[v] = rhs.v[v] - au.v[v];
      sto(c, a.r, c.cur, r0);
      sto(c, a.x, c.cur, u_c);
      if (c.st) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) {
            red[0] += (double)rhs.v[v] * (double)rhs.v[v];
            red[1] += (double)r0.v[v] * (double)r0.v[v];
          }
      }
    }
    u_lo = u_c;
    u_c = u_hi;
    w0_prev = w0_cur;
    c.cur += c.s0b;
    ++j;
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if (c.edge || ZEDGE(a))
      post_t<true>(a, smem);
    else
      post_t<false>(a, smem);
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K4  Jacobi sweep  x_out = rhs/diag + N(x_in)/(diag/ell)       (jacobi.py:101-131)
// ==========================================================================================
template <typename T, int D, bool HET>
struct KSweep {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 1, NRED = 2, NROWS = 2;
  static constexpr int PF2D = 2, PF3D = 4;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* x_in;
    const T* rhs;
    T* x_out;
    const tvdk_stat_t* skip;  // done / jac_stop
    Fin fin;
  };
  C c;
  int j, rowe;
  Vec<T> x_lo, x_c, x_hi, nx_x, nx_rhs, rhs_c;
  double red[NRED];

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && (a.skip[lc.bz].done || a.skip[lc.bz].jac_stop)) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = red[1] = 0.0;
    j = c.z0;
    const long long o = c.at(a.g, j);
    x_c = ldo(a.x_in, o);
    x_lo = c.has_zlo(a.g, j) ? ldo(a.x_in, o - c.s0b) : x_c;
    nx_rhs = ldo(a.rhs, o);
    if (c.has_zhi(a.g, j)) nx_x = ldo(a.x_in, o + c.s0b);
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    x_hi = c.has_zhi(a.g, j) ? nx_x : x_c;
    rhs_c = nx_rhs;
    sst(srow<T>(smem, j & 1, rowe), c, x_c);
    if (j + 1 < c.z1) {
      nx_rhs = ldo(a.rhs, c.cur + c.s0b);
      if (c.has_zhi(a.g, j + 1)) nx_x = ldo(a.x_in, c.cur + 2 * c.s0b);
    }
    if (c.pf_ok(j)) {
      pfo(a.rhs, c.cur + c.pfb);
      pfo(a.x_in, c.cur + c.pfb);
    }
  }
  TVDK_DEV bool ZEDGE(const Args& a) const { return c.zedge(a.g, j); }
  template <bool E>
  TVDK_DEV void post_t(const Args& a, unsigned char* smem) {
    const Vec<T> nb = nsum<T, D, E>(c, x_lo, x_hi, x_c, srow<T>(smem, j & 1, rowe));
    const Vec<T> om = ldcoef<HET>(c, a.k.omega, c.cur), el = ldcoef<HET>(c, a.k.ell, c.cur);
    Vec<T> xn;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const VoxCoef<T> vc = voxcoef<HET>(a.k.sc, (T)0, om.v[v], el.v[v]);
      xn.v[v] = rhs_c.v[v] * vc.inv_diag + nb.v[v] * vc.c_nb;
    }
    sto(c, a.x_out, c.cur, xn);
    if (c.st) {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (c.template OK<E>(v)) {
          const double df = (double)xn.v[v] - (double)x_c.v[v];
          red[0] += df * df;
          red[1] += (double)x_c.v[v] * (double)x_c.v[v];
        }
    }
    x_lo = x_c;
    x_c = x_hi;
    c.cur += c.s0b;
    ++j;
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if (c.edge || ZEDGE(a))
      post_t<true>(a, smem);
    else
      post_t<false>(a, smem);
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K4xS  S Jacobi sweeps in ONE pass (temporal blocking; scalar coefficients, no tol branch):
//         x^(s) = rhs/diag + N(x^(s-1))/(diag/ell),  s = 1..S          (jacobi.py:109-114)
//     Stage s runs one plane behind stage s-1: at march step j it forms x^(s) of plane j-s+1 from
//     the three planes of x^(s-1) it keeps in registers and the in-plane neighbours that stage s-1
//     published one step earlier (one barrier per plane for all stages).  x and rhs are read once
//     and x written once per S sweeps: 3/S instead of 3 words per voxel and sweep.  The price is a
//     halo of S cells (rows / elements) around every tile and 2(S-1) extra planes per chunk.
//     Shared rows: [2 s + slot] = x^(s) of the plane stage s wrote last (s = 0: the input).
// ==========================================================================================
template <typename T, int D, int S>
struct KSweepN {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = S, RH = S, NRED = 0, NROWS = 2 * S;
  static constexpr int HX = (S + V - 1) / V;  // halo vectors left / right
  static constexpr int PF2D = 3, PF3D = 3;
  static constexpr int ZC2D = 64;
  typedef Col<T, D, LH, RH, HX, HX> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* x_in;
    const T* rhs;
    T* x_out;
    const tvdk_stat_t* skip;  // done
    Fin fin;
  };
  C c;
  int j, rowe, zlast;
  Vec<T> xs[S][2];    // xs[s][0], xs[s][1] = x^(s) at planes (j-s-1), (j-s): the two older planes stage s+1 needs
  Vec<T> rh[S];       // rhs at planes j, j-1, .., j-S+1
  Vec<T> nx_x, nx_r;  // x^(0)(j+1) and rhs(j+1) loaded one step ahead
  double red[1];

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    // first plane any stage touches / last march step (clamped to the planes that exist)
    int zs = c.z0 - (S - 1);
    if (zs < 0) zs = 0;
    zlast = c.z1 - 1 + (S - 1);
    j = zs;
    zs0 = zs;
    const long long o = c.at(a.g, j);
    c.cur = o;
    const Vec<T> x0 = ldo(a.x_in, o);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      xs[s][0] = x0;
      xs[s][1] = x0;
      rh[s] = x0;
    }
    xs[0][0] = c.has_zlo(a.g, j) ? ldo(a.x_in, o - c.s0b) : x0;
    nx_r = ldo(a.rhs, o);
    nx_x = c.has_zhi(a.g, j) ? ldo(a.x_in, o + c.s0b) : x0;
    return true;
  }
  TVDK_DEV bool more() const { return j <= zlast; }

  // The sm_100a runner unrolls the march loop by two (UNROLL2): the phase PH = (j - zs) & 1 is a
  // compile-time constant there, so every double-buffer slot below is an immediate and the window
  // rolls are renamed away across the two copies of the body.
  static constexpr bool UNROLL2 = true;
  int zs0;  // first march step (slot parity is counted from it)
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    if ((j - zs0) & 1) pre_p<1>(a, smem); else pre_p<0>(a, smem);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if ((j - zs0) & 1) post_p<1>(a, smem); else post_p<0>(a, smem);
  }
  // pre: plane j of the input becomes current (x^(0)(j) = xs[0][1]); publish its row
  template <int PH>
  TVDK_DEV void pre_p(const Args& a, unsigned char* smem) {
    if (!c.plane_ok(a.g, j)) return;  // past the last plane: only the later stages still run
    sst(srow<T>(smem, 0 + PH, rowe), c, xs[0][1]);
  }

  // E: the block touches the in-plane boundary; ZE: some plane a stage touches in this step is missing or
  // lies on the axis-0 boundary (block-uniform), otherwise every stage runs select-free
  template <bool E, bool ZE, int PH>
  TVDK_DEV void post_t(const Args& a, unsigned char* smem) {
    // roll rhs history: rh[0] = rhs(j)
#pragma unroll
    for (int s = S - 1; s > 0; --s) rh[s] = rh[s - 1];
    rh[0] = nx_r;
    const bool have0 = ZE ? c.plane_ok(a.g, j) : true;
    Vec<T> hi = (!ZE || (have0 && c.has_zhi(a.g, j))) ? nx_x : xs[0][1];  // x^(0)(j+1)
    if (have0 && j + 1 <= zlast && (!ZE || c.plane_ok(a.g, j + 1))) {    // loads for the next step
      nx_r = ldo(a.rhs, c.cur + c.s0b);
      if (j + 2 < a.g.n0) nx_x = ldo(a.x_in, c.cur + 2 * c.s0b);
    }
    if (c.pf_ok(j)) {
      pfo(a.rhs, c.cur + c.pfb);
      pfo(a.x_in, c.cur + c.pfb);
    }
    // stage s (1-based) forms x^(s) of plane p = j - (s - 1) from x^(s-1) at p-1, p, p+1
#pragma unroll
    for (int s = 1; s <= S; ++s) {
      const int p = j - (s - 1);
      const int slot_in = PH ^ ((s - 1) & 1);  // slot of plane p in the rows of x^(s-1) (and of x^(s))
      const bool ok = ZE ? (p >= 0 && p < a.g.n0) : true;  // whole volumes only (tvdk_solve)
      Vec<T> y = xs[s - 1][1];
      if (ok) {
        const Vec<T> lo = xs[s - 1][0], cen = xs[s - 1][1];
        const Vec<T> up = (!ZE || p + 1 < a.g.n0) ? hi : cen;
        const Vec<T> dn = (!ZE || p > 0) ? lo : cen;
        const Vec<T> nb = nsum<T, D, E>(c, dn, up, cen, srow<T>(smem, 2 * (s - 1) + slot_in, rowe));
#pragma unroll
        for (int v = 0; v < V; ++v) y.v[v] = rh[s - 1].v[v] * a.k.sc.inv_diag + nb.v[v] * a.k.sc.c_nb;
        if (s < S) {
          sst(srow<T>(smem, 2 * s + slot_in, rowe), c, y);
        } else if (p >= c.z0 && p < c.z1) {
          sto(c, a.x_out, c.cur - (long long)(S - 1) * c.s0b, y);
        }
      }
      // roll stage s-1's window, hand the new plane to the next stage
      xs[s - 1][0] = xs[s - 1][1];
      xs[s - 1][1] = hi;
      hi = y;
    }
    c.cur += c.s0b;
    ++j;
  }
  template <int PH>
  TVDK_DEV void post_p(const Args& a, unsigned char* smem) {
    const bool ze = j - S < 0 || j + 1 >= a.g.n0;
    if (ze) {
      post_t<true, true, PH>(a, smem);
    } else if (c.edge) {
      post_t<true, false, PH>(a, smem);
    } else {
      post_t<false, false, PH>(a, smem);
    }
  }
  static TVDK_DEV void finalize(const Args&, int, const double*) {}
};

// ==========================================================================================
// K5/K6  CG iteration (scipy.sparse.linalg.cg recurrence, called from cg.py:50-56)
//   PQ:     p_out = first ? r : beta p_in + r ;  reduce p_out . (A p_out)
//   UPDATE: x += alpha p ; r -= alpha (A p) ;    reduce r . r      (q = A p is recomputed,
//           not stored: 8 instead of 10 words per voxel and CG iteration)
//   MATVEC: y = A x (test-granularity operator)
//   RESID:  y = rhs - A x  (multigrid residual, mg.py:212; rhs in r_in)
//   FIRST:  reduce r.(A r) and (A r).(A r)  -- opens the single-pass route (K7): p_0 = r_0 is
//           never written, alpha_0 and the predicted rho_1 come from these two sums
// ==========================================================================================
enum { CG_PQ = 0, CG_UPDATE = 1, CG_MATVEC = 2, CG_FIRST = 3, CG_RESID = 4 };

template <typename T, int D, int MODE, bool HET>
struct KCg {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 1, NRED = MODE == CG_FIRST ? 2 : 1, NROWS = 2;
  static constexpr int PF2D = MODE == CG_UPDATE ? 1 : (MODE == CG_FIRST ? 4 : 3);
  static constexpr int PF3D = MODE == CG_UPDATE ? 3 : 4;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* p_in;  // PQ: old p ; UPDATE: p ; MATVEC: x ; FIRST: r
    const T* r_in;  // PQ: r
    T* p_out;       // PQ: new p ; MATVEC: y
    T* x;           // UPDATE (in place)
    T* r;           // UPDATE (in place)
    DRef coef;      // PQ: beta ; UPDATE: alpha
    IRef first;     // PQ: first spin (p = r)
    int first_all;  // PQ: force first spin for all problems
    const tvdk_stat_t* skip;  // done / cg_stop
    Fin fin;
  };
  C c;
  int j, rowe;
  T coef;
  bool first;
  Vec<T> p_lo, p_c, p_hi;
  Vec<T> nx_p, nx_r;  // raw plane j+1 (PQ: old p and r; else p), loaded one step ahead
  Vec<T> cx, cr;      // UPDATE: x(j), r(j), loaded one step ahead (interior threads only)
  double red[NRED];

  TVDK_DEV void load_raw(const Args& a, long long o) {
    if (MODE == CG_PQ) {
      nx_r = ldo(a.r_in, o);
      if (!first) nx_p = ldo(a.p_in, o);
    } else {
      nx_p = ldo(a.p_in, o);
    }
  }
  TVDK_DEV Vec<T> make_p() const {
    if (MODE == CG_PQ) {
      if (first) return nx_r;
      Vec<T> o;
#pragma unroll
      for (int v = 0; v < V; ++v) o.v[v] = coef * nx_p.v[v] + nx_r.v[v];  // p *= beta; p += r
      return o;
    }
    return nx_p;
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && (a.skip[lc.bz].done || a.skip[lc.bz].cg_stop)) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
#pragma unroll
    for (int r = 0; r < NRED; ++r) red[r] = 0.0;
    coef = (MODE == CG_MATVEC || MODE == CG_FIRST || MODE == CG_RESID || !a.coef.p) ? (T)0 : (T)a.coef.at(c.b);
    first = (MODE == CG_PQ) && (a.first_all || a.first.at(c.b));
    j = c.z0;
    const long long o = c.at(a.g, j);
    load_raw(a, o);
    p_c = make_p();
    p_lo = p_c;
    if (c.has_zlo(a.g, j)) {
      load_raw(a, o - c.s0b);
      p_lo = make_p();
    }
    if (c.has_zhi(a.g, j)) load_raw(a, o + c.s0b);
    if (MODE == CG_UPDATE && c.st) {
      cx = ldw(a.x, o);
      cr = ldw(a.r, o);
    }
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    p_hi = c.has_zhi(a.g, j) ? make_p() : p_c;
    sst(srow<T>(smem, j & 1, rowe), c, p_c);
    if (j + 1 < c.z1 && c.has_zhi(a.g, j + 1)) load_raw(a, c.cur + 2 * c.s0b);
    if (c.pf_ok(j)) {
      const long long o = c.cur + c.pfb;
      if (MODE != CG_PQ || !first) pfo(a.p_in, o);
      if (MODE == CG_PQ) pfo(a.r_in, o);
      if (MODE == CG_UPDATE && c.st) {
        pfo(a.x, o);
        pfo(a.r, o);
      }
    }
  }
  TVDK_DEV bool ZEDGE(const Args& a) const { return c.zedge(a.g, j); }
  template <bool E>
  TVDK_DEV void post_t(const Args& a, unsigned char* smem) {
    const Vec<T> om = ldcoef<HET>(c, a.k.omega, c.cur);
    const Vec<T> q = apply_A<T, D, E, HET, MODE == CG_RESID>(c, a.g, j, c.cur, p_lo, p_c, p_hi, srow<T>(smem, j & 1, rowe), om,
                                                            a.k.ell);
    if (MODE == CG_PQ) {
      sto(c, a.p_out, c.cur, p_c);
      if (c.st) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) red[0] += (double)p_c.v[v] * (double)q.v[v];
      }
    } else if (MODE == CG_UPDATE) {
      if (c.st) {  // x and r are updated in place: only the owning thread touches them
        Vec<T> x = cx, r = cr;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          x.v[v] += coef * p_c.v[v];
          r.v[v] -= coef * q.v[v];
        }
        stw(a.x, c.cur, x);
        stw(a.r, c.cur, r);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) red[0] += (double)r.v[v] * (double)r.v[v];
        if (j + 1 < c.z1) {
          cx = ldw(a.x, c.cur + c.s0b);
          cr = ldw(a.r, c.cur + c.s0b);
        }
      }
    } else if (MODE == CG_RESID) {
      const Vec<T> b = ldo(a.r_in, c.cur);
      Vec<T> y;
#pragma unroll
      for (int v = 0; v < V; ++v) y.v[v] = b.v[v] - q.v[v];
      sto(c, a.p_out, c.cur, y);
    } else if (MODE == CG_FIRST) {
      if (c.st) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) {
            red[0] += (double)p_c.v[v] * (double)q.v[v];
            red[NRED - 1] += (double)q.v[v] * (double)q.v[v];
          }
      }
    } else {
      sto(c, a.p_out, c.cur, q);
    }
    p_lo = p_c;
    p_c = p_hi;
    c.cur += c.s0b;
    ++j;
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if (c.edge || ZEDGE(a))
      post_t<true>(a, smem);
    else
      post_t<false>(a, smem);
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K7  one whole CG iteration in ONE pass (fixed iteration count, tol == 0):
//         r  = r_old - alpha_old (A p_old)        x += alpha_old p_old
//         p  = r + beta p_old                     q  = A p
//         reduce  r.r,  p.q,  r.q,  q.q
//     Same recurrence as K5/K6 (scipy cg, cg.py:50-56), reordered so that x, r and p stream
//     through the SM once per iteration (6 instead of 8 words per voxel): the update that closes
//     iteration i-1 and the direction/matvec that open iteration i share one kernel.  beta needs
//     rho = r.r BEFORE r exists, so the previous kernel predicts it from its own sums,
//         rho_next = ||r - alpha q||^2 = rho - 2 alpha (r.q) + alpha^2 (q.q),
//     while alpha always uses the directly summed rho (finalize_fin: FIN_CGFIRST / FIN_CGFUSED).
//     p (and r) are needed on a +-1 halo of the new p, hence on +-2 of the old one: two halo rows
//     each side in y, one halo vector each side in x, two extra planes at both chunk ends.
//     Shared rows: [slot] = p_old(j), [2 + slot] = p_new(j).
// ==========================================================================================
template <typename T, int D, bool HET>
struct KCgFused {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 2, RH = 2, NRED = 4, NROWS = 4;
  static constexpr int PF2D = 3, PF3D = 2;
  static constexpr int ZC2D = 48;  // planes per CTA: 6 extra planes are read per chunk (0.92 vs 0.90 at 32)
  typedef Col<T, D, LH, RH, 1> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* p_in;   // p_old (== r_in for the first fused iteration: p_0 = r_0)
    const T* r_in;   // r_old
    T* p_out;
    T* r_out;
    T* x;            // in place
    DRef alpha;      // alpha_old
    DRef beta;
    const tvdk_stat_t* skip;  // done / cg_stop
    Fin fin;
  };
  C c;
  int j, rowe;
  T alpha, beta;
  bool same, have_a;
  Vec<T> po_lo, po_c, po_hi, ro_c;  // p_old(j-1), p_old(j), p_old(j+1), r_old(j)
  Vec<T> pn_lo, pn_c, rn_c;         // p_new(j-2), p_new(j-1), r_new(j-1)
  Vec<T> nx_p, nx_r, cx;            // p_old(j+1), r_old(j) loaded one step ahead; x(j-1)
  double red[NRED];

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && (a.skip[lc.bz].done || a.skip[lc.bz].cg_stop)) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
#pragma unroll
    for (int r = 0; r < NRED; ++r) red[r] = 0.0;
    alpha = (T)a.alpha.at(c.b);
    beta = (T)a.beta.at(c.b);
    same = a.p_in == a.r_in;
    j = c.z0 - 1;
    const long long o = c.at(a.g, j);
    c.cur = o;  // the march starts one plane below the chunk
    if (c.plane_ok(a.g, j)) {
      nx_r = ldo(a.r_in, o);
      po_c = same ? nx_r : ldo(a.p_in, o);
      po_lo = c.has_zlo(a.g, j) ? ldo(a.p_in, o - c.s0b) : po_c;
      nx_p = ldo(a.p_in, o + c.s0b);  // plane z0 always exists
    } else {
      nx_p = ldo(a.p_in, o + c.s0b);
      po_c = nx_p;
      po_lo = nx_p;
      nx_r = nx_p;
    }
    pn_lo = pn_c = rn_c = po_c;
    cx = po_c;
    if (c.st) cx = ldw(a.x, o + c.s0b);
    return true;
  }
  TVDK_DEV bool more() const { return j <= c.z1; }

  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    have_a = c.plane_ok(a.g, j);
    po_hi = c.has_zhi(a.g, j) ? nx_p : po_c;
    ro_c = nx_r;
    sst(srow<T>(smem, j & 1, rowe), c, po_c);
    if (j + 1 <= c.z1 && c.plane_ok(a.g, j + 1)) {
      // (first fused iteration: p_0 = r_0, the plane loaded as p one step ago is this step's r)
      nx_r = same ? nx_p : ldo(a.r_in, c.cur + c.s0b);
      if (c.has_zhi(a.g, j + 1)) nx_p = ldo(a.p_in, c.cur + 2 * c.s0b);
    }
    if (c.pf_ok(j)) {
      const long long o = c.cur + c.pfb;
      pfo(a.r_in, o);
      if (!same) pfo(a.p_in, o);
      if (c.st) pfo(a.x, o);
    }
  }

  TVDK_DEV bool ZEDGE(const Args& a) const { return c.zedge(a.g, j) || c.zedge(a.g, j - 1); }
  template <bool E>
  TVDK_DEV void post_t(const Args& a, unsigned char* smem) {
    Vec<T> pn_hi = pn_c;
    Vec<T> rn_hi = rn_c;
    if (have_a) {  // stage A, plane j: close the previous iteration, open this one
      const Vec<T> om = ldcoef<HET>(c, a.k.omega, c.cur);
      const Vec<T> qo = apply_A<T, D, E, HET>(c, a.g, j, c.cur, po_lo, po_c, po_hi, srow<T>(smem, j & 1, rowe), om, a.k.ell);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        rn_hi.v[v] = ro_c.v[v] - alpha * qo.v[v];
        pn_hi.v[v] = beta * po_c.v[v] + rn_hi.v[v];
      }
      sst(srow<T>(smem, 2 + (j & 1), rowe), c, pn_hi);
    }
    if (j > c.z0) {  // stage B, plane j-1: q = A p_new, sums, outputs
      const int zo = j - 1;
      const long long oo = c.cur - c.s0b;
      const Vec<T> om = ldcoef<HET>(c, a.k.omega, oo);
      const Vec<T> q = apply_A<T, D, E, HET>(c, a.g, zo, oo, pn_lo, pn_c, pn_hi, srow<T>(smem, 2 + (zo & 1), rowe), om, a.k.ell);
      if (c.st) {
        Vec<T> x = cx;
#pragma unroll
        for (int v = 0; v < V; ++v) x.v[v] += alpha * po_lo.v[v];  // po_lo still holds p_old(j-1)
        stw(a.x, oo, x);
        stw(a.r_out, oo, rn_c);
        stw(a.p_out, oo, pn_c);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) {
            const double r = (double)rn_c.v[v], pp = (double)pn_c.v[v], qq = (double)q.v[v];
            red[0] += r * r;
            red[1] += pp * qq;
            red[2] += r * qq;
            red[3] += qq * qq;
          }
        if (j < c.z1) cx = ldw(a.x, c.cur);
      }
    }
    po_lo = po_c;
    po_c = po_hi;
    pn_lo = pn_c;
    pn_c = pn_hi;
    rn_c = rn_hi;
    c.cur += c.s0b;
    ++j;
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if (c.edge || ZEDGE(a))
      post_t<true>(a, smem);
    else
      post_t<false>(a, smem);
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K3  shrinkage + Bregman update on the compressed state (split_bregman_tvd.py:136-146),
//     increment partials (:161), in-place update of the iterate (:170).
// ==========================================================================================
template <typename T, int D, bool ISO, bool HET>
struct KShrink {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 0, RH = 1, NRED = 1, NROWS = 2;
  static constexpr int PF2D = 1, PF3D = 3;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* unew[3];  // candidate buffers holding u+ ; selected by stats[b].jac_sel if use_sel
    int use_sel;
    const T* axpy_p;   // optional: u+ = unew[0] + alpha * axpy_p (last x-update of the single-pass CG route)
    DRef axpy_alpha;
    int cg_mode;       // the inner solver was CG: a zero right-hand side makes the iterate ZERO (scipy cg returns b)
    const T* u_old;    // pointwise
    T* u_out;          // may alias u_old (pointwise) or be nullptr
    const T* t_in;     // nullptr: explicit warm-start b in xb
    T* t_out;          // may alias t_in (pointwise)
    long long tcs;
    const T* xb;
    long long xcs;
    const tvdk_stat_t* skip;
    Fin fin;
  };
  C c;
  int j, rowe;
  const T* un;
  const T* ap;
  T aalpha;
  bool zero_up;
  Vec<T> up_c, up_hi, nx_up;
  TVDK_DEV Vec<T> load_up(long long o) const {
    if (zero_up) return vsplat<T>((T)0);
    Vec<T> x = ldo(un, o);
    if (ap) {
      const Vec<T> pv = ldo(ap, o);
#pragma unroll
      for (int v = 0; v < V; ++v) x.v[v] += aalpha * pv.v[v];
    }
    return x;
  }
  // t and u may be updated in place (t_out == t_in, u_out == u_old): only the owning (interior)
  // thread touches them, one plane ahead of their use; halo threads merely publish u+ rows
  RawT<T, D> nx;
  Vec<T> nx_uo;
  double red[NRED];

  TVDK_DEV void prefetch_own(const Args& a, long long o) {
    if (c.st) {
      load_rawt<false, HET, true>(nx, c, a.k, a.t_in, a.tcs, a.xb, a.xb, a.xcs, o);
      nx_uo = ldw(a.u_old, o);
    }
  }
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = 0.0;
    un = a.unew[a.use_sel ? a.skip[lc.bz].jac_sel : 0];
    ap = (a.axpy_p && !(a.skip && a.skip[lc.bz].cg_stop)) ? a.axpy_p : nullptr;
    zero_up = a.cg_mode && a.skip && a.skip[lc.bz].bnrm == 0.0;
    aalpha = ap ? (T)a.axpy_alpha.at(lc.bz) : (T)0;
    j = c.z0;
    const long long o = c.at(a.g, j);
    up_c = load_up(o);
    if (c.has_zhi(a.g, j)) nx_up = load_up(o + c.s0b);
    prefetch_own(a, o);
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    up_hi = c.has_zhi(a.g, j) ? nx_up : up_c;
    sst(srow<T>(smem, j & 1, rowe), c, up_c);
    if (j + 1 < c.z1 && c.has_zhi(a.g, j + 1)) nx_up = load_up(c.cur + 2 * c.s0b);
    if (c.pf_ok(j)) {
      const long long o = c.cur + c.pfb;
      pfo(un, o);
      if (ap) pfo(ap, o);
      if (c.st) {
        pfo(a.u_old, o);
        if (a.t_in) {
#pragma unroll
          for (int d = 0; d < D; ++d) pfo(a.t_in + d * a.tcs, o);
        }
      }
    }
  }
  TVDK_DEV bool ZEDGE(const Args& a) const { return c.zedge(a.g, j); }
  template <bool E>
  TVDK_DEV void post_t(const Args& a, unsigned char* smem) {
    const T* prow = srow<T>(smem, j & 1, rowe);
    const bool expl = a.t_in == nullptr;
    Vec<T> g[D];
    const bool zhi = c.template ZHI<E>(a.g, j);
#pragma unroll
    for (int v = 0; v < V; ++v) g[0].v[v] = zhi ? up_hi.v[v] - up_c.v[v] : (T)0;
    if (D == 3) {
      const Vec<T> yp = syp(prow, c);
#pragma unroll
      for (int v = 0; v < V; ++v) g[1].v[v] = c.template YHI<E>() ? yp.v[v] - up_c.v[v] : (T)0;
    }
    {
      const T xr = sxp(prow, c);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const T right = v < V - 1 ? up_c.v[v < V - 1 ? v + 1 : 0] : xr;
        g[D - 1].v[v] = c.template HR<E>(v) ? right - up_c.v[v] : (T)0;
      }
    }
    if (c.st) {
      Vec<T> bb[D], ww[D], tn[D];
      bw_plane<T, D, ISO, HET>(nx, a.k.sc, expl, bb, ww);
#pragma unroll
      for (int d = 0; d < D; ++d)
#pragma unroll
        for (int v = 0; v < V; ++v) tn[d].v[v] = g[d].v[v] + bb[d].v[v];
#pragma unroll
      for (int d = 0; d < D; ++d) stw(a.t_out + d * a.tcs, c.cur, tn[d]);
      if (a.u_out) stw(a.u_out, c.cur, up_c);
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (c.template OK<E>(v)) {
          const double df = (double)up_c.v[v] - (double)nx_uo.v[v];
          red[0] += df * df;
        }
      if (j + 1 < c.z1) prefetch_own(a, c.cur + c.s0b);
    }
    up_c = up_hi;
    c.cur += c.s0b;
    ++j;
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if (c.edge || ZEDGE(a))
      post_t<true>(a, smem);
    else
      post_t<false>(a, smem);
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// Energy functional (verbose):  1/2 ||omega (x - f)||^2 + sum_a sum |mu G_a x|  (:71-77)
// ==========================================================================================
template <typename T, int D>
struct KEnergy {
  static constexpr int V = Vec<T>::N;
  static constexpr bool HET = true;
  static constexpr int LH = 0, RH = 1, NRED = 2, NROWS = 2;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* x;
    const T* f;
    const tvdk_stat_t* skip;
    Fin fin;
  };
  C c;
  int j, rowe;
  Vec<T> x_c, x_hi;
  double red[NRED];
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = red[1] = 0.0;
    j = c.z0;
    x_c = ldo(a.x, c.cur);
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    x_hi = c.has_zhi(a.g, j) ? ldo(a.x, c.cur + c.s0b) : x_c;
    sst(srow<T>(smem, j & 1, rowe), c, x_c);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const T* prow = srow<T>(smem, j & 1, rowe);
    const Vec<T> f = ldo(a.f, c.cur);
    const Vec<T> mu = ldcoef<HET>(c, a.k.mu, c.cur), om = ldcoef<HET>(c, a.k.omega, c.cur);
    const bool zhi = c.has_zhi(a.g, j);
    Vec<T> yp = x_c;
    if (D == 3) yp = syp(prow, c);
    const T xr = sxp(prow, c);
    if (c.st) {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (c.okc[v]) {
          const T e = om.v[v] * (x_c.v[v] - f.v[v]);
          red[0] += (double)e * (double)e;
          T tv = tabs(mu.v[v] * (zhi ? x_hi.v[v] - x_c.v[v] : (T)0));
          if (D == 3) tv += tabs(mu.v[v] * (c.hyhi ? yp.v[v] - x_c.v[v] : (T)0));
          const T right = v < V - 1 ? x_c.v[v < V - 1 ? v + 1 : 0] : xr;
          tv += tabs(mu.v[v] * (c.hr[v] ? right - x_c.v[v] : (T)0));
          red[1] += (double)tv;
        }
    }
    x_c = x_hi;
    c.cur += c.s0b;
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// Adaptive regularisation (split_bregman_tvd.py:149-158):
//     ell = 1 / max( sum_a |G_a u+| , 1e-12 )         (1-norm of the gradient per voxel)
// ==========================================================================================
template <typename T, int D>
struct KAdapt {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 0, RH = 1, NRED = 0, NROWS = 2;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    const T* x;
    T* ell;
    const tvdk_stat_t* skip;
    Fin fin;
  };
  C c;
  int j, rowe;
  Vec<T> x_c, x_hi;
  double red[1];
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    j = c.z0;
    x_c = ldo(a.x, c.cur);
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    x_hi = c.has_zhi(a.g, j) ? ldo(a.x, c.cur + c.s0b) : x_c;
    sst(srow<T>(smem, j & 1, rowe), c, x_c);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const T* prow = srow<T>(smem, j & 1, rowe);
    const bool zhi = c.has_zhi(a.g, j);
    Vec<T> yp = x_c;
    if (D == 3) yp = syp(prow, c);
    const T xr = sxp(prow, c);
    Vec<T> el;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T g1 = tabs(zhi ? x_hi.v[v] - x_c.v[v] : (T)0);
      if (D == 3) g1 += tabs(c.hyhi ? yp.v[v] - x_c.v[v] : (T)0);
      const T right = v < V - 1 ? x_c.v[v < V - 1 ? v + 1 : 0] : xr;
      g1 += tabs(c.hr[v] ? right - x_c.v[v] : (T)0);
      el.v[v] = (T)1 / tmax(g1, (T)1e-12);
    }
    sto(c, a.ell, c.cur, el);
    x_c = x_hi;
    c.cur += c.s0b;
    ++j;
  }
  static TVDK_DEV void finalize(const Args&, int, const double*) {}
};

// ==========================================================================================
// Streaming kernels: dot products / norms, state expansion, select-copy, pack / unpack
// ==========================================================================================
// red[0] = sum x*y (MODE 0) or sum (x-y)^2 (MODE 1) over the valid voxels of problem b
template <typename T, int D, int MODE>
struct KDot {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 0, RH = 0, NRED = 1, NROWS = 0;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    const T* x;
    const T* y;
    const tvdk_stat_t* skip;
    Fin fin;
  };
  C c;
  double red[NRED];
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    c.setup(a.g, lc);
    red[0] = 0.0;
    if (c.st) {
      long long o = c.cur;
      for (int z = c.z0; z < c.z1; ++z, o += c.s0b) {
        const Vec<T> x = ldo(a.x, o), y = ldo(a.y, o);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.okc[v]) {
            if (MODE == 0)
              red[0] += (double)x.v[v] * (double)y.v[v];
            else {
              const double df = (double)x.v[v] - (double)y.v[v];
              red[0] += df * df;
            }
          }
      }
    }
    return true;
  }
  TVDK_DEV bool more() const { return false; }
  TVDK_DEV void pre(const Args&, unsigned char*) {}
  TVDK_DEV void post(const Args&, unsigned char*) {}
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// d, b from the compressed state (for callers that want the reference's variables back)
template <typename T, int D, bool ISO>
struct KExpand {
  static constexpr int V = Vec<T>::N;
  static constexpr bool HET = true;
  static constexpr int LH = 0, RH = 0, NRED = 0, NROWS = 0;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* t;
    long long tcs;
    T* d;
    T* b;
    long long ocs;
    Fin fin;
  };
  C c;
  double red[1];
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    c.setup(a.g, lc);
    if (c.st) {
      long long o = c.cur;
      for (int z = c.z0; z < c.z1; ++z, o += c.s0b) {
        RawT<T, D> r;
        load_rawt<true, HET, false>(r, c, a.k, a.t, a.tcs, (const T*)nullptr, (const T*)nullptr, 0, o);
        Vec<T> bb[D], ww[D];
        bw_plane<T, D, ISO, HET>(r, a.k.sc, false, bb, ww);
#pragma unroll
        for (int d = 0; d < D; ++d) {
          Vec<T> dv;
#pragma unroll
          for (int v = 0; v < V; ++v) dv.v[v] = r.t[d].v[v] - bb[d].v[v];  // d = t - b
          stw(a.d + d * a.ocs, o, dv);
          stw(a.b + d * a.ocs, o, bb[d]);
        }
      }
    }
    return true;
  }
  TVDK_DEV bool more() const { return false; }
  TVDK_DEV void pre(const Args&, unsigned char*) {}
  TVDK_DEV void post(const Args&, unsigned char*) {}
  static TVDK_DEV void finalize(const Args&, int, const double*) {}
};

// copy src -> dst for the problems whose current iterate lives in buffer `want`
template <typename T, int D>
struct KSelectCopy {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 0, RH = 0, NRED = 0, NROWS = 0;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    const T* src;
    T* dst;
    const tvdk_stat_t* stats;
    int want;
    Fin fin;
  };
  C c;
  double red[1];
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.stats && a.stats[lc.bz].cur_buf != a.want) return false;
    c.setup(a.g, lc);
    if (c.st) {
      long long o = c.cur;
      for (int z = c.z0; z < c.z1; ++z, o += c.s0b) stw(a.dst, o, ldo(a.src, o));
    }
    return true;
  }
  TVDK_DEV bool more() const { return false; }
  TVDK_DEV void pre(const Args&, unsigned char*) {}
  TVDK_DEV void post(const Args&, unsigned char*) {}
  static TVDK_DEV void finalize(const Args&, int, const double*) {}
};

// ==========================================================================================
// Chambolle's projection algorithm -- what darsia.tvd(img) runs by default through
// skimage.restoration.denoise_tv_chambolle (restoration/tvd.py:77-83; third party, PARITY UNPINNED,
// see oracle/chambolle_oracle.py).  One launch = one iteration, every axis spatial:
//     d = div p ; out = f + d ; g = grad out ; n = |g| ; p <- (p - tau g) / (1 + tau/weight n)
//     sums: d.d and sum n (energy).   out(j) is formed for plane j, the update of plane j-1 follows
//     one step later (it needs out(j)); out is needed on a +1 halo, p on -1 of that.
//   MODE 0: the iteration (p_in -> p_out);  MODE 1: the result  out = f + div p  (p chosen per problem
//           by stats[b].cur_buf: the p the last executed iteration read)
//     Shared rows: [slot*(D-1) + a-1] = p_a(j) (a >= 1), [2(D-1) + slot] = out ring.
// ==========================================================================================
template <typename T, int D, int MODE>
struct KChamb {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 1, NRED = MODE == 0 ? 2 : 0, NROWS = 2 * D;
  static constexpr int PF2D = 3, PF3D = 3;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    const T* f;
    const T* p_in;   // MODE 0: D components, stride pcs ; MODE 1: buffer 0
    const T* p_in1;  // MODE 1: buffer 1
    T* p_out;        // MODE 0
    T* out;          // MODE 1
    long long pcs;
    T tau, tau_over_w;
    const tvdk_stat_t* skip;  // MODE 0: done ; MODE 1: cur_buf selects the p buffer
    Fin fin;
  };
  C c;
  int j, rowe;
  const T* pin;
  Vec<T> pc[D], pprev[D], p0_lo, out_prev, f_c;
  Vec<T> nx_p[D], nx_f;
  double red[2];

  TVDK_DEV void load_plane(const Args& a, long long o) {
#pragma unroll
    for (int d = 0; d < D; ++d) nx_p[d] = ldo(pin + d * a.pcs, o);
    nx_f = ldo(a.f, o);
  }
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (MODE == 0 && a.skip && a.skip[lc.bz].done) return false;
    pin = (MODE == 1 && a.skip && a.skip[lc.bz].cur_buf == 1) ? a.p_in1 : a.p_in;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = red[1] = 0.0;
    j = c.z0;
    const long long o = c.at(a.g, j);
    p0_lo = vsplat<T>((T)0);
    if (j > 0) p0_lo = ldo(pin, o - c.s0b);
    out_prev = p0_lo;
#pragma unroll
    for (int d = 0; d < D; ++d) pprev[d] = p0_lo;
    load_plane(a, o);
    return true;
  }
  // the iteration marches one plane past the chunk (out(z1) closes the gradient of plane z1 - 1)
  TVDK_DEV bool more() const { return MODE == 0 ? j <= c.z1 : j < c.z1; }

  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    if (j >= a.g.n0) return;
    const int slot = j & 1;
#pragma unroll
    for (int d = 0; d < D; ++d) pc[d] = nx_p[d];
    f_c = nx_f;
#pragma unroll
    for (int d = 1; d < D; ++d) sst(srow<T>(smem, slot * (D - 1) + d - 1, rowe), c, pc[d]);
    const int last = MODE == 0 ? c.z1 : c.z1 - 1;
    if (j + 1 <= last && j + 1 < a.g.n0) load_plane(a, c.cur + c.s0b);
    if (c.pf_ok(j)) {
      const long long o = c.cur + c.pfb;
      pfo(a.f, o);
#pragma unroll
      for (int d = 0; d < D; ++d) pfo(pin + d * a.pcs, o);
    }
  }

  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const int slot = j & 1;
    Vec<T> out_c = out_prev;
    if (j < a.g.n0) {
      // d = sum_a ( p_a[k - e_a] (k_a >= 1) - p_a[k] )
      Vec<T> d;
#pragma unroll
      for (int v = 0; v < V; ++v) d.v[v] = -pc[0].v[v];
      if (D == 3) {
#pragma unroll
        for (int v = 0; v < V; ++v) d.v[v] = d.v[v] - pc[1].v[v];
      }
#pragma unroll
      for (int v = 0; v < V; ++v) d.v[v] = d.v[v] - pc[D - 1].v[v];
      if (j > 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) d.v[v] += p0_lo.v[v];
      }
      if (D == 3 && c.hylo) {
        const Vec<T> ym = sym(srow<T>(smem, slot * (D - 1) + 0, rowe), c);
#pragma unroll
        for (int v = 0; v < V; ++v) d.v[v] += ym.v[v];
      }
      {
        const T xl = sxm(srow<T>(smem, slot * (D - 1) + (D - 2), rowe), c);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const T left = v > 0 ? pc[D - 1].v[v > 0 ? v - 1 : 0] : xl;
          if (c.hl[v]) d.v[v] += left;
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) out_c.v[v] = f_c.v[v] + d.v[v];
      if (MODE == 1) {
        sto(c, a.out, c.cur, out_c);
      } else {
        sst(srow<T>(smem, 2 * (D - 1) + slot, rowe), c, out_c);
        if (c.st && j < c.z1) {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (c.okc[v]) red[0] += (double)d.v[v] * (double)d.v[v];
        }
      }
    }
    if (MODE == 0 && j > c.z0) {
      // gradient of out at plane j-1, norm, projection step for p(j-1)
      const int zo = j - 1;
      const T* orow = srow<T>(smem, 2 * (D - 1) + (zo & 1), rowe);
      Vec<T> g[D];
      const bool zhi = zo + 1 < a.g.n0;
#pragma unroll
      for (int v = 0; v < V; ++v) g[0].v[v] = zhi ? out_c.v[v] - out_prev.v[v] : (T)0;
      if (D == 3) {
        const Vec<T> yp = syp(orow, c);
#pragma unroll
        for (int v = 0; v < V; ++v) g[1].v[v] = c.hyhi ? yp.v[v] - out_prev.v[v] : (T)0;
      }
      {
        const T xr = sxp(orow, c);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const T right = v < V - 1 ? out_prev.v[v < V - 1 ? v + 1 : 0] : xr;
          g[D - 1].v[v] = c.hr[v] ? right - out_prev.v[v] : (T)0;
        }
      }
      Vec<T> pn[D];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        T s2 = g[0].v[v] * g[0].v[v];
#pragma unroll
        for (int d = 1; d < D; ++d) s2 = s2 + g[d].v[v] * g[d].v[v];
        const T nrm = tsqrt(s2);
        const T den = nrm * a.tau_over_w + (T)1;
#pragma unroll
        for (int d = 0; d < D; ++d) pn[d].v[v] = tdiv(pprev[d].v[v] - a.tau * g[d].v[v], den);
        if (c.st && c.okc[v]) red[1] += (double)nrm;
      }
      const long long oo = c.cur - c.s0b;
#pragma unroll
      for (int d = 0; d < D; ++d) sto(c, a.p_out + d * a.pcs, oo, pn[d]);
    }
    p0_lo = pc[0];
#pragma unroll
    for (int d = 0; d < D; ++d) pprev[d] = pc[d];
    out_prev = out_c;
    c.cur += c.s0b;
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED > 0 ? NRED : 1, b, tot); }
};

// Multigrid grid transfer (utils/linear_solvers/mg.py).  One thread per element of the TARGET grid.
//   MODE 0  restriction (mg.py:107-132): pairwise averages along axis 0, then 1 (, then 2); the last
//           entry of an odd extent is dropped (coarse extent = fine extent / 2)
//   MODE 1  prolongation + edge padding + correction (mg.py:226-239): every coarse value is repeated
//           twice along each axis, an odd fine extent re-uses the last coarse value; fine += that
template <typename T, int D, int MODE>
struct KMgXfer {
  static constexpr int NRED = 0, NROWS = 0;
  struct Args {
    Geom g;   // target grid (coarse for MODE 0, fine for MODE 1)
    Geom go;  // the other grid
    const T* src;
    T* dst;
    Fin fin;
  };
  double red[1];
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    // grid: x = rows (z * ny + y) of the target, y = row segments, z = problem
    const long long row = lc.bx;
    const int z = (int)(row / a.g.ny), y = (int)(row % a.g.ny);
    const T* sb = a.src + (long long)lc.bz * a.go.sb;
    T* db = a.dst + (long long)lc.bz * a.g.sb;
    for (int x = lc.by * lc.ntx + lc.tx; x < a.g.nx; x += lc.nby * lc.ntx) {
      const long long ti = (long long)z * a.g.s0 + (long long)y * a.g.pitch + x;
      if (MODE == 0) {
        T acc[2];
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          T ay[2];
#pragma unroll
          for (int dy = 0; dy < (D == 3 ? 2 : 1); ++dy) {
            const long long o = (long long)(2 * z) * a.go.s0 + (long long)(D == 3 ? 2 * y + dy : 0) * a.go.pitch + 2 * x + dx;
            ay[dy] = (sb[o] + sb[o + a.go.s0]) / (T)2;  // axis 0 first
          }
          acc[dx] = D == 3 ? (ay[0] + ay[1]) / (T)2 : ay[0];  // then axis 1 (3-D)
        }
        db[ti] = (acc[0] + acc[1]) / (T)2;  // last axis
      } else {
        int zc = z >> 1, yc = y >> 1, xc = x >> 1;
        zc = zc < a.go.n0 ? zc : a.go.n0 - 1;
        yc = D == 3 ? (yc < a.go.ny ? yc : a.go.ny - 1) : 0;
        xc = xc < a.go.nx ? xc : a.go.nx - 1;
        db[ti] += sb[(long long)zc * a.go.s0 + (long long)yc * a.go.pitch + xc];
      }
    }
    return true;
  }
  TVDK_DEV bool more() const { return false; }
  TVDK_DEV void pre(const Args&, unsigned char*) {}
  TVDK_DEV void post(const Args&, unsigned char*) {}
  static TVDK_DEV void finalize(const Args&, int, const double*) {}
};

// host-order (voxel-major, problems interleaved) <-> planar padded fields, with dtype cast.
// One thread per (row, x, b) element of the interleaved side: contiguous accesses there.
template <typename TI, typename TF, int PACK>
struct KPack {
  static constexpr int NRED = 0, NROWS = 0;
  struct Args {
    Geom g;
    const void* src;
    void* dst;
    Fin fin;
  };
  double red[1];
  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    // grid: x = rows (z*ny + y) of the local planes, y = row segments ; threads: x*nb + b
    const long long row = lc.bx;
    const long long per_row = (long long)a.g.nx * a.g.nb;
    const long long z = row / a.g.ny, y = row % a.g.ny;
    for (long long e = (long long)lc.by * lc.ntx + lc.tx; e < per_row; e += (long long)lc.nby * lc.ntx) {
      const long long x = e / a.g.nb, b = e % a.g.nb;
      const long long fi = b * a.g.sb + z * a.g.s0 + y * a.g.pitch + x;
      const long long ii = row * per_row + e;
      if (PACK)
        reinterpret_cast<TF*>(a.dst)[fi] = (TF) reinterpret_cast<const TI*>(a.src)[ii];
      else
        reinterpret_cast<TI*>(a.dst)[ii] = (TI) reinterpret_cast<const TF*>(a.src)[fi];
    }
    return true;
  }
  TVDK_DEV bool more() const { return false; }
  TVDK_DEV void pre(const Args&, unsigned char*) {}
  TVDK_DEV void post(const Args&, unsigned char*) {}
  static TVDK_DEV void finalize(const Args&, int, const double*) {}
};

// ==========================================================================================
// TMA staging (sm_100a: cp.async.bulk.tensor + mbarrier; tests/emu: a synchronous copy with zero fill)
//   A TmaMap describes one field as a 3-D tensor (x: pitch, y: ny, z: nb * n0 planes) together with the box a
//   CTA fetches per plane.  Out-of-range box elements are ZERO-filled by the hardware; every boundary rule
//   below selects around them exactly like the march kernels select around their clamped loads.
// ==========================================================================================
#if defined(__CUDACC__)
struct alignas(64) TmaMap {
  unsigned long long opaque[16];  // CUtensorMap (filled by cuTensorMapEncodeTiled on the host)
};
TVDK_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
TVDK_DEV void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
TVDK_DEV void mbar_init_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
TVDK_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
TVDK_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
TVDK_DEV void tma_load_3d(void* dst, const TmaMap* map, int x, int y, int z, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_addr(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_addr(bar))
      : "memory");
}
#else
struct TmaMap {  // emulator: plain description of the tensor and the box
  const unsigned char* base;
  long long dim[3];       // elements: x, y, z
  long long stride_b[2];  // bytes between rows (y) and planes (z)
  int box[3];
  int esize;
};
TVDK_DEV void mbar_init(unsigned long long*, int) {}
TVDK_DEV void mbar_init_fence() {}
TVDK_DEV void mbar_expect_tx(unsigned long long*, unsigned) {}
TVDK_DEV void mbar_wait(unsigned long long*, unsigned) {}
TVDK_DEV void tma_load_3d(void* dst, const TmaMap* m, int x, int y, int z, unsigned long long*) {
  unsigned char* d = reinterpret_cast<unsigned char*>(dst);
  for (int k = 0; k < m->box[2]; ++k)
    for (int j = 0; j < m->box[1]; ++j)
      for (int i = 0; i < m->box[0]; ++i) {
        const long long X = x + i, Y = y + j, Z = z + k;
        unsigned char* q = d + ((long long)(k * m->box[1] + j) * m->box[0] + i) * m->esize;
        const bool in = X >= 0 && X < m->dim[0] && Y >= 0 && Y < m->dim[1] && Z >= 0 && Z < m->dim[2];
        for (int e = 0; e < m->esize; ++e) q[e] = in ? m->base[Z * m->stride_b[1] + Y * m->stride_b[0] + X * m->esize + e] : 0;
      }
}
#endif

// ==========================================================================================
// K1-TMA  fused outer iteration for Jacobi(maxiter=1), 3-D, scalar coefficients
//     (BASELINE config 5; same arithmetic as K1: split_bregman_tvd.py:84-89 rhs, jacobi.py:101-114 sweep,
//      :136-146 shrink + Bregman update, :161 increment)
//   All inputs of a plane arrive as TMA boxes (tile + halo) in a ring of NS shared-memory stages, issued by
//   one thread NS - 1 planes ahead: no halo threads for u or t, no register prefetch, no per-thread load
//   addresses.  The only halo threads left are the one row / one vector on the high side where u+ is recomputed.
//   What the divergence needs of w_a = ell (b_a - d_a) at y - 1 and x - 1 is re-derived from the staged t:
//     anisotropic: w_a is a pointwise function of t_a alone (:80-81, :142-146) -- recomputed on the spot;
//     isotropic:   w_a = ell (t_a (1 - phi) - t_a phi) with the shrinkage factor phi of :117-126, which costs a
//                  square root and a division: phi of plane j + 1 is formed ONE STEP AHEAD by the threads (own
//                  voxels, plus the row y0 - 1 and the column x0 - 1 dealt out one voxel per thread) and published
//                  in a shared ring, so the step keeps its single barrier.
//   Per step j (one barrier):  A: [iso: phi(j+1) -> phi ring]  u+(j) from stage(j) (+ u(j+1) from stage(j+1)) -> ring[j & 1]
//                              B: t+(j-1) from ring[(j-1) & 1] and registers, stores of plane j - 1
//                              barrier ; one thread refills stage(j) with plane j + NS.
//   Stage layout: U (ntx+2)V x (nty+2) | F ntx V x nty | aniso: T0 ntx V x nty, T1 ntx V x (nty+1), T2 (ntx+1)V x nty
//                                                      | iso:   T0, T1, T2 (ntx+1)V x (nty+1) each (origin x0 - V, y0 - 1).
// ==========================================================================================
#ifndef TVDK_TMA_NS
#define TVDK_TMA_NS 3
#endif
template <typename T, bool ISO, int DD = 3>
struct KFusedJ1T {
  static_assert(DD == 3 || (DD == 2 && !ISO), "the 2-D variant is anisotropic only");
  static constexpr int D = DD;
  static constexpr int DX = D - 1;  // index of the x component of t (the contiguous axis)
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 0, RH = 1, NRED = 1;
  static constexpr int NS = TVDK_TMA_NS;
  static constexpr int ZC3D = 64;
  // launch shape (profiles/r2_tma_ab_and_tiles.txt, 512^3): small CTAs win -- 16 x 8 threads (fp64: 0.91 of the HBM peak
  // vs 0.82 with 32 x 16; five CTAs of 40 KB per SM instead of one of 148 KB), 16 x 12 for fp32 (0.91 vs 0.86)
  static constexpr int NTX3D = 16;
  static constexpr int NTY3D = sizeof(T) == 4 ? 12 : 8;
  // 2-D: the "plane" of the march is one image row, a CTA is a strip of ntx vectors; a TMA box holds at most 256 elements
  // per axis, so (ntx + 2) V <= 256: at most 96 threads (fp64) / 32 (fp32).  Measured (profiles/r2_tma2d_ab.txt): SLOWER
  // than the march kernel on the 1788 x 3180 image (0.67-0.79 vs 0.80 of the HBM peak; fp32 0.43 vs 0.66), +2.6 % on
  // 4000 x 7000 x 3 where K1x2 is faster still -- a 2-D row is contiguous, so there are no y-halo threads for the boxes to
  // remove.  Hence not a default route (TVDK_TMA2D=1 / fused = 4 select it).
  static constexpr int NTX2D_MAX = ((256 / V - 2) / 32) * 32;
  static constexpr int NTX2D = sizeof(T) == 8 ? 64 : 32;  // (96 measured 5-9 % faster in fp64; 64 is the GPU-tested default)
  static constexpr bool TMA = true;
  typedef Col<T, D, LH, RH, 0, 1> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    T* u_out;
    T* t_out;
    long long tcs_out;
    const tvdk_stat_t* skip;  // problems with skip[b].done are left untouched
    Fin fin;
    TmaMap tm_u, tm_f, tm_t0, tm_t1, tm_t2;
  };
  // shared-memory layout for a thread block of ntx x nty threads (byte offsets; every buffer 128-byte aligned)
  struct Lay {
    int pU, pF, pT0, pT1, pT2;  // row pitches in elements
    int oU, oF, oT0, oT1, oT2;  // byte offsets inside one stage
    int stage_bytes, ring_bytes, ring_off, phi_bytes, phi_off, stage_off, total;
    unsigned tx_bytes;          // bytes one plane's five boxes deliver
  };
  static TVDK_HD int a128(int x) { return (x + 127) / 128 * 128; }
  // box shapes (elements x rows) -- the host encodes the tensor maps with the same numbers
  static TVDK_HD int box_w(int which, int ntx) {  // 0: U, 1: F, 2..4: T0..T2
    return which == 0 ? (ntx + 2) * V : which == 1 ? ntx * V : ISO ? (ntx + 1) * V : (which == 4 ? (ntx + 1) * V : ntx * V);
  }
  static TVDK_HD int box_h(int which, int nty) {
    if (D == 2) return 1;  // (no y axis: every box is one row high; T1 is not fetched at all)
    return which == 0 ? nty + 2 : which == 1 ? nty : ISO ? nty + 1 : (which == 3 ? nty + 1 : nty);
  }
  static TVDK_HD Lay layout(int ntx, int nty) {
    Lay L;
    const int W = (int)sizeof(T);
    L.pU = box_w(0, ntx);
    L.pF = box_w(1, ntx);
    L.pT0 = box_w(2, ntx);
    L.pT1 = box_w(3, ntx);
    L.pT2 = box_w(4, ntx);
    int b[5];
    for (int i = 0; i < 5; ++i) b[i] = box_w(i, ntx) * box_h(i, nty) * W;
    if (D == 2) b[3] = 0;
    L.oU = 0;
    L.oF = L.oU + a128(b[0]);
    L.oT0 = L.oF + a128(b[1]);
    L.oT1 = L.oT0 + a128(b[2]);
    L.oT2 = L.oT1 + a128(b[3]);
    L.stage_bytes = L.oT2 + a128(b[4]);
    L.tx_bytes = (unsigned)(b[0] + b[1] + b[2] + b[3] + b[4]);
    L.ring_bytes = a128((L.pF * (nty + 1) + V) * W);  // one row + one vector of slack for the halo threads' reads
    L.ring_off = 128;                                 // [0, 128): the NS mbarriers
    L.phi_bytes = ISO ? a128(L.pT0 * (nty + 1) * W) : 0;
    L.phi_off = L.ring_off + 2 * L.ring_bytes;
    L.stage_off = L.phi_off + 2 * L.phi_bytes;
    L.total = L.stage_off + NS * L.stage_bytes;
    return L;
  }
  static TVDK_HD size_t smem_bytes(int ntx, int nty) { return (size_t)layout(ntx, nty).total; }

  C c;
  Lay L;
  int j, x0b, y0b, zb;  // march plane; box origin of the tile; plane offset of the problem
  int sr, kq;           // stage of plane j and how many times that stage has been used before it
  int iU, iF, iT0, iT1, iT2;  // element offsets of the thread's vector inside the stage buffers (ring: iF)
  int tid, nthr, txt;   // linear thread index, threads per block, tile width in elements (iso: halo phi dealing)
  Vec<T> u_lo, u_c;     // u(j-1), u(j)
  Vec<T> w0_prev, up_prev, b_prev[D];
  Vec<T> phi_c;         // iso: phi of the own voxels of plane j (formed one step ahead)
  bool leader;
  double red[NRED];

  static TVDK_DEV bool skipped(const Args& a, const LaunchCtx& lc) { return a.skip && a.skip[lc.bz].done; }
  static TVDK_DEV void bar_init(const Args&, const LaunchCtx& lc, unsigned char* smem) {
    if (lc.tx == 0 && lc.ty == 0) {
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
      for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
      mbar_init_fence();
    }
  }
  // ring bookkeeping without divisions: plane j sits in stage `sr` and is that stage's `kq`-th use
  TVDK_DEV unsigned long long* bar_at(unsigned char* smem, int idx) const { return reinterpret_cast<unsigned long long*>(smem) + idx; }
  TVDK_DEV unsigned char* stage_at(unsigned char* smem, int idx) const { return smem + L.stage_off + idx * L.stage_bytes; }
  TVDK_DEV T* ring(unsigned char* smem, int slot) const { return reinterpret_cast<T*>(smem + L.ring_off + slot * L.ring_bytes); }
  TVDK_DEV T* phiring(unsigned char* smem, int slot) const { return reinterpret_cast<T*>(smem + L.phi_off + slot * L.phi_bytes); }
  TVDK_DEV void issue(const Args& a, unsigned char* smem, int p, int idx) const {  // one thread: plane p -> stage idx
    unsigned long long* b = bar_at(smem, idx);
    unsigned char* s = stage_at(smem, idx);
    mbar_expect_tx(b, L.tx_bytes);
    const int z = zb + p;
    if (D == 2) {  // rows of a 2-D image: tensor (x, 1, nb * n0)
      tma_load_3d(s + L.oU, &a.tm_u, x0b - V, 0, z, b);
      tma_load_3d(s + L.oF, &a.tm_f, x0b, 0, z, b);
      tma_load_3d(s + L.oT0, &a.tm_t0, x0b, 0, z, b);
      tma_load_3d(s + L.oT2, &a.tm_t2, x0b - V, 0, z, b);
      return;
    }
    tma_load_3d(s + L.oU, &a.tm_u, x0b - V, y0b - 1, z, b);
    tma_load_3d(s + L.oF, &a.tm_f, x0b, y0b, z, b);
    if (ISO) {
      tma_load_3d(s + L.oT0, &a.tm_t0, x0b - V, y0b - 1, z, b);
      tma_load_3d(s + L.oT1, &a.tm_t1, x0b - V, y0b - 1, z, b);
      tma_load_3d(s + L.oT2, &a.tm_t2, x0b - V, y0b - 1, z, b);
    } else {
      tma_load_3d(s + L.oT0, &a.tm_t0, x0b, y0b, z, b);
      tma_load_3d(s + L.oT1, &a.tm_t1, x0b, y0b - 1, z, b);
      tma_load_3d(s + L.oT2, &a.tm_t2, x0b - V, y0b, z, b);
    }
  }

  // w = ell (b - d), b of one component from its t (anisotropic: split_bregman_tvd.py:80-81, :142-146)
  static TVDK_DEV void bw1(T t, const Scalars<T>& sc, T& b, T& w) {
    T tt[1] = {t}, bb[1], ww[1];
    bw_from_t<T, 1, false>(tt, sc.kappa, sc.ell, bb, ww);
    b = bb[0];
    w = ww[0];
  }
  // ... and of one component given the voxel's isotropic shrinkage factor (:117-126): d = t phi, b = t (1 - phi)
  static TVDK_DEV void bwphi(T t, T phi, const Scalars<T>& sc, T& b, T& w) {
    const T d = t * phi;
    b = t * ((T)1 - phi);
    w = sc.ell * (b - d);
  }
  static TVDK_DEV T phi3(T t0, T t1, T t2, const Scalars<T>& sc) {
    const T tt[3] = {t0, t1, t2};
    return iso_phi_of<T, 3, true>(tt, sc.kappa);
  }
  // iso: phi of plane p (staged in stage idx) -> phi ring slot: own voxels (returned) + the dealt halo voxels
  TVDK_DEV Vec<T> make_phi(const Args& a, unsigned char* smem, int idx, int slot) const {
    const unsigned char* sp = stage_at(smem, idx);
    const T* T0 = reinterpret_cast<const T*>(sp + L.oT0);
    const T* T1 = reinterpret_cast<const T*>(sp + L.oT1);
    const T* T2 = reinterpret_cast<const T*>(sp + L.oT2);
    T* P = phiring(smem, slot);
    const Vec<T> t0 = vload(T0 + iT0), t1 = vload(T1 + iT0), t2 = vload(T2 + iT0);
    Vec<T> ph;
#pragma unroll
    for (int v = 0; v < V; ++v) ph.v[v] = phi3(t0.v[v], t1.v[v], t2.v[v], a.k.sc);
    vstore(P + iT0, ph);
    // row y0 - 1 (box row 0, elements V .. V + txt) and column x0 - 1 (box column V - 1, rows 1 .. nty)
    const int nty = nthr / (txt / V);
    for (int h = tid; h < txt + nty; h += nthr) {
      const int e = h < txt ? V + h : (h - txt + 1) * L.pT0 + V - 1;
      P[e] = phi3(T0[e], T1[e], T2[e], a.k.sc);
    }
    return ph;
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char* smem) {
    c.setup(a.g, lc);
    L = layout(lc.ntx, lc.nty);
    red[0] = 0.0;
    leader = lc.tx == 0 && lc.ty == 0;
    tid = lc.ty * lc.ntx + lc.tx;
    nthr = lc.ntx * lc.nty;
    txt = lc.ntx * V;
    const int xt = lc.bx % a.g.xtiles, yt = lc.bx / a.g.xtiles;
    x0b = xt * C::iw(lc.ntx);
    y0b = D == 3 ? yt * C::ih(lc.nty) : 0;
    zb = c.b * a.g.n0;
    // (no halo threads on the low side: a tile that STARTS at the boundary must take the selecting path too; and
    //  u+ is formed on every thread row / vector, so a tile whose last thread row / vector reaches the last image
    //  row / column needs the edge-replicating selects as well)
    c.edge = c.edge || x0b == 0 || x0b + lc.ntx * V >= a.g.nx || (D == 3 && (y0b == 0 || y0b + lc.nty >= a.g.ny));
    iU = (D == 3 ? (lc.ty + 1) * L.pU : 0) + (lc.tx + 1) * V;
    iF = lc.ty * L.pF + lc.tx * V;
    if (ISO) {
      iT0 = iT1 = iT2 = (lc.ty + 1) * L.pT0 + (lc.tx + 1) * V;
    } else {
      iT0 = iF;
      iT1 = (lc.ty + 1) * L.pT1 + lc.tx * V;
      iT2 = lc.ty * L.pT2 + (lc.tx + 1) * V;
    }
    j = c.z0;
    sr = 1;  // plane z0 - 1 sits in stage 0, plane z0 in stage 1, ...
    kq = 0;
    if (leader) {
      for (int i = 0; i < NS && c.z0 - 1 + i <= c.z1 + 1; ++i) issue(a, smem, c.z0 - 1 + i, i);
    }
    // planes z0 - 1 (u below, w_0 below) and z0
    mbar_wait(bar_at(smem, 0), 0);
    mbar_wait(bar_at(smem, 1), 0);
    const unsigned char* sm1 = stage_at(smem, 0);
    u_c = vload(reinterpret_cast<const T*>(stage_at(smem, 1) + L.oU) + iU);
    u_lo = u_c;
    w0_prev = vsplat<T>((T)0);
    if (c.has_zlo(a.g, j)) {
      u_lo = vload(reinterpret_cast<const T*>(sm1 + L.oU) + iU);
      const Vec<T> t0 = vload(reinterpret_cast<const T*>(sm1 + L.oT0) + iT0);
      if (ISO) {
        const Vec<T> t1 = vload(reinterpret_cast<const T*>(sm1 + L.oT1) + iT0), t2 = vload(reinterpret_cast<const T*>(sm1 + L.oT2) + iT0);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T b;
          bwphi(t0.v[v], phi3(t0.v[v], t1.v[v], t2.v[v], a.k.sc), a.k.sc, b, w0_prev.v[v]);
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T b;
          bw1(t0.v[v], a.k.sc, b, w0_prev.v[v]);
        }
      }
    }
    phi_c = vsplat<T>((T)0);
    if (ISO) phi_c = make_phi(a, smem, 1, 0);  // phi of plane z0 -> slot 0 (published by the barrier of pre_sync)
    up_prev = vsplat<T>((T)0);
#pragma unroll
    for (int d = 0; d < D; ++d) b_prev[d] = vsplat<T>((T)0);
    return true;
  }
  // iso: the phi ring slot filled in init() must be visible to the whole block before the first step reads it
  static constexpr bool SYNC_AFTER_INIT = ISO;
  TVDK_DEV bool more() const { return j <= c.z1; }

  template <bool E>
  TVDK_DEV void step(const Args& a, unsigned char* smem) {
    const Scalars<T>& sc = a.k.sc;
    const int slot = (j - c.z0) & 1;
    const bool wrap = sr + 1 == NS;
    const int sr1 = wrap ? 0 : sr + 1;  // stage of plane j + 1
    mbar_wait(bar_at(smem, sr1), (unsigned)((kq + (wrap ? 1 : 0)) & 1));  // (every issued plane is waited for, also when unused)
    const bool have_row = (j < c.z1) || c.plane_ok(a.g, j);
    Vec<T> up = vsplat<T>((T)0), u_hi = u_c;
    Vec<T> b_cur[D], w0_cur = w0_prev;
#pragma unroll
    for (int d = 0; d < D; ++d) b_cur[d] = b_prev[d];
    Vec<T> phi_n = phi_c;
    if (ISO && j < c.z1) phi_n = make_phi(a, smem, sr1, slot ^ 1);  // phi of plane j + 1, read after the next barrier
    if (have_row) {
      const unsigned char* sj = stage_at(smem, sr);
      const T* U = reinterpret_cast<const T*>(sj + L.oU);
      const T* T1 = reinterpret_cast<const T*>(sj + L.oT1);
      const T* T2 = reinterpret_cast<const T*>(sj + L.oT2);
      if (c.has_zhi(a.g, j)) u_hi = vload(reinterpret_cast<const T*>(stage_at(smem, sr1) + L.oU) + iU);
      const Vec<T> f = vload(reinterpret_cast<const T*>(sj + L.oF) + iF);
      const Vec<T> t0 = vload(reinterpret_cast<const T*>(sj + L.oT0) + iT0);
      Vec<T> t1 = t0, t1m = t0, ym = u_c, yp = u_c;
      if (D == 3) {
        t1 = vload(T1 + iT1);
        t1m = vload(T1 + iT1 - L.pT1);
        ym = vload(U + iU - L.pU);
        yp = vload(U + iU + L.pU);
      }
      const Vec<T> t2 = vload(T2 + iT2);
      const T t2l = T2[iT2 - 1];
      const T xl = U[iU - 1], xr = U[iU + V];
      Vec<T> w1, w2, w1m;
      T w2l, bdummy;
      if constexpr (ISO) {
        const T* P = phiring(smem, slot);
        const Vec<T> phm = vload(P + iT0 - L.pT0);
        bwphi(t2l, P[iT0 - 1], sc, bdummy, w2l);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          bwphi(t0.v[v], phi_c.v[v], sc, b_cur[0].v[v], w0_cur.v[v]);
          bwphi(t1.v[v], phi_c.v[v], sc, b_cur[1].v[v], w1.v[v]);
          bwphi(t2.v[v], phi_c.v[v], sc, b_cur[2].v[v], w2.v[v]);
          bwphi(t1m.v[v], phm.v[v], sc, bdummy, w1m.v[v]);
        }
      } else {
        bw1(t2l, sc, bdummy, w2l);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          bw1(t0.v[v], sc, b_cur[0].v[v], w0_cur.v[v]);
          if (D == 3) {
            bw1(t1.v[v], sc, b_cur[D == 3 ? 1 : 0].v[v], w1.v[v]);
            bw1(t1m.v[v], sc, bdummy, w1m.v[v]);
          }
          bw1(t2.v[v], sc, b_cur[DX].v[v], w2.v[v]);
        }
      }
      // rhs = omega f + H_0 w_0 + H_1 w_1 + H_2 w_2 in the reference's order (:84-89), then the sweep (jacobi.py:112-114)
      const bool zlo = c.template ZLO<E>(a.g, j);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        T rhs = sc.omega * f.v[v];
        if (zlo) rhs += w0_cur.v[v] - w0_prev.v[v];
        if (D == 3 && c.template YLO<E>()) rhs += w1.v[v] - w1m.v[v];
        const T wl = v > 0 ? w2.v[v > 0 ? v - 1 : 0] : w2l;
        if (c.template HL<E>(v)) rhs += w2.v[v] - wl;
        T nb = u_lo.v[v] + u_hi.v[v];
        if (D == 3) {
          nb += c.template YLO<E>() ? ym.v[v] : u_c.v[v];
          nb += c.template YHI<E>() ? yp.v[v] : u_c.v[v];
        }
        const T left = v > 0 ? u_c.v[v > 0 ? v - 1 : 0] : xl;
        const T right = v < V - 1 ? u_c.v[v < V - 1 ? v + 1 : 0] : xr;
        nb += c.template HL<E>(v) ? left : u_c.v[v];
        nb += c.template HR<E>(v) ? right : u_c.v[v];
        up.v[v] = rhs * sc.inv_diag + nb * sc.c_nb;
      }
      vstore(ring(smem, slot) + iF, up);
    }
    if (j > c.z0) {
      // outputs of plane j-1: t_new = G u+ + b, u_out = u+, increment (:136-146, :161)
      const int zo = j - 1;
      const long long oo = c.cur - c.s0b;
      const T* R = ring(smem, slot ^ 1);
      Vec<T> yp = up_prev;
      if (D == 3) yp = vload(R + iF + L.pF);
      const T xr = R[iF + V];
      const bool zhi = c.template ZHI<E>(a.g, zo);
      Vec<T> tn[D];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        tn[0].v[v] = (zhi ? up.v[v] - up_prev.v[v] : (T)0) + b_prev[0].v[v];
        if (D == 3) tn[D == 3 ? 1 : 0].v[v] = (c.template YHI<E>() ? yp.v[v] - up_prev.v[v] : (T)0) + b_prev[D == 3 ? 1 : 0].v[v];
        const T right = v < V - 1 ? up_prev.v[v < V - 1 ? v + 1 : 0] : xr;
        tn[DX].v[v] = (c.template HR<E>(v) ? right - up_prev.v[v] : (T)0) + b_prev[DX].v[v];
      }
      sto(c, a.u_out, oo, up_prev);
#pragma unroll
      for (int d = 0; d < D; ++d) sto(c, a.t_out + d * a.tcs_out, oo, tn[d]);
      if (c.st) {  // u_lo still holds u(j-1)
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) {
            const double df = (double)up_prev.v[v] - (double)u_lo.v[v];
            red[0] += df * df;
          }
      }
    }
    u_lo = u_c;
    u_c = u_hi;
    w0_prev = w0_cur;
    up_prev = up;
    phi_c = phi_n;
#pragma unroll
    for (int d = 0; d < D; ++d) b_prev[d] = b_cur[d];
  }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    if (c.edge || c.zedge(a.g, j) || c.zedge(a.g, j - 1))
      step<true>(a, smem);
    else
      step<false>(a, smem);
  }
  // after the barrier: stage(j) (and, on the first step, the stage of plane z0 - 1) is free again
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if (leader) {
      if (j == c.z0 && j - 1 + NS <= c.z1 + 1) issue(a, smem, j - 1 + NS, 0);
      if (j + NS <= c.z1 + 1) issue(a, smem, j + NS, sr);
    }
    if (++sr == NS) {
      sr = 0;
      ++kq;
    }
    c.cur += c.s0b;
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K5-TMA / K4-TMA  one stencil field staged by TMA, 3-D, scalar coefficients -- no halo threads at all:
//   MODE TS_PQ     p_out = first ? r : beta p + r ;  reduce p_out . (A p_out)        (K5: scipy cg via cg.py:50-56)
//                  p_out of the six neighbours is re-derived from the staged p and r (one FMA each)
//   MODE TS_SWEEP  x_out = rhs / diag + N(x) / (diag / ell)                           (K4: jacobi.py:101-114)
//   Boxes per plane: A = p / x with a one-cell halo (ntx V + 2V) x (nty + 2); B = r with the same halo (PQ) or the
//   bare tile (SWEEP).  The z-neighbours are carried in registers; the step's only barrier recycles the stage.
// ==========================================================================================
enum { TS_PQ = 0, TS_SWEEP = 1 };

template <typename T, int MODE>
struct KStencilT {
  static constexpr int D = 3;
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 0, RH = 0, NRED = MODE == TS_PQ ? 1 : 0;
  static constexpr int NS = TVDK_TMA_NS;
  static constexpr int ZC3D = 64;
  static constexpr int NTX3D = 16;
  static constexpr int NTY3D = sizeof(T) == 4 ? 12 : 8;
  static constexpr bool TMA = true;
  static constexpr bool SYNC_AFTER_INIT = false;
  typedef Col<T, D, 0, 0, 0, 0> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    T* out;        // PQ: p_out ; SWEEP: x_out
    DRef coef;     // PQ: beta
    IRef first;    // PQ: first spin (p = r), per problem
    int first_all;
    const tvdk_stat_t* skip;  // done ; PQ: cg_stop ; SWEEP: jac_stop
    Fin fin;
    TmaMap tm_a, tm_b;
  };
  struct Lay {
    int pA, pB, oA, oB, stage_bytes, stage_off, total;
    unsigned bytesA, bytesB;
  };
  static TVDK_HD int a128(int x) { return (x + 127) / 128 * 128; }
  static TVDK_HD int box_w(int which, int ntx) { return (which == 0 || MODE == TS_PQ) ? (ntx + 2) * V : ntx * V; }
  static TVDK_HD int box_h(int which, int nty) { return (which == 0 || MODE == TS_PQ) ? nty + 2 : nty; }
  static TVDK_HD Lay layout(int ntx, int nty) {
    Lay L;
    const int W = (int)sizeof(T);
    L.pA = box_w(0, ntx);
    L.pB = box_w(1, ntx);
    L.bytesA = (unsigned)(L.pA * box_h(0, nty) * W);
    L.bytesB = (unsigned)(L.pB * box_h(1, nty) * W);
    L.oA = 0;
    L.oB = a128((int)L.bytesA);
    L.stage_bytes = L.oB + a128((int)L.bytesB);
    L.stage_off = 128;
    L.total = L.stage_off + NS * L.stage_bytes;
    return L;
  }
  static TVDK_HD size_t smem_bytes(int ntx, int nty) { return (size_t)layout(ntx, nty).total; }

  C c;
  Lay L;
  int j, x0b, y0b, zb, sr, kq, iA, iB;
  T beta;
  bool first, leader;
  Vec<T> p_lo, p_c;
  double red[NRED > 0 ? NRED : 1];

  static TVDK_DEV bool skipped(const Args& a, const LaunchCtx& lc) {
    if (!a.skip) return false;
    const tvdk_stat_t& s = a.skip[lc.bz];
    return s.done || (MODE == TS_PQ ? s.cg_stop : s.jac_stop);
  }
  static TVDK_DEV void bar_init(const Args&, const LaunchCtx& lc, unsigned char* smem) {
    if (lc.tx == 0 && lc.ty == 0) {
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
      for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
      mbar_init_fence();
    }
  }
  TVDK_DEV unsigned long long* bar_at(unsigned char* smem, int idx) const { return reinterpret_cast<unsigned long long*>(smem) + idx; }
  TVDK_DEV unsigned char* stage_at(unsigned char* smem, int idx) const { return smem + L.stage_off + idx * L.stage_bytes; }
  TVDK_DEV void issue(const Args& a, unsigned char* smem, int p, int idx) const {
    unsigned long long* b = bar_at(smem, idx);
    unsigned char* s = stage_at(smem, idx);
    const bool needA = !(MODE == TS_PQ && first);  // first CG spin: p = r, the old p is never read
    mbar_expect_tx(b, (needA ? L.bytesA : 0u) + L.bytesB);
    const int z = zb + p;
    if (needA) tma_load_3d(s + L.oA, &a.tm_a, x0b - V, y0b - 1, z, b);
    if (MODE == TS_PQ)
      tma_load_3d(s + L.oB, &a.tm_b, x0b - V, y0b - 1, z, b);
    else
      tma_load_3d(s + L.oB, &a.tm_b, x0b, y0b, z, b);
  }
  // the stencil field at box offset `i` of the stage: PQ: p_new = beta p + r (r alone on the first spin); SWEEP: x
  TVDK_DEV Vec<T> fieldv(const unsigned char* s, int i) const {
    if (MODE == TS_SWEEP) return vload(reinterpret_cast<const T*>(s + L.oA) + i);
    const Vec<T> r = vload(reinterpret_cast<const T*>(s + L.oB) + i);
    if (first) return r;
    const Vec<T> p = vload(reinterpret_cast<const T*>(s + L.oA) + i);
    Vec<T> o;
#pragma unroll
    for (int v = 0; v < V; ++v) o.v[v] = beta * p.v[v] + r.v[v];  // p *= beta; p += r
    return o;
  }
  TVDK_DEV T field1(const unsigned char* s, int i) const {
    if (MODE == TS_SWEEP) return reinterpret_cast<const T*>(s + L.oA)[i];
    const T r = reinterpret_cast<const T*>(s + L.oB)[i];
    if (first) return r;
    return beta * reinterpret_cast<const T*>(s + L.oA)[i] + r;
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char* smem) {
    c.setup(a.g, lc);
    L = layout(lc.ntx, lc.nty);
    if (NRED > 0) red[0] = 0.0;
    leader = lc.tx == 0 && lc.ty == 0;
    const int xt = lc.bx % a.g.xtiles, yt = lc.bx / a.g.xtiles;
    x0b = xt * C::iw(lc.ntx);
    y0b = yt * C::ih(lc.nty);
    zb = c.b * a.g.n0;
    c.edge = x0b == 0 || y0b == 0 || x0b + lc.ntx * V >= a.g.nx || y0b + lc.nty >= a.g.ny;
    iA = (lc.ty + 1) * L.pA + (lc.tx + 1) * V;
    iB = MODE == TS_PQ ? iA : lc.ty * L.pB + lc.tx * V;
    beta = (MODE == TS_PQ && a.coef.p) ? (T)a.coef.at(c.b) : (T)0;
    first = MODE == TS_PQ && (a.first_all || a.first.at(c.b));
    j = c.z0;
    sr = 1;
    kq = 0;
    if (leader) {
      for (int i = 0; i < NS && c.z0 - 1 + i <= c.z1; ++i) issue(a, smem, c.z0 - 1 + i, i);
    }
    mbar_wait(bar_at(smem, 0), 0);
    mbar_wait(bar_at(smem, 1), 0);
    p_c = fieldv(stage_at(smem, 1), iA);
    p_lo = c.has_zlo(a.g, j) ? fieldv(stage_at(smem, 0), iA) : p_c;
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }

  template <bool E>
  TVDK_DEV void step(const Args& a, unsigned char* smem) {
    const Scalars<T>& sc = a.k.sc;
    const bool wrap = sr + 1 == NS;
    const int sr1 = wrap ? 0 : sr + 1;
    mbar_wait(bar_at(smem, sr1), (unsigned)((kq + (wrap ? 1 : 0)) & 1));
    const unsigned char* sj = stage_at(smem, sr);
    const Vec<T> p_hi = c.has_zhi(a.g, j) ? fieldv(stage_at(smem, sr1), iA) : p_c;
    const Vec<T> ym = fieldv(sj, iA - L.pA), yp = fieldv(sj, iA + L.pA);
    const T xl = field1(sj, iA - 1), xr = field1(sj, iA + V);
    const bool zlo = c.template ZLO<E>(a.g, j), zhi = c.template ZHI<E>(a.g, j);
    Vec<T> o;
    if (MODE == TS_PQ) {  // q = omega p - L_ell p with the half-weight boundary rows (derivatives.py:88-104), as apply_A
      const T l = a.k.ell.val;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        T s = lap_axis(p_lo.v[v], p_c.v[v], p_hi.v[v], zlo, zhi, l, l, l);
        s = s + lap_axis(ym.v[v], p_c.v[v], yp.v[v], c.template YLO<E>(), c.template YHI<E>(), l, l, l);
        const T left = v > 0 ? p_c.v[v > 0 ? v - 1 : 0] : xl;
        const T right = v < V - 1 ? p_c.v[v < V - 1 ? v + 1 : 0] : xr;
        s = s + lap_axis(left, p_c.v[v], right, c.template HL<E>(v), c.template HR<E>(v), l, l, l);
        o.v[v] = sc.omega * p_c.v[v] - s;
      }
      sto(c, a.out, c.cur, p_c);
      if (c.st) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.template OK<E>(v)) red[0] += (double)p_c.v[v] * (double)o.v[v];
      }
    } else {  // N(x) in the reference's order (jacobi.py:40-55), as nsum
      const Vec<T> rhs = vload(reinterpret_cast<const T*>(sj + L.oB) + iB);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        T s = (zlo ? p_lo.v[v] : p_c.v[v]) + (zhi ? p_hi.v[v] : p_c.v[v]);
        s += c.template YLO<E>() ? ym.v[v] : p_c.v[v];
        s += c.template YHI<E>() ? yp.v[v] : p_c.v[v];
        const T left = v > 0 ? p_c.v[v > 0 ? v - 1 : 0] : xl;
        const T right = v < V - 1 ? p_c.v[v < V - 1 ? v + 1 : 0] : xr;
        s += c.template HL<E>(v) ? left : p_c.v[v];
        s += c.template HR<E>(v) ? right : p_c.v[v];
        o.v[v] = rhs.v[v] * sc.inv_diag + s * sc.c_nb;
      }
      sto(c, a.out, c.cur, o);
    }
    p_lo = p_c;
    p_c = p_hi;
  }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    if (c.edge || c.zedge(a.g, j))
      step<true>(a, smem);
    else
      step<false>(a, smem);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    if (leader) {
      if (j == c.z0 && j - 1 + NS <= c.z1) issue(a, smem, j - 1 + NS, 0);
      if (j + NS <= c.z1) issue(a, smem, j + NS, sr);
    }
    if (++sr == NS) {
      sr = 0;
      ++kq;
    }
    c.cur += c.s0b;
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED > 0 ? NRED : 1, b, tot); }
};

// ==========================================================================================
// Slab decomposition: halo exchange + all-reduce over peer memory (tvdk_slab.inl; include/tvdk.h)
//   Every rank's arena starts with a SlabHeader.  slot[parity][s] is the mailbox rank s writes into
//   (on EVERY rank, its own included): up to 8 partial sums, then -- after a system-wide fence --
//   the sequence number of the step.  The halo planes a rank pushes into its neighbours' ghost planes
//   are fenced the same way, so "flag of rank s seen" implies "planes and sums of rank s landed".
// ==========================================================================================
struct alignas(128) MailSlot {
  double v[8];
  unsigned long long flag;
  unsigned long long pad[7];
};
struct SlabHeader {
  MailSlot slot[2][TVDK_MAX_SLABS];
  double local_sums[8];   // this rank's partial sums of the current step (written by the compute kernel)
  long long error_seq;    // first step whose flag wait timed out (0 = none)
  unsigned counter;       // last-block detection of the push kernel
  unsigned pad_[27];
};

struct CommPush {
  static constexpr int kMaxSeg = 16;
  struct Seg {
    const unsigned char* src;
    unsigned char* dst;
    long long bytes;  // multiple of 16
  } seg[kMaxSeg];
  int nseg;
  const double* sums;
  int nred;
  MailSlot* mail[TVDK_MAX_SLABS];  // mail[s]: this rank's slot in rank s's header
  int world;
  unsigned long long flag;
  unsigned* counter;
};

struct CommFin {
  const MailSlot* mail;  // this rank's header: slot[parity][0 .. world)
  int world, nred;
  unsigned long long flag;
  long long timeout_ns;
  long long* error;
  Fin fin;
  int skip_mask;  // 1: stats.done, 2: stats.cg_stop, 4: stats.jac_stop -- the compute kernel skipped, keep the scalars
  const tvdk_stat_t* stats;
};

// (mailbox lines are written by peers while this kernel polls them: read them past L1)
TVDK_DEV double ld_mail(const double* p) { return *reinterpret_cast<const volatile double*>(p); }

// ordered sum over the ranks + scalar update; runs in ONE thread once every flag has been seen
TVDK_DEV void comm_finalize_body(const CommFin& a) {
  if (a.stats) {
    const tvdk_stat_t& s = a.stats[0];
    if (((a.skip_mask & 1) && s.done) || ((a.skip_mask & 2) && s.cg_stop) || ((a.skip_mask & 4) && s.jac_stop)) return;
  }
  if (a.nred <= 0) return;
  double tot[8];
  for (int r = 0; r < a.nred; ++r) {
    double t = 0.0;
    for (int s = 0; s < a.world; ++s) t += ld_mail(&a.mail[s].v[r]);  // rank order: bitwise identical on every rank
    tot[r] = t;
  }
  finalize_fin(a.fin, a.nred, 0, tot);
}

}  // namespace tvdk
