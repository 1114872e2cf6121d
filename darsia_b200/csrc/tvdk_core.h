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

TVDK_DEV double tsqrt(double x) { return sqrt(x); }
TVDK_DEV float tsqrt(float x) { return sqrtf(x); }
TVDK_DEV double tabs(double x) { return fabs(x); }
TVDK_DEV float tabs(float x) { return fabsf(x); }
TVDK_DEV double tmax(double a, double b) { return a > b ? a : b; }  // NaN-free inputs assumed
TVDK_DEV float tmax(float a, float b) { return a > b ? a : b; }
template <typename T>
TVDK_DEV T tsign(T x) {
  return (T)((x > (T)0) - (x < (T)0));
}

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
  FIN_CGINIT,  // totals = {sum rhs^2, sum r^2}
  FIN_CGPQ,    // totals[0] = p.q
  FIN_CGRR,    // totals[0] = r.r
  FIN_JTOL,    // totals = {sum (x_new - x)^2, sum x_in^2}
  FIN_ENERGY   // totals = {sum (omega (x - f))^2, sum |mu G x|}
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
  int buf_id;          // FIN_OUTER: buffer now holding u ; FIN_JTOL: buffer the sweep wrote
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
      if (f.eps >= 0.0 && inc < f.eps) {
        s.done = 1;
        if (f.ndone) atomic_inc(f.ndone);
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
    case FIN_JTOL: {  // jacobi.py:118-131 (old iterate kept on convergence, Q7)
      if (f.first_sweep) s.jac_x0nrm2 = tot[1];
      double err = sqrt(tot[0]) / sqrt(s.jac_x0nrm2);
      s.jac_err = err;
      if (err < f.tol)
        s.jac_stop = 1;
      else
        s.jac_sel = f.buf_id;
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
template <typename T, int D, int LH, int RH>
struct Col {
  static constexpr int V = Vec<T>::N;
  static constexpr int RHX = RH > 0 ? 1 : 0;
  int X, Y, b;    // global coordinates of the first element of the thread's vector; problem
  int soff;       // shared-memory element offset of the vector inside one row array
  int spitch;     // shared-memory row pitch (elements)
  long long off;  // element offset of (b, plane 0, Y, X)
  int z0, z1;     // planes [z0, z1) owned by this CTA
  bool ld;        // the vector may be loaded (inside the allocation)
  bool st;        // interior thread whose results are stored
  bool okc[V];    // element inside the image
  bool hl[V], hr[V];  // element has a left / right neighbour along x
  bool hylo, hyhi;    // row has a lower / upper neighbour along y (3-D)

  static TVDK_HD int iw(int ntx) { return (ntx - LH - RHX) * V; }
  static TVDK_HD int ih(int nty) { return D == 3 ? nty - LH - RH : 1; }
  static TVDK_HD int row_elems(int ntx, int nty) { return (nty + 2) * (ntx + 2) * V; }

  TVDK_DEV void setup(const Geom& g, const LaunchCtx& lc) {
    const int xt = lc.bx % g.xtiles, yt = lc.bx / g.xtiles;
    X = xt * iw(lc.ntx) + (lc.tx - LH) * V;
    bool rowok = true;
    Y = 0;
    if (D == 3) {
      Y = yt * ih(lc.nty) + (lc.ty - LH);
      rowok = (Y >= 0) && (Y < g.ny);
    }
    b = lc.bz;
    z0 = g.z_begin + lc.by * g.zchunk;
    z1 = z0 + g.zchunk < g.z_end ? z0 + g.zchunk : g.z_end;
    ld = rowok && X >= 0 && X < g.pitch;
    const bool xin = lc.tx >= LH && lc.tx < lc.ntx - RHX;
    const bool yin = (D == 3) ? (lc.ty >= LH && lc.ty < lc.nty - RH) : true;
    st = ld && xin && yin && X < g.nx;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      okc[v] = rowok && (X + v >= 0) && (X + v < g.nx);
      hl[v] = X + v > 0;
      hr[v] = X + v < g.nx - 1;
    }
    hylo = Y > 0;
    hyhi = Y < g.ny - 1;
    spitch = (lc.ntx + 2) * V;
    soff = (lc.ty + 1) * spitch + (lc.tx + 1) * V;
    off = (long long)b * g.sb + (long long)Y * g.pitch + X;
  }
  // plane z exists in the local allocation and in the global domain
  TVDK_DEV bool plane_ok(const Geom& g, int z) const {
    const int gz = z + g.g0_off;
    return z >= 0 && z < g.n0 && gz >= 0 && gz < g.g0_n;
  }
  TVDK_DEV bool has_zlo(const Geom& g, int z) const { return z + g.g0_off > 0; }
  TVDK_DEV bool has_zhi(const Geom& g, int z) const { return z + g.g0_off < g.g0_n - 1; }
};

template <typename T, typename C>
TVDK_DEV Vec<T> ldv(const C& c, const Geom& g, const T* base, int z) {
  if (c.ld) return vload_ro(base + c.off + (long long)z * g.s0);
  return vsplat<T>((T)0);
}
template <typename T, typename C>
TVDK_DEV Vec<T> ldv_rw(const C& c, const Geom& g, const T* base, int z) {
  if (c.ld) return vload(base + c.off + (long long)z * g.s0);
  return vsplat<T>((T)0);
}
template <typename T, typename C>
TVDK_DEV Vec<T> ldcoef(const C& c, const Geom& g, const Coef<T>& k, int z) {
  if (k.p && c.ld)
    return vload_ro(k.p + (long long)c.b * k.sb + (c.off - (long long)c.b * g.sb) + (long long)z * g.s0);
  return vsplat<T>(k.val);
}
template <typename T, typename C>
TVDK_DEV void stv(const C& c, const Geom& g, T* base, int z, const Vec<T>& x) {
  if (c.st) vstore(base + c.off + (long long)z * g.s0, x);
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
template <typename T, int D, bool ISO>
TVDK_DEV void bw_from_t(const T (&t)[D], T kappa, T ell, T (&b)[D], T (&w)[D]) {
  if (ISO) {
    T s = t[0] * t[0];
#pragma unroll
    for (int a = 1; a < D; ++a) s = s + t[a] * t[a];
    s = tsqrt(s);
    const T phi = tmax(s - kappa, (T)0) / (s + (T)1e-18);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const T d = t[a] * phi;
      b[a] = t[a] * ((T)1 - phi);
      w[a] = ell * (b[a] - d);
    }
  } else {
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const T d = tmax(tabs(t[a]) - kappa, (T)0) * tsign(t[a]);
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
template <typename T>
TVDK_DEV VoxCoef<T> voxcoef(const Scalars<T>& sc, T mu, T om, T el) {
  VoxCoef<T> r;
  if (!sc.hetero) {
    r.kappa = sc.kappa;
    r.ell = sc.ell;
    r.omega = sc.omega;
    r.inv_diag = sc.inv_diag;
    r.c_nb = sc.c_nb;
  } else {
    r.kappa = mu / el;
    r.ell = el;
    r.omega = om;
    const T diag = om + el * (T)2 * sc.dim_t;  // omega + ell*2*dim/h**2, h = 1 (jacobi.py:72)
    r.inv_diag = (T)1 / diag;
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
template <typename T, int D, typename C>
TVDK_DEV Vec<T> nsum(const C& c, const Vec<T>& lo, const Vec<T>& hi, const Vec<T>& xc, const T* row) {
  constexpr int V = Vec<T>::N;
  Vec<T> s;
#pragma unroll
  for (int v = 0; v < V; ++v) s.v[v] = lo.v[v] + hi.v[v];
  if (D == 3) {
    const Vec<T> ym = sym(row, c), yp = syp(row, c);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      s.v[v] += c.hylo ? ym.v[v] : xc.v[v];
      s.v[v] += c.hyhi ? yp.v[v] : xc.v[v];
    }
  }
  const T xl = sxm(row, c), xr = sxp(row, c);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const T left = v > 0 ? xc.v[v > 0 ? v - 1 : 0] : xl;
    const T right = v < V - 1 ? xc.v[v < V - 1 ? v + 1 : 0] : xr;
    s.v[v] += c.hl[v] ? left : xc.v[v];
    s.v[v] += c.hr[v] ? right : xc.v[v];
  }
  return s;
}

// A x = omega x - L_ell x (cg.py:37-44) with x's z-neighbours in registers and its row in smem.
// Scalar ell only here; the per-voxel-ell variant loads ell's neighbours from global memory.
template <typename T, int D, typename C>
TVDK_DEV Vec<T> apply_A(const C& c, const Geom& g, int z, const Vec<T>& lo, const Vec<T>& xc,
                        const Vec<T>& hi, const T* row, const Vec<T>& om, const Coef<T>& ell) {
  constexpr int V = Vec<T>::N;
  const bool zlo = c.has_zlo(g, z), zhi = c.has_zhi(g, z);
  Vec<T> lap;
  const T xl = sxm(row, c), xr = sxp(row, c);
  Vec<T> ym = xc, yp = xc;
  if (D == 3) {
    ym = sym(row, c);
    yp = syp(row, c);
  }
  if (!ell.p) {
    const T l = ell.val;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T s = lap_axis(lo.v[v], xc.v[v], hi.v[v], zlo, zhi, l, l, l);
      if (D == 3) s = s + lap_axis(ym.v[v], xc.v[v], yp.v[v], c.hylo, c.hyhi, l, l, l);
      const T left = v > 0 ? xc.v[v > 0 ? v - 1 : 0] : xl;
      const T right = v < V - 1 ? xc.v[v < V - 1 ? v + 1 : 0] : xr;
      s = s + lap_axis(left, xc.v[v], right, c.hl[v], c.hr[v], l, l, l);
      lap.v[v] = s;
    }
  } else {
    // per-voxel ell inside the divergence (cg.py:43): rare path, ell's neighbours
    // come straight from global memory (L1/L2 cached)
    const T* lb = ell.p + (long long)c.b * ell.sb + (c.off - (long long)c.b * g.sb) + (long long)z * g.s0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (!c.okc[v] || !c.ld) {
        lap.v[v] = (T)0;
        continue;
      }
      const T lc = sload_ro(lb + v);
      T s = lap_axis(lo.v[v], xc.v[v], hi.v[v], zlo, zhi, zlo ? sload_ro(lb + v - g.s0) : lc, lc,
                     zhi ? sload_ro(lb + v + g.s0) : lc);
      if (D == 3)
        s = s + lap_axis(ym.v[v], xc.v[v], yp.v[v], c.hylo, c.hyhi,
                         c.hylo ? sload_ro(lb + v - g.pitch) : lc, lc,
                         c.hyhi ? sload_ro(lb + v + g.pitch) : lc);
      const T left = v > 0 ? xc.v[v > 0 ? v - 1 : 0] : xl;
      const T right = v < V - 1 ? xc.v[v < V - 1 ? v + 1 : 0] : xr;
      s = s + lap_axis(left, xc.v[v], right, c.hl[v], c.hr[v], c.hl[v] ? sload_ro(lb + v - 1) : lc,
                       lc, c.hr[v] ? sload_ro(lb + v + 1) : lc);
      lap.v[v] = s;
    }
  }
  Vec<T> y;
#pragma unroll
  for (int v = 0; v < V; ++v) y.v[v] = om.v[v] * xc.v[v] - lap.v[v];
  return y;
}

// ==========================================================================================
// K1  fused outer iteration for Jacobi(maxiter=1):  rhs -> sweep -> shrink -> Bregman update
//     split_bregman_tvd.py:84-89 (rhs), jacobi.py:101-114 (sweep), :136-146 (shrink/update),
//     :161 (increment).  u+ is recomputed on a +1 halo (threads of the RH rows / right vector).
// ==========================================================================================
template <typename T, int D, bool ISO>
struct KFusedJ1 {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 2, NRED = 1, NROWS = 2 * D + 2;
  // shared rows: [slot*D + 0] = u(j), [slot*D + d] = w_d(j) (d >= 1), [2D + slot] = u+ ring
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
  Vec<T> u_lo, u_c, u_hi;   // u(j-1), u(j), u(j+1)
  Vec<T> w0_prev, w0_cur;   // ell (b_0 - d_0) at planes j-1, j
  Vec<T> wa_cur[D];         // w_a(j), a >= 1 kept for the own-vector neighbour
  Vec<T> b_prev[D], b_cur[D];
  Vec<T> up_prev, f_c, invd, cnb, om_f;  // u+(j-1); plane-j data
  bool have_row;
  double red[NRED];

  TVDK_DEV void load_w0(const Args& a, int z, Vec<T>& w0) {
    // w_0 of plane z (prologue only): needs all components of t for the isotropic norm
    Vec<T> tv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) tv[d] = ldv(c, a.g, a.t_in + d * a.tcs_in, z);
    const Vec<T> mu = ldcoef(c, a.g, a.k.mu, z), el = ldcoef(c, a.g, a.k.ell, z);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T t[D], b[D], w[D];
#pragma unroll
      for (int d = 0; d < D; ++d) t[d] = tv[d].v[v];
      const VoxCoef<T> vc = voxcoef(a.k.sc, mu.v[v], (T)1, el.v[v]);
      bw_from_t<T, D, ISO>(t, vc.kappa, vc.ell, b, w);
      w0.v[v] = w[0];
    }
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = 0.0;
    j = c.z0;
    u_c = ldv(c, a.g, a.u_in, j);
    u_lo = u_c;
    w0_prev = vsplat<T>((T)0);
    if (c.has_zlo(a.g, j)) {
      u_lo = ldv(c, a.g, a.u_in, j - 1);
      load_w0(a, j - 1, w0_prev);
    }
    up_prev = vsplat<T>((T)0);
#pragma unroll
    for (int d = 0; d < D; ++d) b_prev[d] = vsplat<T>((T)0);
    have_row = false;
    return true;
  }
  TVDK_DEV bool more() const { return j <= c.z1; }

  // phase 1: load plane j, form b/w from t, publish u(j) and w_a(j) (a >= 1)
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    have_row = (j < c.z1) || c.plane_ok(a.g, j);
    if (!have_row) return;
    const int slot = j & 1;
    u_hi = c.has_zhi(a.g, j) ? ldv(c, a.g, a.u_in, j + 1) : u_c;
    f_c = ldv(c, a.g, a.f, j);
    Vec<T> tv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) tv[d] = ldv(c, a.g, a.t_in + d * a.tcs_in, j);
    const Vec<T> mu = ldcoef(c, a.g, a.k.mu, j), om = ldcoef(c, a.g, a.k.omega, j),
                 el = ldcoef(c, a.g, a.k.ell, j);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T t[D], b[D], w[D];
#pragma unroll
      for (int d = 0; d < D; ++d) t[d] = tv[d].v[v];
      const VoxCoef<T> vc = voxcoef(a.k.sc, mu.v[v], om.v[v], el.v[v]);
      bw_from_t<T, D, ISO>(t, vc.kappa, vc.ell, b, w);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        b_cur[d].v[v] = b[d];
        if (d == 0)
          w0_cur.v[v] = w[0];
        else
          wa_cur[d].v[v] = w[d];
      }
      invd.v[v] = vc.inv_diag;
      cnb.v[v] = vc.c_nb;
      om_f.v[v] = vc.omega * f_c.v[v];
    }
    sst(srow<T>(smem, slot * D + 0, rowe), c, u_c);
#pragma unroll
    for (int d = 1; d < D; ++d) sst(srow<T>(smem, slot * D + d, rowe), c, wa_cur[d]);
  }

  // phase 2: u+(j); then outputs of plane j-1
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const int slot = j & 1;
    Vec<T> up = vsplat<T>((T)0);
    if (have_row) {
      const T* urow = srow<T>(smem, slot * D + 0, rowe);
      // rhs = omega f + H_0 w_0 + H_1 w_1 (+ H_2 w_2), in the reference's order (:84-89)
      Vec<T> rhs = om_f;
      if (c.has_zlo(a.g, j)) {
#pragma unroll
        for (int v = 0; v < V; ++v) rhs.v[v] += w0_cur.v[v] - w0_prev.v[v];
      }
      if (D == 3) {
        const T* wrow = srow<T>(smem, slot * D + 1, rowe);
        const Vec<T> wm = sym(wrow, c);
        if (c.hylo) {
#pragma unroll
          for (int v = 0; v < V; ++v) rhs.v[v] += wa_cur[1].v[v] - wm.v[v];
        }
      }
      {
        const T* wrow = srow<T>(smem, slot * D + (D - 1), rowe);
        const T wl = sxm(wrow, c);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const T left = v > 0 ? wa_cur[D - 1].v[v > 0 ? v - 1 : 0] : wl;
          if (c.hl[v]) rhs.v[v] += wa_cur[D - 1].v[v] - left;
        }
      }
      const Vec<T> nb = nsum<T, D>(c, u_lo, u_hi, u_c, urow);
#pragma unroll
      for (int v = 0; v < V; ++v) up.v[v] = rhs.v[v] * invd.v[v] + nb.v[v] * cnb.v[v];
      sst(srow<T>(smem, 2 * D + slot, rowe), c, up);
    }
    if (j > c.z0) {
      // outputs of plane j-1: t_new = G u+ + b, u_out = u+, increment (:136-146, :161)
      const int zo = j - 1;
      const T* prow = srow<T>(smem, 2 * D + (zo & 1), rowe);
      Vec<T> tn[D];
      const bool zhi = c.has_zhi(a.g, zo);
#pragma unroll
      for (int v = 0; v < V; ++v)
        tn[0].v[v] = (zhi ? up.v[v] - up_prev.v[v] : (T)0) + b_prev[0].v[v];
      if (D == 3) {
        const Vec<T> yp = syp(prow, c);
#pragma unroll
        for (int v = 0; v < V; ++v)
          tn[1].v[v] = (c.hyhi ? yp.v[v] - up_prev.v[v] : (T)0) + b_prev[1].v[v];
      }
      {
        const T xr = sxp(prow, c);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const T right = v < V - 1 ? up_prev.v[v < V - 1 ? v + 1 : 0] : xr;
          tn[D - 1].v[v] = (c.hr[v] ? right - up_prev.v[v] : (T)0) + b_prev[D - 1].v[v];
        }
      }
      stv(c, a.g, a.u_out, zo, up_prev);
#pragma unroll
      for (int d = 0; d < D; ++d) stv(c, a.g, a.t_out + d * a.tcs_out, zo, tn[d]);
      if (c.st) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          if (c.okc[v]) {
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
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};
// NOTE on u_lo in the increment: when plane j-1 is output, u_lo still holds u(j-1) (it is rolled
// afterwards), so (u+ - u)(j-1) uses the right old value.

// ==========================================================================================
// K2  rhs from the compressed state, optionally fused with the first Jacobi sweep (MODE 1)
//     or with CG's r0 = rhs - A u, x = u (MODE 2).   split_bregman_tvd.py:84-89; jacobi.py:107-114;
//     cg.py:37-47 + scipy cg (r = b - A x).  With `expl` the first outer iteration of a warm
//     start reads the reference's explicit d and b (split_bregman_tvd.py:93-97).
// ==========================================================================================
enum { RHS_ONLY = 0, RHS_SWEEP = 1, RHS_CGINIT = 2 };

template <typename T, int D, bool ISO, int MODE>
struct KRhs {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 1, NRED = 2, NROWS = 2 * D;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* f;
    const T* t;        // compressed state, or nullptr when expl
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
  Vec<T> u_lo, u_c, u_hi, w0_prev, w0_cur, wa_cur[D], om_c, f_c, invd, cnb;
  double red[NRED];

  TVDK_DEV void load_w(const Args& a, int z, Vec<T> (&w)[D]) {
    Vec<T> tv[D], dv[D];
    const bool expl = a.t == nullptr;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (!expl)
        tv[d] = ldv(c, a.g, a.t + d * a.tcs, z);
      else {
        tv[d] = ldv(c, a.g, a.xb + d * a.xcs, z);
        dv[d] = ldv(c, a.g, a.xd + d * a.xcs, z);
      }
    }
    const Vec<T> mu = ldcoef(c, a.g, a.k.mu, z), el = ldcoef(c, a.g, a.k.ell, z);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const VoxCoef<T> vc = voxcoef(a.k.sc, mu.v[v], (T)1, el.v[v]);
      if (!expl) {
        T t[D], b[D], ww[D];
#pragma unroll
        for (int d = 0; d < D; ++d) t[d] = tv[d].v[v];
        bw_from_t<T, D, ISO>(t, vc.kappa, vc.ell, b, ww);
#pragma unroll
        for (int d = 0; d < D; ++d) w[d].v[v] = ww[d];
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) w[d].v[v] = vc.ell * (tv[d].v[v] - dv[d].v[v]);
      }
    }
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = red[1] = 0.0;
    j = c.z0;
    w0_prev = vsplat<T>((T)0);
    u_c = vsplat<T>((T)0);
    if (MODE != RHS_ONLY) u_c = ldv(c, a.g, a.u, j);
    u_lo = u_c;
    if (c.has_zlo(a.g, j)) {
      Vec<T> w[D];
      load_w(a, j - 1, w);
      w0_prev = w[0];
      if (MODE != RHS_ONLY) u_lo = ldv(c, a.g, a.u, j - 1);
    }
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }

  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    const int slot = j & 1;
    Vec<T> w[D];
    load_w(a, j, w);
    w0_cur = w[0];
#pragma unroll
    for (int d = 1; d < D; ++d) wa_cur[d] = w[d];
    f_c = ldv(c, a.g, a.f, j);
    const Vec<T> om = ldcoef(c, a.g, a.k.omega, j);
    om_c = om;
    if (MODE != RHS_ONLY) {
      u_hi = c.has_zhi(a.g, j) ? ldv(c, a.g, a.u, j + 1) : u_c;
      sst(srow<T>(smem, slot * D + 0, rowe), c, u_c);
      if (MODE == RHS_SWEEP) {
        const Vec<T> el = ldcoef(c, a.g, a.k.ell, j);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const VoxCoef<T> vc = voxcoef(a.k.sc, (T)0, om.v[v], el.v[v]);
          invd.v[v] = vc.inv_diag;
          cnb.v[v] = vc.c_nb;
        }
      }
    }
#pragma unroll
    for (int d = 1; d < D; ++d) sst(srow<T>(smem, slot * D + d, rowe), c, wa_cur[d]);
  }

  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const int slot = j & 1;
    Vec<T> rhs;
#pragma unroll
    for (int v = 0; v < V; ++v) rhs.v[v] = om_c.v[v] * f_c.v[v];
    if (c.has_zlo(a.g, j)) {
#pragma unroll
      for (int v = 0; v < V; ++v) rhs.v[v] += w0_cur.v[v] - w0_prev.v[v];
    }
    if (D == 3) {
      const Vec<T> wm = sym(srow<T>(smem, slot * D + 1, rowe), c);
      if (c.hylo) {
#pragma unroll
        for (int v = 0; v < V; ++v) rhs.v[v] += wa_cur[1].v[v] - wm.v[v];
      }
    }
    {
      const T wl = sxm(srow<T>(smem, slot * D + (D - 1), rowe), c);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const T left = v > 0 ? wa_cur[D - 1].v[v > 0 ? v - 1 : 0] : wl;
        if (c.hl[v]) rhs.v[v] += wa_cur[D - 1].v[v] - left;
      }
    }
    if (MODE == RHS_ONLY || MODE == RHS_SWEEP) stv(c, a.g, a.rhs, j, rhs);
    if (MODE == RHS_SWEEP) {
      const Vec<T> nb = nsum<T, D>(c, u_lo, u_hi, u_c, srow<T>(smem, slot * D + 0, rowe));
      Vec<T> xn;
#pragma unroll
      for (int v = 0; v < V; ++v) xn.v[v] = rhs.v[v] * invd.v[v] + nb.v[v] * cnb.v[v];
      stv(c, a.g, a.x, j, xn);
      if (c.st) {  // Jacobi tol branch, first sweep: ||x_new - x||^2 and ||x0||^2
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.okc[v]) {
            const double df = (double)xn.v[v] - (double)u_c.v[v];
            red[0] += df * df;
            red[1] += (double)u_c.v[v] * (double)u_c.v[v];
          }
      }
    }
    if (MODE == RHS_CGINIT) {
      const Vec<T> au = apply_A<T, D>(c, a.g, j, u_lo, u_c, u_hi, srow<T>(smem, slot * D + 0, rowe), om_c,
                                      a.k.ell);
      Vec<T> r0;
#pragma unroll
      for (int v = 0; v < V; ++v) r0.v[v] = rhs.v[v] - au.v[v];
      stv(c, a.g, a.r, j, r0);
      stv(c, a.g, a.x, j, u_c);
      if (c.st) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.okc[v]) {
            red[0] += (double)rhs.v[v] * (double)rhs.v[v];
            red[1] += (double)r0.v[v] * (double)r0.v[v];
          }
      }
    }
    u_lo = u_c;
    u_c = u_hi;
    w0_prev = w0_cur;
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K4  Jacobi sweep  x_out = rhs/diag + N(x_in)/(diag/ell)       (jacobi.py:101-131)
// ==========================================================================================
template <typename T, int D>
struct KSweep {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 1, NRED = 2, NROWS = 2;
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
  Vec<T> x_lo, x_c, x_hi;
  double red[NRED];

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && (a.skip[lc.bz].done || a.skip[lc.bz].jac_stop)) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = red[1] = 0.0;
    j = c.z0;
    x_c = ldv(c, a.g, a.x_in, j);
    x_lo = c.has_zlo(a.g, j) ? ldv(c, a.g, a.x_in, j - 1) : x_c;
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    x_hi = c.has_zhi(a.g, j) ? ldv(c, a.g, a.x_in, j + 1) : x_c;
    sst(srow<T>(smem, j & 1, rowe), c, x_c);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const Vec<T> nb = nsum<T, D>(c, x_lo, x_hi, x_c, srow<T>(smem, j & 1, rowe));
    const Vec<T> rhs = ldv(c, a.g, a.rhs, j);
    const Vec<T> om = ldcoef(c, a.g, a.k.omega, j), el = ldcoef(c, a.g, a.k.ell, j);
    Vec<T> xn;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const VoxCoef<T> vc = voxcoef(a.k.sc, (T)0, om.v[v], el.v[v]);
      xn.v[v] = rhs.v[v] * vc.inv_diag + nb.v[v] * vc.c_nb;
    }
    stv(c, a.g, a.x_out, j, xn);
    if (c.st) {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (c.okc[v]) {
          const double df = (double)xn.v[v] - (double)x_c.v[v];
          red[0] += df * df;
          red[1] += (double)x_c.v[v] * (double)x_c.v[v];
        }
    }
    x_lo = x_c;
    x_c = x_hi;
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K5/K6  CG iteration (scipy.sparse.linalg.cg recurrence, called from cg.py:50-56)
//   PQ:     p_out = first ? r : beta p_in + r ;  reduce p_out . (A p_out)
//   UPDATE: x += alpha p ; r -= alpha (A p) ;    reduce r . r      (q = A p is recomputed,
//           not stored: 8 instead of 10 words per voxel and CG iteration)
//   MATVEC: y = A x (test-granularity operator)
// ==========================================================================================
enum { CG_PQ = 0, CG_UPDATE = 1, CG_MATVEC = 2 };

template <typename T, int D, int MODE>
struct KCg {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 1, RH = 1, NRED = 1, NROWS = 2;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* p_in;  // PQ: old p ; UPDATE: p ; MATVEC: x
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
  double red[NRED];

  TVDK_DEV Vec<T> load_p(const Args& a, int z) {
    if (MODE == CG_PQ) {
      const Vec<T> r = ldv(c, a.g, a.r_in, z);
      if (first) return r;
      const Vec<T> p = ldv(c, a.g, a.p_in, z);
      Vec<T> o;
#pragma unroll
      for (int v = 0; v < V; ++v) o.v[v] = coef * p.v[v] + r.v[v];  // p *= beta; p += r
      return o;
    }
    return ldv(c, a.g, a.p_in, z);
  }

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && (a.skip[lc.bz].done || a.skip[lc.bz].cg_stop)) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = 0.0;
    coef = (MODE == CG_MATVEC || !a.coef.p) ? (T)0 : (T)a.coef.at(c.b);
    first = (MODE == CG_PQ) && (a.first_all || a.first.at(c.b));
    j = c.z0;
    p_c = load_p(a, j);
    p_lo = c.has_zlo(a.g, j) ? load_p(a, j - 1) : p_c;
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    p_hi = c.has_zhi(a.g, j) ? load_p(a, j + 1) : p_c;
    sst(srow<T>(smem, j & 1, rowe), c, p_c);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const Vec<T> om = ldcoef(c, a.g, a.k.omega, j);
    const Vec<T> q = apply_A<T, D>(c, a.g, j, p_lo, p_c, p_hi, srow<T>(smem, j & 1, rowe), om, a.k.ell);
    if (MODE == CG_PQ) {
      stv(c, a.g, a.p_out, j, p_c);
      if (c.st) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.okc[v]) red[0] += (double)p_c.v[v] * (double)q.v[v];
      }
    } else if (MODE == CG_UPDATE) {
      if (c.st) {  // x and r are updated in place: only the owning thread touches them
        Vec<T> x = ldv_rw(c, a.g, a.x, j), r = ldv_rw(c, a.g, a.r, j);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          x.v[v] += coef * p_c.v[v];
          r.v[v] -= coef * q.v[v];
        }
        stv(c, a.g, a.x, j, x);
        stv(c, a.g, a.r, j, r);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (c.okc[v]) red[0] += (double)r.v[v] * (double)r.v[v];
      }
    } else {
      stv(c, a.g, a.p_out, j, q);
    }
    p_lo = p_c;
    p_c = p_hi;
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// K3  shrinkage + Bregman update on the compressed state (split_bregman_tvd.py:136-146),
//     increment partials (:161), in-place update of the iterate (:170).
// ==========================================================================================
template <typename T, int D, bool ISO>
struct KShrink {
  static constexpr int V = Vec<T>::N;
  static constexpr int LH = 0, RH = 1, NRED = 1, NROWS = 2;
  typedef Col<T, D, LH, RH> C;
  struct Args {
    Geom g;
    Coefs<T> k;
    const T* unew[3];  // candidate buffers holding u+ ; selected by stats[b].jac_sel if use_sel
    int use_sel;
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
  Vec<T> up_c, up_hi;
  double red[NRED];

  TVDK_DEV bool init(const Args& a, const LaunchCtx& lc, unsigned char*) {
    if (a.skip && a.skip[lc.bz].done) return false;
    c.setup(a.g, lc);
    rowe = C::row_elems(lc.ntx, lc.nty);
    red[0] = 0.0;
    un = a.unew[a.use_sel ? a.skip[lc.bz].jac_sel : 0];
    j = c.z0;
    up_c = ldv(c, a.g, un, j);
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    up_hi = c.has_zhi(a.g, j) ? ldv(c, a.g, un, j + 1) : up_c;
    sst(srow<T>(smem, j & 1, rowe), c, up_c);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const T* prow = srow<T>(smem, j & 1, rowe);
    const bool expl = a.t_in == nullptr;
    // t and u may be updated in place (t_out == t_in, u_out == u_old): only the owning
    // (interior) thread touches them; halo threads merely publish u+ rows.
    Vec<T> tv[D];
    Vec<T> uo = vsplat<T>((T)0);
#pragma unroll
    for (int d = 0; d < D; ++d) tv[d] = vsplat<T>((T)0);
    if (c.st) {
#pragma unroll
      for (int d = 0; d < D; ++d)
        tv[d] = expl ? ldv_rw(c, a.g, a.xb + d * a.xcs, j) : ldv_rw(c, a.g, a.t_in + d * a.tcs, j);
      uo = ldv_rw(c, a.g, a.u_old, j);
    }
    const Vec<T> mu = ldcoef(c, a.g, a.k.mu, j), el = ldcoef(c, a.g, a.k.ell, j);
    Vec<T> g[D];
    const bool zhi = c.has_zhi(a.g, j);
#pragma unroll
    for (int v = 0; v < V; ++v) g[0].v[v] = zhi ? up_hi.v[v] - up_c.v[v] : (T)0;
    if (D == 3) {
      const Vec<T> yp = syp(prow, c);
#pragma unroll
      for (int v = 0; v < V; ++v) g[1].v[v] = c.hyhi ? yp.v[v] - up_c.v[v] : (T)0;
    }
    {
      const T xr = sxp(prow, c);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const T right = v < V - 1 ? up_c.v[v < V - 1 ? v + 1 : 0] : xr;
        g[D - 1].v[v] = c.hr[v] ? right - up_c.v[v] : (T)0;
      }
    }
    Vec<T> tn[D];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T b[D];
      if (expl) {
#pragma unroll
        for (int d = 0; d < D; ++d) b[d] = tv[d].v[v];
      } else {
        T t[D], w[D];
#pragma unroll
        for (int d = 0; d < D; ++d) t[d] = tv[d].v[v];
        const VoxCoef<T> vc = voxcoef(a.k.sc, mu.v[v], (T)1, el.v[v]);
        bw_from_t<T, D, ISO>(t, vc.kappa, vc.ell, b, w);
      }
#pragma unroll
      for (int d = 0; d < D; ++d) tn[d].v[v] = g[d].v[v] + b[d];
    }
#pragma unroll
    for (int d = 0; d < D; ++d) stv(c, a.g, a.t_out + d * a.tcs, j, tn[d]);
    if (a.u_out) stv(c, a.g, a.u_out, j, up_c);
    if (c.st) {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (c.okc[v]) {
          const double df = (double)up_c.v[v] - (double)uo.v[v];
          red[0] += df * df;
        }
    }
    up_c = up_hi;
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
};

// ==========================================================================================
// Energy functional (verbose):  1/2 ||omega (x - f)||^2 + sum_a sum |mu G_a x|  (:71-77)
// ==========================================================================================
template <typename T, int D>
struct KEnergy {
  static constexpr int V = Vec<T>::N;
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
    x_c = ldv(c, a.g, a.x, j);
    return true;
  }
  TVDK_DEV bool more() const { return j < c.z1; }
  TVDK_DEV void pre(const Args& a, unsigned char* smem) {
    x_hi = c.has_zhi(a.g, j) ? ldv(c, a.g, a.x, j + 1) : x_c;
    sst(srow<T>(smem, j & 1, rowe), c, x_c);
  }
  TVDK_DEV void post(const Args& a, unsigned char* smem) {
    const T* prow = srow<T>(smem, j & 1, rowe);
    const Vec<T> f = ldv(c, a.g, a.f, j);
    const Vec<T> mu = ldcoef(c, a.g, a.k.mu, j), om = ldcoef(c, a.g, a.k.omega, j);
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
    ++j;
  }
  static TVDK_DEV void finalize(const Args& a, int b, const double* tot) { finalize_fin(a.fin, NRED, b, tot); }
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
      for (int z = c.z0; z < c.z1; ++z) {
        const Vec<T> x = ldv(c, a.g, a.x, z), y = ldv(c, a.g, a.y, z);
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
      for (int z = c.z0; z < c.z1; ++z) {
        Vec<T> tv[D], dv[D], bv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) tv[d] = ldv(c, a.g, a.t + d * a.tcs, z);
        const Vec<T> mu = ldcoef(c, a.g, a.k.mu, z), el = ldcoef(c, a.g, a.k.ell, z);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T t[D], b[D], w[D];
#pragma unroll
          for (int d = 0; d < D; ++d) t[d] = tv[d].v[v];
          const VoxCoef<T> vc = voxcoef(a.k.sc, mu.v[v], (T)1, el.v[v]);
          bw_from_t<T, D, ISO>(t, vc.kappa, vc.ell, b, w);
#pragma unroll
          for (int d = 0; d < D; ++d) {
            bv[d].v[v] = b[d];
            dv[d].v[v] = t[d] - b[d];  // d = t - b  (b = t - d in the reference)
          }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
          stv(c, a.g, a.d + d * a.ocs, z, dv[d]);
          stv(c, a.g, a.b + d * a.ocs, z, bv[d]);
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
    if (c.st)
      for (int z = c.z0; z < c.z1; ++z) stv(c, a.g, a.dst, z, ldv(c, a.g, a.src, z));
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

}  // namespace tvdk
