// tvdk_slab.inl -- slab-decomposed split-Bregman solve (one volume cut along axis 0 over several
// GPUs): the host driver behind tvdk_solve_slab / tvdk_slab_exchange (include/tvdk.h).
//
// Included after tvdk_host.inl by exactly one translation unit per backend.  The backend provides
//   namespace tvdk::backend {
//     int comm_push(const CommPush&, void* stream);       // boundary planes + partial sums -> peers, then the flag
//     int comm_finalize(const CommFin&, void* stream);    // wait for every rank's flag, ordered sum, finalize_fin
//     int comm_sync_local(int n, void* const* streams, const int* devices);  // event edges between local streams
//     int set_device(int dev);                            // -1: no-op
//     int read_int(const int* dev_ptr, void* stream, int* value);            // blocking read of a device int
//     int read_i64(const long long* dev_ptr, void* stream, long long* value);
//   }
// Same arithmetic per voxel as solve_impl (the kernels evaluate every boundary rule on GLOBAL plane indices,
// tvdk_geom_t::g0_off / g0_n); only the order in which the per-rank partial sums are added differs.
//
// Reference: restoration/split_bregman_tvd.py:14-177 with utils/linear_solvers/jacobi.py:74-133 or
// cg.py:16-59 + SciPy's cg recurrence; the reference itself has no distributed code.
#include <stddef.h>

#include <vector>

namespace tvdk {

// one rank of a lock-step group driven by this host thread
template <typename T>
struct SlabRank {
  const tvdk_problem_t* p;
  const tvdk_slab_comm_t* cm;
  void* stream;
  int device;
  Geom g;
  Scratch sc;
  SlabHeader* hdr;          // this rank's arena header
  Coefs<T> k;
  const T* f;
  T* ubuf[2];
  T* tbuf[2];
  long long tcs[2];
  T *RB, *XA, *XB, *X, *P[2];
  tvdk_stat_t* st;
  int cur;
};

static int check_slab(const tvdk_problem_t* p, const tvdk_slab_comm_t* c) {
  if (!p || !c) return fail(TVDK_ERR_INVALID, "problem / comm is NULL");
  TVDK_TRY(check_geom(&p->geom));
  if (c->world < 1 || c->world > TVDK_MAX_SLABS) return fail(TVDK_ERR_INVALID, "world = %d out of range [1, %d]", c->world, TVDK_MAX_SLABS);
  if (c->rank < 0 || c->rank >= c->world) return fail(TVDK_ERR_INVALID, "rank %d not in [0, %d)", c->rank, c->world);
  if (!c->seq) return fail(TVDK_ERR_INVALID, "comm.seq is NULL");
  if (p->geom.nb != 1) return fail(TVDK_ERR_INVALID, "slab decomposition handles one volume (nb = 1)");
  int cap = 0;
  long long total = 0;
  for (int r = 0; r < c->world; ++r) {
    if (c->nown[r] < TVDK_SLAB_GH && c->world > 1) return fail(TVDK_ERR_INVALID, "every slab needs at least %d planes", TVDK_SLAB_GH);
    if (!c->arena[r]) return fail(TVDK_ERR_INVALID, "arena[%d] is NULL", r);
    cap = std::max(cap, c->nown[r]);
    total += c->nown[r];
  }
  const tvdk_geom_t& g = p->geom;
  if (g.n[0] != TVDK_SLAB_GL + cap + TVDK_SLAB_GH)
    return fail(TVDK_ERR_INVALID, "geom.n[0] = %d must be GL + max nown + GH = %d on every rank", g.n[0], TVDK_SLAB_GL + cap + TVDK_SLAB_GH);
  if (g.z_begin != TVDK_SLAB_GL || g.z_end != TVDK_SLAB_GL + c->nown[c->rank])
    return fail(TVDK_ERR_INVALID, "owned planes must be [GL, GL + nown[rank])");
  long long first = 0;
  for (int r = 0; r < c->rank; ++r) first += c->nown[r];
  if (g.g0_n != total || g.g0_off != first - TVDK_SLAB_GL)
    return fail(TVDK_ERR_INVALID, "g0_n / g0_off do not match the slab sizes (expected %lld / %lld)", total, first - TVDK_SLAB_GL);
  if (p->x0_b || p->x0_d) return fail(TVDK_ERR_UNSUPPORTED, "warm starts are not implemented for slab-decomposed solves");
  if (p->adaptive_mask) return fail(TVDK_ERR_UNSUPPORTED, "the adaptive schedule is not implemented for slab-decomposed solves");
  if (p->verbose_every > 0) return fail(TVDK_ERR_UNSUPPORTED, "verbose energies are not implemented for slab-decomposed solves");
  if (p->solver_kind != TVDK_SOLVER_JACOBI && p->solver_kind != TVDK_SOLVER_CG)
    return fail(TVDK_ERR_UNSUPPORTED, "unknown solver kind %d (Jacobi and CG only)", p->solver_kind);
  if (p->solver_maxiter < 1) return fail(TVDK_ERR_INVALID, "solver maxiter must be >= 1");
  if (p->solver_kind == TVDK_SOLVER_CG && p->solver_tol < 0.0)
    return fail(TVDK_ERR_INVALID, "CG needs a numeric tol (SciPy rejects rtol=None); use 0.0 for a fixed count");
  // every pointer the neighbours write through must lie inside the arena
  const unsigned char* a0 = reinterpret_cast<const unsigned char*>(c->arena[c->rank]);
  auto inside = [&](const void* q) {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(q);
    return q && b >= a0 + sizeof(SlabHeader) && b < a0 + c->arena_bytes;
  };
  if (!inside(p->f) || !inside(p->u) || !inside(p->t) || !inside(p->workspace))
    return fail(TVDK_ERR_INVALID, "f, u, t and the workspace must live inside the rank's arena, behind the header");
  if (reinterpret_cast<const unsigned char*>(p->workspace) + p->workspace_bytes > a0 + c->arena_bytes)
    return fail(TVDK_ERR_INVALID, "the workspace exceeds the arena");
  return 0;
}

// pointer of the same field in rank r's arena, as mapped for this rank
template <typename T>
static T* peer_ptr(const tvdk_slab_comm_t* c, const T* mine, int r) {
  const ptrdiff_t off = reinterpret_cast<const unsigned char*>(mine) - reinterpret_cast<const unsigned char*>(c->arena[c->rank]);
  return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(c->arena[r]) + off);
}

struct HaloSpec {
  const void* field;  // base of one field (one component) in this rank's arena
  int n_up, n_down;
};

// Enqueue, for every local rank: push (halo planes + partial sums + flag) ; local stream edges ; finalize.
template <typename T>
static int comm_step(std::vector<SlabRank<T>>& R, const std::vector<std::vector<HaloSpec>>& halos, int nred,
                     const std::vector<Fin>& fins, int skip_mask) {
  const int nl = (int)R.size();
  const unsigned long long seq = ++(*R[0].cm->seq);
  for (int i = 1; i < nl; ++i) *R[i].cm->seq = seq;  // (ranks of one process may share or own their counters)
  const int parity = (int)(seq & 1ull);
  for (int i = 0; i < nl; ++i) {
    SlabRank<T>& r = R[i];
    const tvdk_slab_comm_t* c = r.cm;
    CommPush a;
    memset(&a, 0, sizeof a);
    const long long plane_bytes = r.g.s0 * (long long)sizeof(T);
    const int zb = r.g.z_begin, ze = r.g.z_end;
    for (const HaloSpec& h : halos[i]) {
      const unsigned char* base = reinterpret_cast<const unsigned char*>(h.field);
      if (h.n_up > 0 && c->rank + 1 < c->world) {  // my last owned planes -> the ghost planes below the upper neighbour's range
        if (a.nseg >= CommPush::kMaxSeg) return fail(TVDK_ERR_INVALID, "internal: too many halo segments");
        CommPush::Seg& s = a.seg[a.nseg++];
        s.src = base + (long long)(ze - h.n_up) * plane_bytes;
        s.dst = reinterpret_cast<unsigned char*>(peer_ptr(c, base, c->rank + 1)) + (long long)(TVDK_SLAB_GL - h.n_up) * plane_bytes;
        s.bytes = (long long)h.n_up * plane_bytes;
      }
      if (h.n_down > 0 && c->rank > 0) {  // my first owned planes -> the ghost planes above the lower neighbour's range
        if (a.nseg >= CommPush::kMaxSeg) return fail(TVDK_ERR_INVALID, "internal: too many halo segments");
        CommPush::Seg& s = a.seg[a.nseg++];
        s.src = base + (long long)zb * plane_bytes;
        s.dst = reinterpret_cast<unsigned char*>(peer_ptr(c, base, c->rank - 1)) +
                (long long)(TVDK_SLAB_GL + c->nown[c->rank - 1]) * plane_bytes;
        s.bytes = (long long)h.n_down * plane_bytes;
      }
    }
    a.sums = r.hdr->local_sums;
    a.nred = nred;
    a.world = c->world;
    for (int s = 0; s < c->world; ++s)
      a.mail[s] = &reinterpret_cast<SlabHeader*>(c->arena[s])->slot[parity][c->rank];
    a.flag = seq;
    a.counter = &r.hdr->counter;
    TVDK_TRY(backend::set_device(r.device));
    ++g_launches;
    TVDK_TRY(backend::comm_push(a, r.stream));
  }
  if (nl > 1) {
    std::vector<void*> streams(nl);
    std::vector<int> devs(nl);
    for (int i = 0; i < nl; ++i) {
      streams[i] = R[i].stream;
      devs[i] = R[i].device;
    }
    TVDK_TRY(backend::comm_sync_local(nl, streams.data(), devs.data()));
  }
  for (int i = 0; i < nl; ++i) {
    SlabRank<T>& r = R[i];
    CommFin a;
    memset(&a, 0, sizeof a);
    a.mail = &r.hdr->slot[parity][0];
    a.world = r.cm->world;
    a.nred = nred;
    a.flag = seq;
    a.timeout_ns = (long long)(r.cm->timeout_ms > 0 ? r.cm->timeout_ms : 20000) * 1000000LL;
    a.error = &r.hdr->error_seq;
    a.fin = fins[i];
    a.skip_mask = skip_mask;
    a.stats = r.st;
    TVDK_TRY(backend::set_device(r.device));
    ++g_launches;
    TVDK_TRY(backend::comm_finalize(a, r.stream));
  }
  return 0;
}

enum { SKIP_DONE = 1, SKIP_CG = 2, SKIP_JAC = 4 };

// The workspace layout must not depend on the rank (peers address each other's fields by arena offset): size the
// scratch for the LARGEST slab -- every rank's array has GL + max nown + GH planes.
static Geom capacity_geom(Geom g) {
  g.z_begin = TVDK_SLAB_GL;
  g.z_end = g.n0 - TVDK_SLAB_GH;
  g.zchunk = g.z_end - g.z_begin;
  return g;
}
// ... and for the launch shape with the MOST thread blocks any rank can get (the planner picks 8..64 planes per block
// depending on how many planes a rank owns; 8 is its minimum), so that every rank carves the same scratch
template <typename T, int D>
static size_t slab_partial_capacity(const Geom& g, const Tune& tune) {
  const Geom cg = capacity_geom(g);
  Tune fine = tune;
  if (fine.zchunk == 0) fine.zchunk = 8;
  return std::max(partial_capacity<T, D>(cg, tune), partial_capacity<T, D>(cg, fine));
}

template <typename T, int D>
static size_t slab_workspace_bytes_impl(const tvdk_problem_t* p) {
  Geom g = capacity_geom(make_geom(p->geom));
  Tune tune = merge_tune(p->tile_x, p->tile_y, p->zchunk);
  const size_t cap = slab_partial_capacity<T, D>(g, tune);
  const bool cg = p->solver_kind == TVDK_SOLVER_CG;
  const bool fusedj1 = !cg && p->solver_maxiter == 1 && p->solver_tol < 0.0;
  const int n_fields = fusedj1 ? 1 + D : (cg ? 4 : 3);
  const size_t field_bytes = align_up((size_t)g.sb * g.nb * sizeof(T), 256);
  return scratch_bytes(cap, g.nb) + n_fields * field_bytes + 256;
}

template <typename T, int D>
static int solve_slab_impl(int nlocal, const tvdk_problem_t* const* problems, const tvdk_slab_comm_t* const* comms,
                           void* const* streams, const int* devices, int* iters_launched) {
  const tvdk_problem_t* p0 = problems[0];
  const bool iso = p0->isotropic != 0;
  const bool cg = p0->solver_kind == TVDK_SOLVER_CG;
  const bool fusedj1 = !cg && p0->solver_maxiter == 1 && p0->solver_tol < 0.0;
  const bool tolmode = !cg && p0->solver_tol >= 0.0;
  const Tune tune = merge_tune(p0->tile_x, p0->tile_y, p0->zchunk);
  std::vector<SlabRank<T>> R(nlocal);
  for (int i = 0; i < nlocal; ++i) {
    const tvdk_problem_t* p = problems[i];
    if (p->isotropic != p0->isotropic || p->solver_kind != p0->solver_kind || p->solver_maxiter != p0->solver_maxiter ||
        p->solver_tol != p0->solver_tol || p->max_num_iter != p0->max_num_iter || p->eps != p0->eps)
      return fail(TVDK_ERR_INVALID, "all ranks must pass the same solver parameters");
    const size_t need = slab_workspace_bytes_impl<T, D>(p);
    if (p->workspace_bytes < need)
      return fail(TVDK_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", need, p->workspace_bytes);
    if (!p->stats) return fail(TVDK_ERR_INVALID, "stats must be non-NULL");
    SlabRank<T>& r = R[i];
    r.p = p;
    r.cm = comms[i];
    r.stream = streams ? streams[i] : nullptr;
    r.device = devices ? devices[i] : -1;
    r.g = make_geom(p->geom);
    const size_t cap = slab_partial_capacity<T, D>(r.g, tune);
    r.sc = carve_scratch(p->workspace, cap, r.g.nb);
    r.hdr = reinterpret_cast<SlabHeader*>(comms[i]->arena[comms[i]->rank]);
    r.k = make_coefs<T>(&p->mu, &p->omega, &p->ell, D);
    r.f = reinterpret_cast<const T*>(p->f);
    const size_t fe = align_up((size_t)r.g.sb * r.g.nb * sizeof(T), 256) / sizeof(T);
    T* wsf = reinterpret_cast<T*>(r.sc.fields);
    int next = 0;
    auto take = [&](int n) {
      T* q = wsf + (size_t)next * fe;
      next += n;
      return q;
    };
    r.ubuf[0] = reinterpret_cast<T*>(p->u);
    r.tbuf[0] = reinterpret_cast<T*>(p->t);
    r.tcs[0] = p->t_comp_stride;
    r.tcs[1] = (long long)fe;
    r.ubuf[1] = nullptr;
    r.tbuf[1] = nullptr;
    r.RB = r.XA = r.XB = r.X = r.P[0] = r.P[1] = nullptr;
    if (fusedj1) {
      r.ubuf[1] = take(1);
      r.tbuf[1] = take(D);
    } else if (cg) {
      r.X = take(1);
      r.RB = take(1);
      r.P[0] = take(1);
      r.P[1] = take(1);
    } else {
      r.RB = take(1);
      r.XA = take(1);
      r.XB = take(1);
    }
    r.st = p->stats;
    r.cur = 0;
  }
  const bool het = R[0].k.sc.hetero != 0;
  const bool het0 = R[0].k.omega.p != nullptr || R[0].k.ell.p != nullptr;
  const int m = p0->solver_maxiter;

  auto each = [&](auto&& fn) -> int {
    for (int i = 0; i < nlocal; ++i) {
      TVDK_TRY(backend::set_device(R[i].device));
      TVDK_TRY(fn(R[i]));
    }
    return 0;
  };
  // kernels write their raw per-rank sums into the header; the scalar update happens in comm_finalize
  auto local_fin = [&](SlabRank<T>& r) {
    Fin f = no_fin();
    f.sum_out = r.hdr->local_sums;
    return f;
  };
  auto fins_of = [&](auto&& mk) {
    std::vector<Fin> v(nlocal);
    for (int i = 0; i < nlocal; ++i) v[i] = mk(R[i]);
    return v;
  };
  auto halos_of = [&](auto&& mk) {
    std::vector<std::vector<HaloSpec>> v(nlocal);
    for (int i = 0; i < nlocal; ++i) v[i] = mk(R[i]);
    return v;
  };
  auto plain_fin = [&](int kind, double tol = 0.0) {
    return fins_of([&](SlabRank<T>& r) {
      Fin f = no_fin();
      f.kind = kind;
      f.stats = r.st;
      f.tol = tol;
      return f;
    });
  };
  auto field_halo = [&](auto&& get, int up, int down) {
    return halos_of([&](SlabRank<T>& r) { return std::vector<HaloSpec>{{get(r), up, down}}; });
  };

  TVDK_TRY(each([&](SlabRank<T>& r) {
    TVDK_TRY(backend::memset_async(r.st, 0, sizeof(tvdk_stat_t) * (size_t)r.g.nb, r.stream));
    TVDK_TRY(backend::memset_async(r.sc.counters, 0, sizeof(unsigned) * (size_t)r.g.nb, r.stream));
    return backend::memset_async(r.sc.ndone, 0, sizeof(int), r.stream);
  }));

  // handshake only: every rank has entered this solve (its uploads into f / u / t are stream-ordered before
  // this point), so nobody pushes ghosts into an arena that is still being filled
  TVDK_TRY(comm_step(R, std::vector<std::vector<HaloSpec>>(nlocal), 0, std::vector<Fin>(nlocal, no_fin()), 0));

  // ||img|| over all ranks (split_bregman_tvd.py:58), together with the ghosts of the initial state:
  // u: 1 plane up / 2 down, t: 1 each way (what the fused kernel reads; a superset for the other routes)
  TVDK_TRY(each([&](SlabRank<T>& r) {
    typename KDot<T, D, 0>::Args a;
    memset(&a, 0, sizeof a);
    a.x = r.f;
    a.y = r.f;
    a.fin = local_fin(r);
    return run<KDot<T, D, 0>, T, D>(a, r.g, tune, false, r.sc, r.stream);
  }));
  TVDK_TRY(comm_step(R, halos_of([&](SlabRank<T>& r) {
                       std::vector<HaloSpec> h{{r.f, TVDK_SLAB_GL, TVDK_SLAB_GH}, {r.ubuf[0], 1, 2}};
                       for (int d = 0; d < D; ++d) h.push_back({r.tbuf[0] + d * r.tcs[0], 1, 1});
                       return h;
                     }),
                     1, plain_fin(FIN_NRMF), 0));

  auto outer_fins = [&](int it, int buf_id) {
    return fins_of([&](SlabRank<T>& r) {
      Fin fn = no_fin();
      fn.kind = FIN_OUTER;
      fn.stats = r.st;
      fn.eps = p0->eps;
      fn.history = r.p->history;
      fn.iter = it;
      fn.nb = r.g.nb;
      fn.ndone = r.sc.ndone;
      fn.buf_id = buf_id;
      return fn;
    });
  };
  // shrink + Bregman update + in-place update of u (K3), then the ghosts of the new u and t
  auto shrink_step = [&](int it, auto&& unew, bool use_sel) -> int {
    TVDK_TRY(each([&](SlabRank<T>& r) {
      return dispatch2(iso, het, [&](auto I, auto H) {
        typedef KShrink<T, D, decltype(I)::value, decltype(H)::value> K;
        typename K::Args a;
        memset(&a, 0, sizeof a);
        a.k = r.k;
        a.unew[0] = unew(r, 0); a.unew[1] = unew(r, 1); a.unew[2] = unew(r, 2);
        a.use_sel = use_sel ? 1 : 0;
        a.cg_mode = cg ? 1 : 0;
        a.u_old = r.ubuf[0]; a.u_out = r.ubuf[0];
        a.t_in = r.tbuf[0]; a.t_out = r.tbuf[0]; a.tcs = r.tcs[0];
        a.skip = r.st; a.fin = local_fin(r);
        return run<K, T, D>(a, r.g, tune, false, r.sc, r.stream);
      });
    }));
    return comm_step(R, halos_of([&](SlabRank<T>& r) {
                       std::vector<HaloSpec> h{{r.ubuf[0], 1, 1}};
                       for (int d = 0; d < D; ++d) h.push_back({r.tbuf[0] + d * r.tcs[0], 1, 1});
                       return h;
                     }),
                     1, outer_fins(it, 0), SKIP_DONE);
  };

  const int check_every = p0->check_every > 0 ? p0->check_every : 8;
  int launched = 0;
  bool all_done = false;
  for (int it = 0; it < p0->max_num_iter && !all_done; ++it) {
    if (fusedj1) {
      // K1: one launch per outer iteration, ping-pong (u, t); its outputs' ghosts: u 1 up / 2 down, t 1 each way
      const int cur = R[0].cur;
      TVDK_TRY(each([&](SlabRank<T>& r) {
        return run_fj1<T, D>(iso, het, r.g, tune, r.k, r.f, r.ubuf[cur], r.ubuf[1 - cur], r.tbuf[cur], r.tbuf[1 - cur], r.tcs[cur],
                             r.tcs[1 - cur], r.st, local_fin(r), r.sc, r.stream);
      }));
      TVDK_TRY(comm_step(R, halos_of([&](SlabRank<T>& r) {
                           std::vector<HaloSpec> h{{r.ubuf[1 - cur], 1, 2}};
                           for (int d = 0; d < D; ++d) h.push_back({r.tbuf[1 - cur] + d * r.tcs[1 - cur], 1, 1});
                           return h;
                         }),
                         1, outer_fins(it, 1 - cur), SKIP_DONE));
      for (auto& r : R) r.cur = 1 - cur;
    } else if (cg) {
      // K2 (rhs, r0 = rhs - A u, x = u) ; m x [K5 (p, p.Ap) ; K6 (x, r, r.r)] ; K3 -- scipy's recurrence via cg.py:50-56
      TVDK_TRY(each([&](SlabRank<T>& r) {
        return dispatch2(iso, het, [&](auto I, auto H) {
          typedef KRhs<T, D, decltype(I)::value, RHS_CGINIT, decltype(H)::value> K;
          typename K::Args a;
          memset(&a, 0, sizeof a);
          a.k = r.k; a.f = r.f; a.t = r.tbuf[0]; a.tcs = r.tcs[0]; a.u = r.ubuf[0];
          a.x = r.X; a.r = r.RB; a.skip = r.st; a.fin = local_fin(r);
          return run<K, T, D>(a, r.g, tune, false, r.sc, r.stream);
        });
      }));
      // (the upper ghost of x = u is what the shrink kernel reads if CG stops before its first update)
      TVDK_TRY(comm_step(R, halos_of([&](SlabRank<T>& r) { return std::vector<HaloSpec>{{r.RB, 1, 1}, {r.X, 0, 1}}; }), 2,
                         plain_fin(FIN_CGINIT, p0->solver_tol), SKIP_DONE));
      const int sd = (int)(sizeof(tvdk_stat_t) / sizeof(double)), si = (int)(sizeof(tvdk_stat_t) / sizeof(int));
      for (int i = 0; i < m; ++i) {
        TVDK_TRY(each([&](SlabRank<T>& r) {
          DRef beta;
          beta.p = &r.st[0].beta; beta.stride = sd;
          IRef frst;
          frst.p = &r.st[0].cg_first; frst.stride = si;
          return run_cgpq<T, D>(het, r.g, tune, r.k, r.P[(i & 1) ^ 1], r.RB, r.P[i & 1], beta, frst, 0, r.st, local_fin(r), r.sc, r.stream);
        }));
        TVDK_TRY(comm_step(R, field_halo([&](SlabRank<T>& r) { return r.P[i & 1]; }, 1, 1), 1, plain_fin(FIN_CGPQ),
                           SKIP_DONE | SKIP_CG));
        TVDK_TRY(each([&](SlabRank<T>& r) {
          return dispatch2(het, false, [&](auto H, auto) {
            typedef KCg<T, D, CG_UPDATE, decltype(H)::value> K;
            typename K::Args a;
            memset(&a, 0, sizeof a);
            a.k = r.k; a.p_in = r.P[i & 1]; a.x = r.X; a.r = r.RB;
            a.coef.p = &r.st[0].alpha; a.coef.stride = sd;
            a.skip = r.st; a.fin = local_fin(r);
            return run<K, T, D>(a, r.g, tune, false, r.sc, r.stream);
          });
        }));
        // (x's upper ghost is read by the shrink kernel only: after the last update, or whenever a tolerance may stop CG)
        const bool xg = i == m - 1 || p0->solver_tol > 0.0;
        TVDK_TRY(comm_step(R, halos_of([&](SlabRank<T>& r) {
                             std::vector<HaloSpec> h{{r.RB, 1, 1}};
                             if (xg) h.push_back({r.X, 0, 1});
                             return h;
                           }),
                           1, plain_fin(FIN_CGRR), SKIP_DONE | SKIP_CG));
      }
      TVDK_TRY(shrink_step(it, [&](SlabRank<T>& r, int which) -> const T* { return which == 0 ? r.X : nullptr; }, false));
    } else {
      // Jacobi(K >= 2 or tol): K2 (rhs + sweep 1) ; K - 1 x K4 ; K3 (jacobi.py:109-131)
      std::vector<Fin> jf = plain_fin(tolmode ? FIN_JTOL : FIN_NONE, p0->solver_tol);
      for (auto& f : jf) {
        f.buf_id = 1;
        f.first_sweep = 1;
      }
      TVDK_TRY(each([&](SlabRank<T>& r) {
        return dispatch2(iso, het, [&](auto I, auto H) {
          typedef KRhs<T, D, decltype(I)::value, RHS_SWEEP, decltype(H)::value> K;
          typename K::Args a;
          memset(&a, 0, sizeof a);
          a.k = r.k; a.f = r.f; a.t = r.tbuf[0]; a.tcs = r.tcs[0]; a.u = r.ubuf[0];
          a.rhs = r.RB; a.x = r.XA; a.skip = r.st; a.fin = local_fin(r);
          return run<K, T, D>(a, r.g, tune, false, r.sc, r.stream);
        });
      }));
      TVDK_TRY(comm_step(R, field_halo([&](SlabRank<T>& r) { return r.XA; }, 1, 1), 2, jf, SKIP_DONE));
      bool in_a = true;  // the latest sweep wrote XA
      for (int s = 2; s <= m; ++s) {
        const int out_id = in_a ? 2 : 1;
        TVDK_TRY(each([&](SlabRank<T>& r) {
          // (the sums of the tol branch are only formed when a tolerance is set)
          return run_sweep<T, D>(het0, r.g, tune, r.k, in_a ? r.XA : r.XB, r.RB, in_a ? r.XB : r.XA, r.st,
                                 tolmode ? local_fin(r) : no_fin(), r.sc, r.stream);
        }));
        std::vector<Fin> sf = plain_fin(tolmode ? FIN_JTOL : FIN_NONE, p0->solver_tol);
        for (auto& f : sf) {
          f.buf_id = out_id;
          f.first_sweep = 0;
        }
        TVDK_TRY(comm_step(R, field_halo([&](SlabRank<T>& r) { return in_a ? r.XB : r.XA; }, 1, 1), 2, sf,
                           SKIP_DONE | SKIP_JAC));
        in_a = !in_a;
      }
      if (tolmode) {
        // jacobi.py:118-131: on convergence the PREVIOUS iterate is kept; stats.jac_sel names the buffer
        // (0: u itself -- its upper ghost is current since the last shrink)
        TVDK_TRY(shrink_step(it, [&](SlabRank<T>& r, int which) -> const T* {
          return which == 0 ? r.ubuf[0] : (which == 1 ? r.XA : r.XB);
        }, true));
      } else {
        TVDK_TRY(shrink_step(it, [&](SlabRank<T>& r, int which) -> const T* {
          return which == 0 ? (in_a ? r.XA : r.XB) : nullptr;
        }, false));
      }
    }
    ++launched;
    // split_bregman_tvd.py:173-175: a BLOCKING read every check_every iterations -- every rank reads the same
    // counter value after the same iteration, so all ranks stop enqueueing at the same point
    if (p0->eps >= 0.0 && (launched % check_every) == 0 && it + 1 < p0->max_num_iter) {
      int v = 0;
      TVDK_TRY(backend::set_device(R[0].device));
      TVDK_TRY(backend::read_int(R[0].sc.ndone, R[0].stream, &v));
      if (v >= R[0].g.nb) all_done = true;
    }
  }
  if (fusedj1) {  // bring the iterate (and state) back into p->u / p->t
    TVDK_TRY(each([&](SlabRank<T>& r) {
      typename KSelectCopy<T, D>::Args a;
      memset(&a, 0, sizeof a);
      a.src = r.ubuf[1]; a.dst = r.ubuf[0]; a.stats = r.st; a.want = 1; a.fin = no_fin();
      TVDK_TRY((run<KSelectCopy<T, D>, T, D>(a, r.g, tune, false, r.sc, r.stream)));
      for (int d = 0; d < D; ++d) {
        typename KSelectCopy<T, D>::Args b;
        memset(&b, 0, sizeof b);
        b.src = r.tbuf[1] + d * r.tcs[1]; b.dst = r.tbuf[0] + d * r.tcs[0]; b.stats = r.st; b.want = 1; b.fin = no_fin();
        TVDK_TRY((run<KSelectCopy<T, D>, T, D>(b, r.g, tune, false, r.sc, r.stream)));
      }
      return 0;
    }));
  }
  if (iters_launched) *iters_launched = launched;
  return 0;
}

template <typename T, int D>
static int slab_exchange_impl(int nlocal, const tvdk_geom_t* const* geoms, const tvdk_slab_comm_t* const* comms, int nfields,
                              const void* const* const* fields, int n_up, int n_down, void* const* streams,
                              const int* devices) {
  std::vector<SlabRank<T>> R(nlocal);
  std::vector<std::vector<HaloSpec>> halos(nlocal);
  std::vector<Fin> fins(nlocal, no_fin());
  for (int i = 0; i < nlocal; ++i) {
    R[i].p = nullptr;
    R[i].cm = comms[i];
    R[i].stream = streams ? streams[i] : nullptr;
    R[i].device = devices ? devices[i] : -1;
    R[i].g = make_geom(*geoms[i]);
    R[i].hdr = reinterpret_cast<SlabHeader*>(comms[i]->arena[comms[i]->rank]);
    R[i].st = nullptr;
    for (int k = 0; k < nfields; ++k) halos[i].push_back({fields[i][k], n_up, n_down});
  }
  // handshake first (see solve_slab_impl), then the planes
  TVDK_TRY(comm_step(R, std::vector<std::vector<HaloSpec>>(nlocal), 0, fins, 0));
  return comm_step(R, halos, 0, fins, 0);
}

}  // namespace tvdk

size_t tvdk_slab_header_bytes(void) { return sizeof(tvdk::SlabHeader); }

size_t tvdk_slab_workspace_bytes(const tvdk_problem_t* p) {
  if (!p || check_geom(&p->geom)) return 0;
  size_t r = 0;
  TVDK_DISPATCH(p->geom, r = (slab_workspace_bytes_impl<T, D>(p)));
  return r;
}

int tvdk_solve_slab(int nlocal, const tvdk_problem_t* const* problems, const tvdk_slab_comm_t* const* comms,
                    void* const* streams, const int* devices, int* iters_launched) {
  if (nlocal < 1 || nlocal > TVDK_MAX_SLABS || !problems || !comms) return fail(TVDK_ERR_INVALID, "bad rank group");
  for (int i = 0; i < nlocal; ++i) {
    TVDK_TRY(check_slab(problems[i], comms[i]));
    if (problems[i]->geom.dim != problems[0]->geom.dim || problems[i]->geom.dtype != problems[0]->geom.dtype)
      return fail(TVDK_ERR_INVALID, "all ranks must share dim and dtype");
    if (problems[i]->max_num_iter < 0) return fail(TVDK_ERR_INVALID, "max_num_iter must be >= 0");
  }
  int rc = 0;
  TVDK_RANGE("tvdk_solve_slab");
  g_alternate = 0;  // (several ranks share this thread's toggle: every rank must see the same direction per launch)
  TVDK_DISPATCH(problems[0]->geom, rc = (solve_slab_impl<T, D>(nlocal, problems, comms, streams, devices, iters_launched)));
  return rc;
}

int tvdk_slab_exchange(int nlocal, const tvdk_geom_t* const* geoms, const tvdk_slab_comm_t* const* comms, int nfields,
                       const void* const* const* fields, int n_up, int n_down, void* const* streams, const int* devices) {
  if (nlocal < 1 || nlocal > TVDK_MAX_SLABS || !geoms || !comms || !fields || nfields < 1)
    return fail(TVDK_ERR_INVALID, "bad rank group");
  if (n_up < 0 || n_up > TVDK_SLAB_GL || n_down < 0 || n_down > TVDK_SLAB_GH)
    return fail(TVDK_ERR_INVALID, "halo depth out of range (up <= %d, down <= %d)", TVDK_SLAB_GL, TVDK_SLAB_GH);
  for (int i = 0; i < nlocal; ++i) {
    TVDK_TRY(check_geom(geoms[i]));
    if (!comms[i] || !comms[i]->seq) return fail(TVDK_ERR_INVALID, "comm / comm.seq is NULL");
  }
  int rc = 0;
  TVDK_DISPATCH(*geoms[0], rc = (slab_exchange_impl<T, D>(nlocal, geoms, comms, nfields, fields, n_up, n_down, streams, devices)));
  return rc;
}

int tvdk_slab_error(const void* arena, void* stream, int64_t* seq_out) {
  if (!arena || !seq_out) return fail(TVDK_ERR_INVALID, "NULL pointer");
  long long v = 0;
  TVDK_TRY(backend::read_i64(&reinterpret_cast<const tvdk::SlabHeader*>(arena)->error_seq, stream, &v));
  *seq_out = v;
  return 0;
}
