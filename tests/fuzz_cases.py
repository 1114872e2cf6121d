"""Seeded random problems over every route of split_bregman_tvd (shape, tile, batch, dtype, isotropy,
per-voxel maps, solver family, kernel route, eps): shared by the CPU thread-stepper and the GPU suites."""

import numpy as np

from golden.generate_golden import coeff_field, phantom


def draw(seed: int, n: int):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        dim = int(rng.choice([2, 3]))
        if dim == 2:
            shape = (int(rng.integers(1, 45)), int(rng.integers(2, 200)))
            tile = (int(rng.choice([0, 32, 64])), 0, int(rng.choice([0, 3, 5, 8, 17])))
        else:
            shape = (int(rng.integers(1, 14)), int(rng.integers(1, 22)), int(rng.integers(2, 75)))
            tile = (int(rng.choice([0, 8, 16])), int(rng.choice([0, 6, 8, 11])), int(rng.choice([0, 2, 3, 5])))
        c = dict(
            dim=dim, shape=shape, tile=tile,
            batch=() if rng.random() < 0.7 else (int(rng.integers(1, 4)),),
            dtype=np.float64 if rng.random() < 0.75 else np.float32,
            iso=bool(rng.random() < 0.5), het=bool(rng.random() < 0.35),
            kind=str(rng.choice(["j1", "jK", "cg0", "cgtol", "jtol", "mg"], p=[0.25, 0.2, 0.25, 0.1, 0.1, 0.1])),
            fused=int(rng.choice([1, 2, 0, 3])), nit=int(rng.integers(1, 8)), K=int(rng.integers(2, 12)),
            seed=int(rng.integers(0, 1000)), omega=float(rng.choice([0.5, 1.0, 2.0])), ell=float(rng.choice([0.5, 1.0])),
            het_ell=bool(rng.random() < 0.4), eps=None if rng.random() < 0.8 else 5e-3,
        )
        if c["kind"] == "mg" and (min(shape) < 2 or c["het"]):
            continue
        out.append(c)
    return out


def solver(api, c):
    K = c["K"]
    return {
        "j1": lambda: api.Jacobi(maxiter=1), "jK": lambda: api.Jacobi(maxiter=K),
        "cg0": lambda: api.CG(maxiter=min(K, 6), tol=0.0), "cgtol": lambda: api.CG(maxiter=30, tol=1e-3),
        "jtol": lambda: api.Jacobi(maxiter=15, tol=1e-2),
        "mg": lambda: api.MG(depth=0, smoother_iterations=min(K, 4), maxiter=2),
    }[c["kind"]]()


def run(da, oracle, c):
    dt = c["dtype"]
    img = phantom(c["shape"] + c["batch"], c["seed"], dt)
    kw = dict(mu=0.1, omega=c["omega"], ell=c["ell"])
    if c["het"]:
        kw["mu"] = coeff_field(c["shape"], 8, 0.05, 0.1, 0.1).astype(dt)
        kw["omega"] = coeff_field(c["shape"], 9, 0.5, 2.0).astype(dt)
        if c["het_ell"]:
            kw["ell"] = coeff_field(c["shape"], 10, 0.5, 1.5).astype(dt)
    common = dict(dim=c["dim"], max_num_iter=c["nit"], isotropic=c["iso"], eps=c["eps"], **kw)
    ref = oracle.split_bregman_tvd(img, solver=solver(oracle, c), **common)
    extra = {} if c["kind"] == "mg" else dict(tile=c["tile"], fused=c["fused"])
    out = da.split_bregman_tvd(img, solver=solver(da, c), **extra, **common)
    return out, ref
