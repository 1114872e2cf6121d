"""Cases for the TMA-staged 3-D anisotropic fused Jacobi-1 kernel (KFusedJ1T), shared by the emulator (CPU) and
the GPU test modules: shapes around tile / chunk seams, thin and tiny volumes (every plane on a boundary), batches
(the tensor maps address plane b * n0 + z), fp32, forced tiny tiles."""

import numpy as np

from golden.generate_golden import phantom


def tma_cases():
    f8, f4 = np.float64, np.float32
    return [
        # shape, batch, tile (x, y, planes per CTA), dtype, outer iterations
        ((11, 19, 70), (), (0, 0, 0), f8, 6),
        ((11, 19, 70), (), (8, 6, 3), f8, 6),
        ((11, 19, 70), (), (8, 6, 3), f4, 6),
        ((13, 33, 131), (), (16, 8, 4), f8, 5),
        ((9, 17, 66), (2,), (8, 5, 4), f8, 5),
        ((9, 17, 66), (3,), (0, 0, 0), f4, 5),
        ((2, 1, 7), (), (0, 0, 0), f8, 4),
        ((1, 3, 5), (), (0, 0, 0), f8, 3),
        ((9, 5, 6), (), (0, 0, 2), f8, 6),
        ((5, 16, 64), (), (0, 0, 1), f8, 4),
        ((7, 15, 62), (), (0, 0, 0), f8, 4),
        # the halo row / vector of an interior tile IS the last image row / column (ny = k * ih + 1, nx = x0 + ntx * V)
        ((6, 31, 126), (), (0, 0, 0), f8, 4),
        ((6, 46, 125), (), (0, 0, 0), f8, 4),
        ((6, 31, 124), (), (0, 0, 0), f4, 4),
        ((6, 29, 62), (), (0, 0, 0), f8, 4),
        ((6, 22, 92), (), (0, 0, 0), f8, 4),
        ((6, 34, 124), (), (0, 0, 0), f4, 4),
        ((7, 16, 30), (), (8, 6, 3), f8, 4),
        ((7, 16, 46), (), (8, 6, 3), f4, 4),
    ]


def run_tma_case(da, oracle, shape, batch, tile, dtype, nit, iso=False):
    img = phantom(tuple(shape) + tuple(batch), 13, dtype)
    kw = dict(mu=0.1, omega=1.3, ell=0.9, dim=3, max_num_iter=nit, isotropic=iso)
    hist = []
    ref = oracle.split_bregman_tvd(img, solver=oracle.Jacobi(maxiter=1), history=hist, **kw)
    # (fused = 2 forces the isotropic TMA variant, which is not the default route: the march kernel measured faster)
    out, info = da.split_bregman_tvd(img, solver=da.Jacobi(maxiter=1), tile=tile, return_info=True, fused=2 if iso else True, **kw)
    return out, ref, info, hist


def stencil_cases():
    """3-D CG / Jacobi(K) through the TMA-staged single-field stencil kernel (KStencilT: tiles WITHOUT halo threads, so the
    tile edges coincide with the image edges for exact multiples): shape, tile, dtype."""
    f8, f4 = np.float64, np.float32
    return [
        ((6, 16, 64), (0, 0, 0), f8), ((6, 17, 65), (0, 0, 0), f8), ((7, 24, 97), (0, 0, 0), f8), ((6, 24, 128), (0, 0, 0), f4),
        ((6, 25, 129), (0, 0, 0), f4), ((11, 19, 70), (8, 6, 3), f8), ((11, 19, 70), (8, 6, 3), f4), ((2, 1, 7), (0, 0, 0), f8),
        ((9, 5, 6), (0, 0, 2), f8), ((5, 13, 40), (0, 0, 1), f8),
    ]


def run_stencil_case(da, oracle, shape, tile, dtype, solver):
    img = phantom(tuple(shape), 17, dtype)
    kw = dict(mu=0.1, omega=1.2, ell=0.8, dim=3, max_num_iter=4, isotropic=(solver == "cg"))
    mk = (lambda m: m.CG(maxiter=4, tol=0.0)) if solver == "cg" else (lambda m: m.Jacobi(maxiter=3))
    ref = oracle.split_bregman_tvd(img, solver=mk(oracle), **kw)
    out = da.split_bregman_tvd(img, solver=mk(da), tile=tile, **kw)
    return out, ref


def tma2d_cases():
    """2-D anisotropic Jacobi(1) through the TMA-staged kernel (KFusedJ1T<T, false, 2>: the march "plane" is one image row,
    a CTA a strip of columns): shape, batch, tile (threads, -, rows per CTA), dtype, outer iterations."""
    f8, f4 = np.float64, np.float32
    return [
        ((37, 150), (), (32, 0, 5), f8, 6),
        ((37, 150), (), (32, 0, 5), f4, 6),
        ((37, 150), (2,), (32, 0, 7), f8, 5),
        ((41, 300), (), (0, 0, 0), f8, 6),
        ((70, 300), (), (64, 0, 0), f8, 7),
        ((33, 126), (), (0, 0, 0), f8, 4),   # one tile whose halo vector is the last image column
        ((33, 127), (), (0, 0, 0), f8, 4),
        ((33, 128), (), (0, 0, 0), f8, 4),
        ((33, 252), (), (0, 0, 0), f8, 4),
        ((33, 124), (), (0, 0, 0), f4, 4),
        ((33, 125), (), (0, 0, 0), f4, 4),
        ((20, 62), (), (32, 0, 3), f8, 4),
        ((3, 40), (), (0, 0, 0), f8, 6),
        ((1, 40), (), (0, 0, 0), f8, 4),
        ((2, 7), (), (0, 0, 0), f8, 5),
        ((5, 9), (3,), (0, 0, 1), f4, 3),
        ((64, 1), (), (0, 0, 0), f8, 3),
    ]


def run_tma2d_case(da, oracle, shape, batch, tile, dtype, nit):
    img = phantom(tuple(shape) + tuple(batch), 13, dtype)
    kw = dict(mu=0.1, omega=1.3, ell=0.9, dim=2, max_num_iter=nit, isotropic=False)
    hist = []
    ref = oracle.split_bregman_tvd(img, solver=oracle.Jacobi(maxiter=1), history=hist, **kw)
    out, info = da.split_bregman_tvd(img, solver=da.Jacobi(maxiter=1), tile=tile, return_info=True, fused=4, **kw)
    march = da.split_bregman_tvd(img, solver=da.Jacobi(maxiter=1), tile=tile, fused=3, **kw)  # K1, one iteration per launch
    return out, ref, info, hist, march
