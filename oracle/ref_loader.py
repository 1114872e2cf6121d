"""Run the UNMODIFIED DarSIA reference hot path (TEST / BASELINE INFRASTRUCTURE ONLY).

``import darsia`` fails in this image (matplotlib, scikit-image, ... are absent), so the ten
hot-path source files are executed VERBATIM inside a stub ``darsia`` namespace plus a stub
``skimage`` exposing the ``img_as_*`` identities the path uses on float arrays.  The only
shims are ``scipy.sparse.linalg.cg(tol=...)`` -> ``rtol=...`` (SciPy >= 1.14 renamed the
argument; same meaning) and ``np.lib.pad = np.pad`` (NumPy 2 removed the alias
that ``utils/linear_solvers/mg.py:233`` still uses).

Where the files come from, in this order:

1. ``$DARSIA_REFERENCE`` if set,
2. ``/root/reference/src/darsia`` (the build container),
3. ``oracle/_ref/darsia`` -- a git-ignored staging copy made by ``oracle/make_ref.py`` during
   ``__graft_entry__.build()`` so that the files travel to the GPU box with the snapshot (the
   box has no ``/root/reference``).  Nothing of it enters the repository's history.

Used by ``tests/golden/generate_golden.py`` (fixtures), ``tests/test_oracle_vs_reference_live.py``
and the ``--impl reference`` / ``cpu_baseline`` legs of ``bench.py``.  The product never imports it.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "darsia")

# the hot-path files, in dependency order
HOT_PATH_FILES = [
    "utils/array_slice.py",
    "utils/derivatives.py",
    "utils/dtype.py",
    "utils/linear_solvers/solver.py",
    "utils/linear_solvers/jacobi.py",
    "utils/linear_solvers/cg.py",
    "utils/linear_solvers/mg.py",
    "restoration/tvd.py",
    "restoration/split_bregman_tvd.py",
    "restoration/h1_regularization.py",
]


def reference_dir():
    """Directory holding the reference's ``src/darsia`` tree, or None."""
    cands = [os.environ.get("DARSIA_REFERENCE"), "/root/reference/src/darsia", STAGED]
    for c in cands:
        if c and all(os.path.isfile(os.path.join(c, rel)) for rel in HOT_PATH_FILES):
            return c
    return None


def available() -> bool:
    return reference_dir() is not None


def load_reference(ref_dir=None):
    """Execute the reference hot-path files verbatim under a stub namespace; returns the stub
    ``darsia`` module exposing ``tvd``, ``TVD``, ``split_bregman_tvd``, ``Jacobi``, ``CG``, ``MG``, ..."""
    import scipy.sparse.linalg as spla

    ref_dir = ref_dir or reference_dir()
    if ref_dir is None:
        raise FileNotFoundError("DarSIA reference sources not found (no /root/reference, no oracle/_ref)")

    skimage = types.ModuleType("skimage")

    def _as(dtype):
        def f(x):
            x = np.asarray(x)
            if not np.issubdtype(x.dtype, np.floating):
                raise TypeError("stub skimage: float arrays only")
            return x if dtype is None else x.astype(dtype, copy=False)

        return f

    skimage.img_as_float = _as(None)
    skimage.img_as_float32 = _as(np.float32)
    skimage.img_as_float64 = _as(np.float64)
    sys.modules["skimage"] = skimage

    darsia = types.ModuleType("darsia")

    class Image:  # duck-typed stand-in for isinstance checks in tvd.py
        def __init__(self, img):
            self.img = img

        def copy(self):
            import copy

            return copy.deepcopy(self)

    darsia.Image = Image
    sys.modules["darsia"] = darsia

    if not getattr(spla.cg, "_tol_shim", False):
        _cg = spla.cg

        def cg(A, b, x0=None, *, tol=None, rtol=1e-5, atol=0.0, maxiter=None, M=None, callback=None):
            if tol is not None:
                rtol = tol
            return _cg(A, b, x0=x0, rtol=rtol, atol=atol, maxiter=maxiter, M=M, callback=callback)

        cg._tol_shim = True
        spla.cg = cg

    if not hasattr(np.lib, "pad"):  # NumPy >= 2 removed the alias mg.py:233 uses
        np.lib.pad = np.pad

    for rel in HOT_PATH_FILES:
        name = "darsia._ref_" + rel.replace("/", "_")[:-3]
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_dir, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for k, v in vars(mod).items():
            if not k.startswith("_") and getattr(v, "__module__", None) == name:
                setattr(darsia, k, v)
    darsia.__ref_dir__ = ref_dir
    return darsia
