"""``darsia.H1_regularization`` on the GPU ("next" row 1).

Reference: ``restoration/h1_regularization.py:15-146``:  min_u 1/2 ||u - img||^2_omega +
1/2 ||grad u||^2_mu, i.e. ONE inner solve ``omega u - mu laplace(u) = omega img`` started from
``img`` with the same ``Jacobi`` / ``CG`` kernels as the TVD path; channels are solved
separately (``h1_regularization.py:54-67``).
"""

from __future__ import annotations

from typing import Optional

import numpy as np

from ..image import is_image_like
from ..utils.linear_solvers import Jacobi, Solver


def _h1_array(img: np.ndarray, mu, omega, dim: int, solver: Solver) -> np.ndarray:
    if not np.issubdtype(img.dtype, np.floating):
        raise TypeError("float images only")
    solver.update_params(mass_coeff=omega, diffusion_coeff=mu, dim=dim)  # :46-50
    if img.ndim == dim or solver.tol is None:
        # trailing axes are independent problems inside the solver (one batched launch sequence)
        return solver(x0=img, rhs=omega * img)
    # with a tolerance every channel has its own stopping iteration (:54-67 loops over channels)
    out = np.empty_like(img)
    for idx in np.ndindex(img.shape[dim:]):
        sel = (slice(None),) * dim + idx
        chan = np.ascontiguousarray(img[sel])
        out[sel] = solver(x0=chan, rhs=omega * chan)
    return out


def H1_regularization(img, mu: float, omega: float = 1.0, dim: int = 2, solver: Optional[Solver] = None):
    """Inline application of H1 regularization (``h1_regularization.py:110-146``); ndarray or Image."""
    if solver is None:
        solver = Jacobi()
    if isinstance(img, np.ndarray):
        return _h1_array(img, mu, omega, dim, solver)
    elif is_image_like(img):
        out = img.copy()
        out.img = _h1_array(img.img, mu, omega, dim, solver)
        return out
    raise TypeError(f"Type {type(img)} not supported.")
