"""CPU restatement of scikit-image's Chambolle TV denoiser -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  ``darsia.tvd(img)`` defaults to ``method="chambolle"``, which the reference
forwards to ``skimage.restoration.denoise_tv_chambolle`` (``restoration/tvd.py:77-83``).  The
algorithm therefore lives in a third-party dependency that is neither under ``/root/reference``
nor installed in this image (``pyproject.toml`` lists a bare ``"scikit-image"``, no pin), and the
reference holds no golden vector for it.  What follows restates the published algorithm of
``skimage/restoration/_denoise.py::_denoise_tv_chambolle_nd`` (unchanged across scikit-image
0.19 - 0.25 apart from the ``n_iter_max`` -> ``max_num_iter`` rename) from its description; it
could not be checked against scikit-image itself here, so every parity claim that rests on this
file is capped at "partial".

The iteration (Chambolle 2004, projection algorithm), all axes spatial, ``tau = 1 / (2 ndim)``::

    d    = div p            d[k] = sum_a ( p_a[k - e_a] (k_a >= 1)  -  p_a[k] )        (0 in iteration 0)
    out  = image + d
    E    = sum d^2
    g_a  = out[k + e_a] - out[k]   (0 in the last slice of axis a)
    nrm  = sqrt(sum_a g_a^2);  E += weight * sum nrm;  E /= image.size
    p_a  = (p_a - tau g_a) / (1 + tau / weight * nrm)
    stop after iteration i > 0 when |E_prev - E| < eps * E_0, else E_prev = E

The value returned is ``out`` of the last iteration that ran (formed BEFORE that iteration's update
of ``p``).
"""

from __future__ import annotations

import numpy as np


def denoise_tv_chambolle(image: np.ndarray, weight: float = 0.1, eps: float = 2.0e-4, max_num_iter: int = 200,
                         return_info: bool = False):
    """``skimage.restoration.denoise_tv_chambolle(image, weight, eps, max_num_iter)`` with
    ``channel_axis=None`` (every axis spatial), float input (float32 stays float32)."""
    image = np.asarray(image)
    if image.dtype.kind != "f":
        raise TypeError("float images only (scikit-image would rescale integers with img_as_float)")
    if image.dtype not in (np.float32, np.float64):
        image = image.astype(np.float64)
    ndim = image.ndim
    p = np.zeros((ndim,) + image.shape, dtype=image.dtype)
    g = np.zeros_like(p)
    d = np.zeros_like(image)
    out = image
    energies = []
    i = 0
    e_init = e_prev = 0.0
    while i < max_num_iter:
        if i > 0:
            d = -p.sum(0)  # minus the divergence's "own" term ...
            for ax in range(ndim):  # ... plus the lower neighbour along every axis
                sd = [slice(None)] * ndim
                sp = [slice(None)] * (ndim + 1)
                sd[ax] = slice(1, None)
                sp[ax + 1] = slice(0, -1)
                sp[0] = ax
                d[tuple(sd)] += p[tuple(sp)]
            out = image + d
        else:
            out = image
        e = (d**2).sum()
        for ax in range(ndim):
            sg = [slice(None)] * (ndim + 1)
            sg[ax + 1] = slice(0, -1)
            sg[0] = ax
            g[tuple(sg)] = np.diff(out, axis=ax)
        norm = np.sqrt((g**2).sum(axis=0))[np.newaxis, ...]
        e += weight * norm.sum()
        tau = 1.0 / (2.0 * ndim)
        norm *= tau / weight
        norm += 1.0
        p -= tau * g
        p /= norm
        e /= float(image.size)
        energies.append(float(e))
        if i == 0:
            e_init = e
            e_prev = e
        else:
            if np.abs(e_prev - e) < eps * e_init:
                break
            e_prev = e
        i += 1
    if return_info:
        return out, dict(iterations=len(energies), energies=energies)
    return out
