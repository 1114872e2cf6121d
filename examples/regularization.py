"""The reference's ``examples/regularization.py`` on the B200 path: heterogeneous anisotropic TV denoising of a
3-D volume with a per-voxel weight and ``Jacobi(maxiter=20)``, then H1 regularisation -- the same calls,
``darsia`` replaced by ``darsia_b200`` (reference lines cited).

    python examples/regularization.py [volume.npy]        # default: a synthetic 128 x 160 x 192 CT-like volume

The reference loads ``examples/images/dicom_3d.npy`` through ``darsia.imread`` and plots slices with
matplotlib; image I/O and plotting are outside this package's scope, so the volume comes from a ``.npy``
file (or is synthesised) and the result is summarised on stdout.
"""

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout
import darsia_b200 as darsia  # noqa: E402

if len(sys.argv) > 1:
    vol = np.load(sys.argv[1]).astype(np.float64)
else:
    rng = np.random.default_rng(0)
    z, y, x = np.indices((128, 160, 192), sparse=True)
    vol = 0.2 + 0.3 * (((z // 32) + (y // 40) + (x // 48)) % 3) + 0.1 * rng.standard_normal((128, 160, 192))
    vol = np.clip(vol, 0.0, 1.0)
img = darsia.Image(vol)  # duck-typed stand-in for darsia.Image: .img + .copy()

# Make regularization parameter heterogenous (for illustration purposes)      (examples/regularization.py:13-15)
mu = np.ones_like(img.img)
mu[:, 0:100, :] = 0.5

# (the first GPU call of a process pays for the CUDA context and for loading the kernels -- about 3 s; keep it
#  out of the timings below)
darsia.tvd(np.zeros((8, 8, 32)) + 0.5, method="heterogeneous bregman", dim=3, max_num_iter=1)

t0 = time.perf_counter()
img_regularized_tvd = darsia.tvd(  # examples/regularization.py:18-30
    img=img,
    method="heterogeneous bregman",
    isotropic=False,
    weight=mu,
    omega=0.1,
    max_num_iter=30,
    eps=1e-6,
    dim=3,
    verbose=False,
    solver=darsia.Jacobi(maxiter=20),
    # solver = darsia.CG(maxiter = 20, tol = 1e-3)
)
t1 = time.perf_counter()
img_regularized_h1 = darsia.H1_regularization(  # examples/regularization.py:33-40
    img=img,
    mu=1,
    omega=1,
    dim=3,
    solver=darsia.Jacobi(maxiter=30, tol=1e-6),
    # solver = darsia.MG(3)
)
t2 = time.perf_counter()


def tv(a):
    return sum(np.abs(np.diff(a, axis=ax)).sum() for ax in range(a.ndim))


print(f"volume {vol.shape}: total variation {tv(vol):.4g}")
print(f"TVD  (heterogeneous bregman, Jacobi(20), 30 outer): {t1 - t0:.3f} s, total variation {tv(img_regularized_tvd.img):.4g}")
print(f"H1   (Jacobi(30, tol=1e-6)):                        {t2 - t1:.3f} s, total variation {tv(img_regularized_h1.img):.4g}")
