"""CPU oracle for the split-Bregman TVD hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: the
only legitimate importers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as
the checker / the timed CPU arm, never as the thing shipped.  The product
(``darsia_b200``) never imports this package and fails loudly when its CUDA
extension is missing.

Parity status: the DarSIA reference holds no golden vectors for this path
(SURVEY.md section 4), so the restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF, generated in the build container by executing the reference's
hot-path files verbatim (``tests/golden/generate_golden.py``) and committed as
fixtures under ``tests/golden/*.npz``.
"""

from .chambolle_oracle import denoise_tv_chambolle  # noqa: F401  (parity unpinned, see its header)
from .tvd_oracle import (  # noqa: F401
    CG,
    H1_regularization,
    Jacobi,
    MG,
    Solver,
    backward_diff,
    cg_matvec,
    energy_functional,
    forward_diff,
    jacobi_sweep,
    laplace,
    mg_prolongation,
    mg_restriction,
    neighbor_accumulation,
    shrink_anisotropic,
    shrink_isotropic,
    split_bregman_tvd,
    tvd,
    TVD,
)
