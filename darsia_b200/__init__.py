"""darsia_b200 -- B200-native (sm_100a) split-Bregman TV denoising, drop-in for the hot path of
pmgbergen/DarSIA: ``tvd`` / ``TVD`` / ``split_bregman_tvd`` with ``Jacobi`` / ``CG`` inner solves.

The names mirror DarSIA's flattened namespace (``src/darsia/__init__.py:41-48,112-118``).
Everything runs in hand-written CUDA kernels behind a C ABI (``include/tvdk.h``); PyTorch is
used for device memory, streams and ``torch.distributed`` only.  There is no CPU fallback.
"""

from .image import Image, is_image_like  # noqa: F401
from .utils.linear_solvers import CG, MG, Jacobi, Solver  # noqa: F401
from .restoration.split_bregman_tvd import split_bregman_tvd  # noqa: F401
from .restoration.tvd import TVD, tvd  # noqa: F401
from .restoration.h1_regularization import H1_regularization  # noqa: F401
from .engine import Engine, SolveInfo, SolverSpec  # noqa: F401
from .series import frame_range, tvd_series, tvd_series_stream  # noqa: F401
from .slab import slab_range, split_bregman_tvd_slab  # noqa: F401
from .slab_p2p import SlabGroup, split_bregman_tvd_slab_p2p  # noqa: F401

__version__ = "0.1.0"
