from .split_bregman_tvd import split_bregman_tvd  # noqa: F401
from .tvd import TVD, tvd  # noqa: F401
from .h1_regularization import H1_regularization  # noqa: F401
