"""The ``darsia.Image`` in/out contract, duck-typed.

``import darsia`` is not possible where this package is built, so the
boundary accepts any object with an ``img`` ndarray attribute and a ``copy()`` method --
exactly what ``TVD._tvd_image`` uses (``restoration/tvd.py:117-129``).  When DarSIA is
installed, ``darsia.Image`` instances satisfy this test.
"""

from __future__ import annotations

import copy as _copy

import numpy as np


def is_image_like(obj) -> bool:
    return hasattr(obj, "img") and isinstance(getattr(obj, "img"), np.ndarray) and callable(getattr(obj, "copy", None))


class Image:
    """Minimal stand-in for ``darsia.Image`` (``image/image.py``): ``img`` + metadata + ``copy``."""

    def __init__(self, img: np.ndarray, **metadata) -> None:
        self.img = img
        self._metadata = dict(metadata)

    def copy(self) -> "Image":
        return _copy.deepcopy(self)  # image/image.py:363-370

    def metadata(self) -> dict:
        return _copy.copy(self._metadata)
