"""Stage the reference's hot-path files for the GPU box (TEST / BASELINE INFRASTRUCTURE ONLY).

    python oracle/make_ref.py            # called by __graft_entry__.build()

Copies the ten files ``oracle/ref_loader.HOT_PATH_FILES`` VERBATIM from ``/root/reference/src/darsia``
(or ``$DARSIA_REFERENCE``) into ``oracle/_ref/darsia/`` together with a SHA-256 manifest.  ``oracle/_ref/``
is git-ignored (never enters the history) but not gpurun-ignored, so it travels with the snapshot
the way a built ``.so`` does: ``bench.py --impl reference`` can then time the UNMODIFIED reference on
the GPU box's host cores, where ``/root/reference`` does not exist.  A no-op (keeping whatever is
already staged) when the reference tree is not present.
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(HERE))

from oracle.ref_loader import HOT_PATH_FILES, STAGED  # noqa: E402


def stage(src=None) -> bool:
    src = src or os.environ.get("DARSIA_REFERENCE") or "/root/reference/src/darsia"
    if not all(os.path.isfile(os.path.join(src, rel)) for rel in HOT_PATH_FILES):
        print(f"make_ref: no reference tree at {src}; keeping {STAGED} as it is")
        return os.path.isdir(STAGED)
    manifest = {}
    for rel in HOT_PATH_FILES:
        dst = os.path.join(STAGED, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(STAGED, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "sha256": manifest}, fh, indent=1)
    print(f"make_ref: staged {len(manifest)} reference files into {STAGED}")
    return True


if __name__ == "__main__":
    stage()
