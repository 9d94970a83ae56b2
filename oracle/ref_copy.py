"""Where the UNMODIFIED reference sources can be found at run time — TEST / BENCH INFRASTRUCTURE.

In the build container the reference is mounted read-only at /root/reference.  The GPU box has no such mount; `gpurun` ships
/root/repo only.  SURVEY.md Appendix B: copy the handful of files the oracle pin and the reference arms need VERBATIM into the
git-ignored ``baseline/_ref/`` (it travels with the snapshot, it is never committed, nothing under ``lemevit_b200/`` reads it):

    benchmark.py                                                   the reference's own benchmark driver (external clock)
    models/__init__.py, models/lemevit.py                          the classification model
    semantic_segmentation/mmseg/models/backbones/lemevit.py        the mmseg backbone copy

``ensure()`` (called by ``__graft_entry__.build()``) refreshes the copy whenever /root/reference is present; ``root()`` returns
the directory to load from (the mount if present, else the copy, else None).
"""
from __future__ import annotations

import filecmp
import os
import shutil

MOUNT = os.environ.get("LEMEVIT_REFERENCE_ROOT", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COPY = os.path.join(REPO, "baseline", "_ref")
FILES = [
    "benchmark.py",
    "models/__init__.py",
    "models/lemevit.py",
    "semantic_segmentation/mmseg/models/backbones/lemevit.py",
]


def ensure() -> str | None:
    """Refresh baseline/_ref from the mounted reference (no-op when the mount is absent).  Returns the copy's path or None."""
    if not os.path.isfile(os.path.join(MOUNT, FILES[0])):
        return COPY if os.path.isfile(os.path.join(COPY, FILES[0])) else None
    for rel in FILES:
        src, dst = os.path.join(MOUNT, rel), os.path.join(COPY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    return COPY


def root() -> str | None:
    for cand in (MOUNT, COPY):
        if os.path.isfile(os.path.join(cand, "models", "lemevit.py")):
            return cand
    return None
