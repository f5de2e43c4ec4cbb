"""Overlay of the drop-in packages on a checkout of the reference.

The reference's scripts import top-level packages ``utils``, ``diffusion`` and ``config`` and bind the
hot-path functions by name at import time (utils/stats.py:9, diffusion/scheduler/scheduler.py:10), so the
drop-in has to BE those packages.  This directory provides ``utils`` and ``diffusion`` with the hot-path
modules re-implemented on the CUDA engine; everything else (data loading, FID, config, trainers, the other
schedulers) is out of scope and is taken unchanged from the reference when ``PDM_REFERENCE_ROOT`` points at
a checkout: its directories are appended to the package ``__path__`` so that sub-modules we do not provide
resolve to the reference's files, while ``from .scheduler import Scheduler`` / ``from utils import
compute_pw_dist_sqr`` inside those files pick up ours.
"""
from __future__ import annotations

import os
import sys


def reference_root() -> str | None:
    root = os.environ.get("PDM_REFERENCE_ROOT")
    if root and os.path.isdir(os.path.join(root, "utils")):
        return root
    return None


def extend_package_path(pkg_path: list, *relative: str) -> bool:
    """Append <reference>/<relative...> to a package __path__ (after our own directory)."""
    root = reference_root()
    if root is None:
        return False
    p = os.path.join(root, *relative)
    if os.path.isdir(p) and p not in pkg_path:
        pkg_path.append(p)
    if root not in sys.path:
        sys.path.append(root)          # lets `import config` resolve to the reference's package
    return True
