"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

Test infrastructure only (see ``oracle/__init__.py``).  The reference's
``utils``/``diffusion`` packages import four third-party packages that are not
installed here and are not used by the hot path (``diffusers``,
``denoising_diffusion_pytorch``, ``torchmetrics``, ``torch_ema``) plus
``matplotlib`` in the scripts.  Empty stand-ins are registered in
``sys.modules`` so the reference's own hot-path code runs untouched.

The reference tree does not exist on the GPU box: callers must check
``reference_available()`` and the ``-m gpu`` tests never call this module.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PDM_REFERENCE_ROOT", "/root/reference")

_STUBS = {
    "diffusers": ("DDPMPipeline", "UNet2DModel", "DDPMScheduler"),
    "diffusers.models": (),
    "diffusers.models.attention_processor": ("AttnProcessor2_0",),
    "denoising_diffusion_pytorch": ("Unet",),
    "torchmetrics": (),
    "torchmetrics.image": (),
    "torchmetrics.image.fid": ("FrechetInceptionDistance",),
    "torch_ema": ("ExponentialMovingAverage",),
    "matplotlib": (),
    "matplotlib.pyplot": (),
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "stats.py"))


def _install_stubs() -> None:
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave like a package so sub-imports resolve
        for a in attrs:
            setattr(mod, a, type(a, (), {}))
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)


class ReferenceModules:
    """Handle on the reference's packages, imported under their own names."""

    def __init__(self) -> None:
        if not reference_available():
            raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
        _install_stubs()
        # The reference's packages are top-level `utils`, `diffusion`, `config`.
        # Make sure no same-named package (e.g. our drop-in mirror) shadows them.
        for name in list(sys.modules):
            root = name.split(".")[0]
            if root in ("utils", "diffusion", "config"):
                f = getattr(sys.modules[name], "__file__", "") or ""
                if not f.startswith(REFERENCE_ROOT):
                    del sys.modules[name]
        self._saved_path = list(sys.path)
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            self.utils = importlib.import_module("utils")
            self.distance = importlib.import_module("utils.distance")
            self.stats = importlib.import_module("utils.stats")
            self.metric_utils = importlib.import_module("utils.metric_utils")
            self.synthetic = importlib.import_module("utils.synthetic_datasets")
            self.diffusion = importlib.import_module("diffusion")
            self.scheduler = importlib.import_module("diffusion.scheduler")
            self.true_model = importlib.import_module("diffusion.ddpm.true_model")
            self.sampling = importlib.import_module("diffusion.ddpm_sampling")
        finally:
            sys.path[:] = self._saved_path

    def unload(self) -> None:
        """Drop the reference's modules so our same-named mirror can be imported."""
        for name in list(sys.modules):
            root = name.split(".")[0]
            if root in ("utils", "diffusion", "config"):
                f = getattr(sys.modules[name], "__file__", "") or ""
                if f.startswith(REFERENCE_ROOT):
                    del sys.modules[name]


def load_reference() -> ReferenceModules:
    return ReferenceModules()
