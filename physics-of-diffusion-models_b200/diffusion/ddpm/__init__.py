"""DDPM wrappers of the ideal-denoiser path.  ``DDPMUnet`` / ``DDPMDiffusers`` / ``ddpm_from_config`` are
neural-network code outside the hot path and load from the reference checkout (PDM_REFERENCE_ROOT)."""
from __future__ import annotations

import importlib

from pdm_b200.overlay import extend_package_path

from .ddpm import DDPM as DDPM, DDPMPredictions as DDPMPredictions  # noqa: F401
from .true_model import DDPMTrue as DDPMTrue  # noqa: F401

_HAS_REFERENCE = extend_package_path(__path__, "diffusion", "ddpm")

_LAZY = {"DDPMUnet": ".unet", "set_processor_recursively": ".unet", "DDPMDiffusers": ".diffusers_model",
         "ddpm_from_config": ".from_config"}


def __getattr__(name: str):
    if name in _LAZY:
        if not _HAS_REFERENCE:
            raise AttributeError(f"diffusion.ddpm.{name} is provided by the reference checkout; set PDM_REFERENCE_ROOT")
        value = getattr(importlib.import_module(_LAZY[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
