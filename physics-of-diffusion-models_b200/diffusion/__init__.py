"""Drop-in ``diffusion`` package: the ideal-denoiser path of the reference (``Scheduler``,
``DDPM``/``DDPMPredictions``, ``DDPMTrue``) on the B200 engine.  Neural models, trainer, sampler and the
other noise schedules are out of scope and come from the reference checkout named
by ``PDM_REFERENCE_ROOT`` when it is set (pdm_b200/overlay.py).
"""
from __future__ import annotations

import importlib

from pdm_b200.overlay import extend_package_path

from .scheduler import Scheduler, alpha_bar_from_log_temp, cast_log_temp  # noqa: F401
from .ddpm import DDPM, DDPMPredictions, DDPMTrue  # noqa: F401

_HAS_REFERENCE = extend_package_path(__path__, "diffusion")

_LAZY = {
    "ddpm_from_config": ".ddpm",
    "DDPMTrainer": ".ddpm_trainer",
    "DDPMSampler": ".ddpm_sampling",
    "get_samples": ".ddpm_sampling",
}


def __getattr__(name: str):
    if name in _LAZY:
        if not _HAS_REFERENCE:
            raise AttributeError(f"diffusion.{name} is provided by the reference checkout; set PDM_REFERENCE_ROOT")
        value = getattr(importlib.import_module(_LAZY[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
