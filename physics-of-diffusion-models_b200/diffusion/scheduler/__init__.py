"""Noise schedules.  ``Scheduler`` (with the ideal denoiser) and the linear-beta schedule live here; the
remaining schedules are scalar math outside the hot path and load from the reference checkout when
``PDM_REFERENCE_ROOT`` is set (their ``from .scheduler import Scheduler`` then binds to this package's)."""
from __future__ import annotations

import importlib

from pdm_b200.overlay import extend_package_path

from .scheduler import (  # noqa: F401
    Scheduler as Scheduler,
    log_temp_from_alpha_bar as log_temp_from_alpha_bar,
    alpha_bar_from_log_temp as alpha_bar_from_log_temp,
    cast_log_temp as cast_log_temp,
)
from .linear import LinearBetaScheduler as LinearBetaScheduler  # noqa: F401

_HAS_REFERENCE = extend_package_path(__path__, "diffusion", "scheduler")

_LAZY = {
    "CosineScheduler": ".cosine", "LogSNRScheduler": ".log_snr",
    "InterpolatedDiscreteTimeScheduler": ".interpolated", "CustomScheduler": ".custom",
    "EntropyScheduler": ".entropy", "FromDiffusersScheduler": ".diffusers", "MetricScheduler": ".metric",
    "scheduler_from_config": ".from_config",
}


def __getattr__(name: str):
    if name in _LAZY:
        if not _HAS_REFERENCE:
            raise AttributeError(f"diffusion.scheduler.{name} is provided by the reference checkout; set PDM_REFERENCE_ROOT")
        value = getattr(importlib.import_module(_LAZY[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
