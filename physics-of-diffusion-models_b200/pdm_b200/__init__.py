"""pdm_b200 -- B200-native engine for the empirical (ideal) denoiser and its thermodynamic statistics.

Python host side above the C ABI of ``include/pdm_b200.h``.  The reference-facing drop-in modules live next
to this package (``utils/``, ``diffusion/``); they call into ``PosteriorEngine``.
"""
from ._cabi import PdmError, LIB_PATH  # noqa: F401
from .engine import EmpiricalDataset, PosteriorEngine, EngineConfig, STAT_KEYS, pow2_scale_for  # noqa: F401
from .sampling import IdealSampler, step_coefficients  # noqa: F401

__all__ = ["PdmError", "LIB_PATH", "EmpiricalDataset", "PosteriorEngine", "EngineConfig", "STAT_KEYS",
           "pow2_scale_for", "IdealSampler", "step_coefficients"]
