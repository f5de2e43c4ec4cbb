"""Drop-in ``utils`` package: the hot-path modules of the reference's ``utils`` re-implemented on the
B200 engine (distance, stats, metric_utils, synthetic_datasets).  Everything else the reference's package
exports (data loading, FID, LeNet, config helpers -- out of scope) resolves to the
reference's own files when ``PDM_REFERENCE_ROOT`` points at a checkout (see pdm_b200/overlay.py and
INTEGRATION.md); without it those names raise an informative AttributeError.
"""
from __future__ import annotations

import importlib

from pdm_b200.overlay import extend_package_path

from .distance import (  # noqa: F401
    compute_gram_matrix as compute_gram_matrix,
    compute_pw_dist_sqr as compute_pw_dist_sqr,
    norm_sqr as norm_sqr,
)
from .synthetic_datasets import (  # noqa: F401
    generate_simplex as generate_simplex,
    generate_cross_polytope as generate_cross_polytope,
    sample_on_hypersphere as sample_on_hypersphere,
    generate_gaussian as generate_gaussian,
    generate_dataset as generate_dataset,
)
from .stats import (  # noqa: F401
    compute_stats as compute_stats,
    compute_stats_batch as compute_stats_batch,
    extrapolate_entropy as extrapolate_entropy,
    compute_metric_stats as compute_metric_stats,
    compute_metric_stats_batch as compute_metric_stats_batch,
    compute_model_metric_stats as compute_model_metric_stats,
    compute_thermo_stats as compute_thermo_stats,       # extension: legacy notebook schema (log_Z, U, full_U, var_H)
)
from .metric_utils import (  # noqa: F401
    compute_metric_scalar as compute_metric_scalar,
    compute_metric_matrix as compute_metric_matrix,
    compute_rescaled_metric_matrix as compute_rescaled_metric_matrix,
)

_HAS_REFERENCE = extend_package_path(__path__, "utils")

# names the reference's utils/__init__.py re-exports from modules we do not provide (utils/__init__.py:1-46)
_REFERENCE_EXPORTS = {
    "data": ("get_default_num_workers", "get_dataset", "get_data_tensor", "get_data_generator",
             "compute_dataset_average", "to_uint8"),
    "utils": ("dict_map", "append_dict", "add_dict", "extend_dict", "batch_jacobian", "get_diffusers_pipeline",
              "load_config", "with_config", "interp1d", "get_default_device", "parse_value"),
    "fid": ("extract_features_statistics", "compute_fid", "get_compute_fid"),
    "lenet": ("LeNet", "train_lenet"),
}


def __getattr__(name: str):
    for module, names in _REFERENCE_EXPORTS.items():
        if name in names:
            if not _HAS_REFERENCE:
                raise AttributeError(
                    f"utils.{name} lives in the reference's utils/{module}.py, which this drop-in does not "
                    "re-implement; set PDM_REFERENCE_ROOT to a checkout of the reference")
            value = getattr(importlib.import_module(f"{__name__}.{module}"), name)
            globals()[name] = value
            return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
