"""Drop-in for the reference's ``utils/distance.py`` on the B200 engine.

Same names, arguments and return conventions as the reference (utils/distance.py:5-21): tensors of any
trailing shape are flattened per row, ``y=None`` means self-distance, the result is a dense float32 matrix
on the input's device evaluated as ``||x||^2 - 2 x.y + ||y||^2`` in that order and never clamped.
The arithmetic runs in the CUDA library (pdm_b200); there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import weakref
from collections import OrderedDict
from typing import Optional

import torch
from torch import Tensor

from pdm_b200 import EmpiricalDataset, PosteriorEngine
from pdm_b200.engine import default_backend

_backend = None


def _be():
    global _backend
    if _backend is None:
        _backend = default_backend()
    return _backend


def _to_engine(x: Tensor) -> Tensor:
    be = _be()
    return x.reshape(x.shape[0], -1).to(device=be.device, dtype=torch.float32).contiguous()


def compute_gram_matrix(x: Tensor, y: Tensor) -> Tensor:
    """x @ y.T (utils/distance.py:5-6) through the exact fp32 CUDA-core contraction."""
    be = _be()
    xf, yf = _to_engine(x), _to_engine(y)
    out = be.weighted_mean_exact(xf, yf.t().contiguous())
    return out.to(x.device)


def norm_sqr(x: Tensor) -> Tensor:
    """Row-wise squared norm of a 2-D tensor (utils/distance.py:9-10)."""
    return _be().row_norms(_to_engine(x)).to(x.device)


_DATASETS: "OrderedDict[tuple, tuple]" = OrderedDict()


def _dataset_for(ref: Tensor) -> EmpiricalDataset:
    """The prepared form of ``ref`` (row norms, operand split) is kept for the last few tensors seen, keyed by storage
    address / shape / version and tied to the tensor by a weak reference: loops that call compute_pw_dist_sqr(x, y)
    against the same y (the k-NN searches of the scripts) prepare it once."""
    key = (ref.data_ptr(), tuple(ref.shape), ref.dtype, str(ref.device), ref._version)
    hit = _DATASETS.get(key)
    if hit is not None and hit[0]() is ref:
        _DATASETS.move_to_end(key)
        return hit[1]
    ds = EmpiricalDataset(ref, backend=_be())
    _DATASETS[key] = (weakref.ref(ref), ds)
    for k in [k for k, (r, _) in _DATASETS.items() if r() is None]:
        del _DATASETS[k]
    while len(_DATASETS) > 2:
        _DATASETS.popitem(last=False)
    return ds


def compute_pw_dist_sqr(x: Tensor, y: Optional[Tensor] = None) -> Tensor:
    """Dense pairwise squared distances (utils/distance.py:13-21)."""
    ref = x if y is None else y
    return PosteriorEngine(_dataset_for(ref)).pairwise_sqdist(x).to(x.device)
