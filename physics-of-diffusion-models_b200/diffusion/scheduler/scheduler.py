"""``Scheduler`` base class with the closed-form ideal denoiser on the B200 engine.

Mirror of the reference's diffusion/scheduler/scheduler.py:13-69: same free functions, same abstract
interface (``log_temp_from_tau`` / ``tau_from_log_temp``), same ``add_noise`` and the same
``true_posterior_mean_x0(xt, tau, data)`` signature and result (float32, shaped like ``xt``, on ``xt``'s
device).  The reference evaluates  h = 1/2 ||xt - sqrt(ab) y||^2, p = softmax(-h / (1 - ab)), p @ data  with
a fresh sqrt(ab)-scaled copy of the whole dataset per call (:64); here the identity
||x - a y||^2 = a^2 ||x/a - y||^2 turns it into the VE form with queries x/sqrt(ab) at temperature
T = (1 - ab)/ab, so the resident dataset (norms, fp16 split, transposed split) is reused untouched.
"""
from __future__ import annotations

import os
import weakref
from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Optional

import torch
from torch import Tensor

from pdm_b200 import EmpiricalDataset, PosteriorEngine
from pdm_b200.engine import default_backend


def log_temp_from_alpha_bar(alpha_bar: Tensor) -> Tensor:
    return (1 - alpha_bar).log() - alpha_bar.log()


def alpha_bar_from_log_temp(log_temp: Tensor) -> Tensor:
    return torch.sigmoid(-log_temp)


def cast_log_temp(log_temp: Tensor, target: Tensor) -> Tensor:
    return log_temp.view(-1, *[1] * (target.ndim - 1))


_DENOISER_ENGINES: "OrderedDict[tuple, tuple]" = OrderedDict()
_MAX_RESIDENT_DATASETS = 4


def _engine_for_data(data: Tensor) -> PosteriorEngine:
    """One resident dataset per training-data tensor: keyed by storage address, shape, dtype, device and version, and
    tied to the tensor itself by a weak reference -- a different tensor that happens to be allocated at a freed
    tensor's address (loops over equally shaped synthetic datasets) never hits a stale entry.  Least recently used out."""
    key = (data.data_ptr(), tuple(data.shape), data.dtype, str(data.device), data._version)
    hit = _DENOISER_ENGINES.get(key)
    if hit is not None and hit[0]() is data:
        _DENOISER_ENGINES.move_to_end(key)
        return hit[1]
    eng = PosteriorEngine(EmpiricalDataset(data, backend=default_backend()))
    _DENOISER_ENGINES[key] = (weakref.ref(data), eng)
    _DENOISER_ENGINES.move_to_end(key)
    for k in [k for k, (ref, _) in _DENOISER_ENGINES.items() if ref() is None]:       # their tensors are gone
        del _DENOISER_ENGINES[k]
    while len(_DENOISER_ENGINES) > _MAX_RESIDENT_DATASETS:
        _DENOISER_ENGINES.popitem(last=False)
    return eng


def _query_shards():
    """(group, rank, world) over which ``true_posterior_mean_x0`` splits the rows of a batch when PDM_SHARD_QUERIES=1 and
    torch.distributed is initialised (mode 2: the dataset is replicated, every rank evaluates a
    contiguous slice of the queries and the slices are all-gathered -- the reference's own DDPMSampler loop, run with the
    same seed on every rank, then samples on all GPUs of the node unchanged)."""
    if os.environ.get("PDM_SHARD_QUERIES", "0") != "1":
        return None, 0, 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None, 0, 1
    return dist.group.WORLD, dist.get_rank(), dist.get_world_size()


class Scheduler(ABC):
    def __init__(self):
        pass

    @abstractmethod
    def log_temp_from_tau(self, tau: Tensor) -> Tensor:
        pass

    @abstractmethod
    def tau_from_log_temp(self, log_temp: Tensor) -> Tensor:
        pass

    def alpha_bar_from_tau(self, tau: Tensor) -> Tensor:
        return alpha_bar_from_log_temp(self.log_temp_from_tau(tau))

    def add_noise(self, x0: Tensor, tau: Optional[Tensor] = None) -> tuple[Tensor, Tensor, Tensor]:
        """xt = sqrt(ab) x0 + sqrt(1 - ab) eps with tau ~ U[0,1) per sample unless given (scheduler.py:40-45)."""
        tau = torch.rand((len(x0),), device=x0.device) if tau is None else tau
        alpha_bar = cast_log_temp(self.alpha_bar_from_tau(tau), x0)
        eps = torch.randn_like(x0)
        return tau, eps, alpha_bar.sqrt() * x0 + eps * (1 - alpha_bar).sqrt()

    def true_score(self, xt: Tensor, tau: Tensor, train_data: Tensor) -> Tensor:
        """Ideal score (x0_hat sqrt(ab) - xt)/(1 - ab).  The reference's version (scheduler.py:47-56, never
        called by its own code) materialises an (N, B, ...) tensor; this one goes through the fused path."""
        alpha_bar = cast_log_temp(self.alpha_bar_from_tau(tau), xt)
        x0_hat = self.true_posterior_mean_x0(xt, tau, train_data)
        return (x0_hat * alpha_bar.sqrt() - xt.float()) / (1 - alpha_bar)

    @torch.autocast(device_type="cuda", enabled=False)
    def true_posterior_mean_x0(self, xt: Tensor, tau: Tensor, data: Tensor) -> Tensor:
        x = xt.float()
        b = x.shape[0]
        alpha_bar = self.alpha_bar_from_tau(tau).float().reshape(-1)
        if alpha_bar.numel() not in (1, b):
            raise ValueError(f"tau must be scalar-like or hold one value per sample, got {alpha_bar.numel()} for batch {b}")
        eng = _engine_for_data(data)
        if torch.is_grad_enabled() and (x.requires_grad or alpha_bar.requires_grad):
            return _IdealDenoiser.apply(x, alpha_bar, eng)       # scripts/optimize_schedule.py differentiates this
        return _denoise(eng, x, alpha_bar)


def _denoise(eng: PosteriorEngine, x: Tensor, alpha_bar: Tensor) -> Tensor:
    ab = alpha_bar.detach().expand(x.shape[0])
    temp_rows = ((1 - ab) / ab).clamp_min(1e-30)
    group, rank, world = _query_shards()
    if world == 1 or x.shape[0] < world:
        return eng.posterior_mean(x.detach(), temp_rows, post=ab.rsqrt()).to(x.device).view_as(x)
    import torch.distributed as dist
    b = x.shape[0]
    per = (b + world - 1) // world
    lo, hi = min(b, rank * per), min(b, (rank + 1) * per)
    mine = eng.posterior_mean(x.detach()[lo:hi], temp_rows[lo:hi], post=ab.rsqrt()[lo:hi])
    send = torch.zeros(per, mine.shape[1], dtype=mine.dtype, device=mine.device)
    send[:hi - lo] = mine
    recv = torch.empty(world * per, mine.shape[1], dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:b].to(x.device).view_as(x)


class _IdealDenoiser(torch.autograd.Function):
    """x0_hat(xt, alpha_bar) with its vector-Jacobian product on the engine.  In the VE variables
    q = xt / sqrt(ab), T = (1 - ab)/ab the posterior is p_j ~ exp(-||q - y_j||^2 / 2T), so
        d x0_hat / d q = Cov_p(y) / T,        d x0_hat / d T = Cov_p(E, y) / T^2,
    chained through dq/dxt = ab^-1/2, dq/dab = -xt ab^-3/2 / 2, dT/dab = -1/ab^2.  ``data`` is a constant."""

    @staticmethod
    def forward(ctx, x: Tensor, alpha_bar: Tensor, eng: PosteriorEngine) -> Tensor:
        out = _denoise(eng, x, alpha_bar)
        ctx.eng = eng
        ctx.save_for_backward(x.detach(), alpha_bar.detach())
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        x, alpha_bar = ctx.saved_tensors
        b = x.shape[0]
        ab = alpha_bar.expand(b)
        temp_rows = ((1 - ab) / ab).clamp_min(1e-30)
        g_q, g_t = ctx.eng.posterior_mean_backward(x, temp_rows, ab.rsqrt(), grad_out.float())
        g_q = g_q.to(x.device).view_as(x)
        g_t = g_t.to(x.device)
        lead = (-1,) + (1,) * (x.ndim - 1)
        grad_x = g_q * ab.rsqrt().view(lead) if ctx.needs_input_grad[0] else None
        grad_ab = None
        if ctx.needs_input_grad[1]:
            per_row = -0.5 * (g_q * x).reshape(b, -1).sum(1) * ab.pow(-1.5) - g_t / (ab * ab)
            grad_ab = per_row.sum().reshape(alpha_bar.shape) if alpha_bar.numel() == 1 else per_row.reshape(alpha_bar.shape)
        return grad_x, grad_ab, None
