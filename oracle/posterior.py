"""CPU restatement of the reference's empirical-denoiser hot path (torch, CPU).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity status: PINNED
against golden vectors generated from the unmodified reference
(``oracle/make_golden.py``, ``tests/test_oracle_golden.py``).

The reference's arithmetic for this path lives in third-party torch
(ATen -> MKL on CPU; pinned torch==2.5.1 in its ``poetry.lock:4753``, torch
2.11.0 in this image), so the restatement issues the same torch CPU ops in the
same order; every function cites the reference lines it follows.  Passing
``dtype=torch.float64`` re-evaluates the same formulas in double precision --
that is the arbiter used by the parity tests where the reference's own fp32
cancellation noise exceeds 1e-4.

All functions take the *noised* queries ``xt`` explicitly (no hidden RNG), so
the CUDA path and the oracle see identical inputs.  ``draw_noised_queries``
reproduces the reference's RNG call order for end-to-end comparisons.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import Tensor


# --------------------------------------------------------------------------
# utils/distance.py
# --------------------------------------------------------------------------

def row_norm_sqr(x: Tensor) -> Tensor:
    """||x_i||^2 per row via a batched (1,d)@(d,1) product. utils/distance.py:9-10."""
    return torch.bmm(x[:, None, :], x[:, :, None]).reshape(x.shape[0])


def gram(x: Tensor, y: Tensor) -> Tensor:
    """x @ y^T. utils/distance.py:5-6."""
    return torch.matmul(x, y.transpose(0, 1))


def pairwise_sqdist(x: Tensor, y: Optional[Tensor] = None) -> Tensor:
    """Squared distances by the norm expansion, in the reference's op order
    ``xn[:,None] - 2*G + yn`` and WITHOUT clamping (small negatives survive).
    utils/distance.py:13-21."""
    fx = x.reshape(x.shape[0], -1)
    fy = fx if y is None else y.reshape(y.shape[0], -1)
    xn = row_norm_sqr(fx)
    yn = row_norm_sqr(fy)
    g = gram(fx, fy)
    return xn[:, None] - 2 * g + yn


# --------------------------------------------------------------------------
# Boltzmann row statistics shared by utils/stats.py and scheduler.py
# --------------------------------------------------------------------------

def boltzmann_rows(energy: Tensor, t: Tensor | float, aux: Optional[Tensor] = None) -> dict[str, Tensor]:
    """Per-row statistics of p_j ~ exp(-(E_j - E_min)/T).

    Returns the min-shifted quantities the reference actually uses:
      e_min, argmin, log_l = log sum_j exp(-e_j), mean_e = <e>, mean_e2 = <e^2>,
      var_e = max(<e^2> - <e>^2, 0), aux_mean = <aux_j>, with e_j = (E_j - E_min)/T.
    Follows utils/stats.py:80-90 (logsumexp form) which is algebraically the
    same as utils/stats.py:282-288 (exp/sum/log form, see entropy_rows).
    """
    emin, amin = energy.min(dim=-1, keepdim=True)
    e = (energy - emin) / t
    log_l = torch.logsumexp(-e, dim=-1, keepdim=True)
    w = (-e - log_l).exp()
    mean_e = (w * e).sum(-1)
    mean_e2 = (w * e.pow(2)).sum(-1)
    out = {
        "e_min": emin.squeeze(-1),
        "argmin": amin.squeeze(-1),
        "log_l": log_l.squeeze(-1),
        "mean_e": mean_e,
        "mean_e2": mean_e2,
        "var_e": torch.clamp(mean_e2 - mean_e.pow(2), min=0),
        "weights": w,
    }
    if aux is not None:
        out["aux_mean"] = torch.matmul(w, aux.to(w.dtype))
    return out


def entropy_rows(energy: Tensor, t: Tensor | float, num_objects: int) -> Tensor:
    """Entropy per query exactly as utils/stats.py:282-289 computes it
    (subtract row min, exp, sum, log, renormalise, <E'>, S = logZ' + <E'>/T - log N)."""
    e = energy - energy.min(-1, keepdim=True).values
    z = -e / t
    log_part = z.exp().sum(-1).log()
    p = (z - log_part.unsqueeze(-1)).exp()
    avg_e = (p * e).sum(-1)
    return log_part + avg_e / t - math.log(num_objects)


# --------------------------------------------------------------------------
# utils/stats.py
# --------------------------------------------------------------------------

def _dataloader_seed_draw() -> None:
    """Creating a DataLoader iterator draws one int64 ``base_seed`` from the global
    CPU generator (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__).
    On CPU that interleaves with the reference's ``torch.randn`` calls; on CUDA the
    noise comes from the device generator and is unaffected."""
    torch.empty((), dtype=torch.int64).random_()


def draw_noised_queries(x0: Tensor, temp: Tensor, *, loader_iters: str = "none",
                        dtype=torch.float32) -> Tensor:
    """xt[i] = randn(B,...) * sqrt(T_i) + x0 with one ``torch.randn`` call per
    temperature, in schedule order -- the reference's RNG stream
    (utils/stats.py:74 and :273).  ``loader_iters`` replays the CPU-generator
    draws of the reference's DataLoader passes so a CPU run reproduces its
    numbers exactly: "per_temp" = one pass after each draw (compute_stats_batch,
    utils/stats.py:276), "once_before" = one pass up front
    (compute_metric_stats_batch, utils/stats.py:33).  Returns (n_T, B, ...)."""
    out = []
    if loader_iters == "once_before":
        _dataloader_seed_draw()
    for i in range(len(temp)):
        t = temp[i]
        out.append(torch.randn(*x0.shape) * t.sqrt() + x0)
        if loader_iters == "per_temp":
            _dataloader_seed_draw()
    return torch.stack(out).to(dtype)


def draw_noise(shape, n_temps: int, *, loader_iters: str = "none") -> Tensor:
    """The raw standard-normal draws behind ``draw_noised_queries`` (same RNG call order), (n_T, *shape)."""
    out = []
    if loader_iters == "once_before":
        _dataloader_seed_draw()
    for _ in range(n_temps):
        out.append(torch.randn(*shape))
        if loader_iters == "per_temp":
            _dataloader_seed_draw()
    return torch.stack(out)


def entropy_batch(xt: Tensor, data: Tensor, temp: Tensor, *, chunk: Optional[int] = None,
                  dtype=torch.float32) -> Tensor:
    """``compute_stats_batch`` for given noised queries xt (n_T,B,...).
    utils/stats.py:261-292; ``chunk`` reproduces the DataLoader chunking of the
    dataset (:276-280), which only affects which rows share one GEMM call."""
    data = data.to(dtype)
    n = data.shape[0]
    rows = []
    for i in range(len(temp)):
        q = xt[i].to(dtype)
        if chunk is None:
            energy = 0.5 * pairwise_sqdist(q, data)
        else:
            energy = torch.zeros(q.shape[0], n, dtype=dtype)
            for s in range(0, n, chunk):
                energy[:, s:s + chunk] = 0.5 * pairwise_sqdist(q, data[s:s + chunk])
        rows.append(entropy_rows(energy, temp[i].to(dtype), n))
    return torch.stack(rows)


def gaussian_cluster_metric(sigma_sq: Tensor, t: Tensor) -> Tensor:
    """G_reg = 1/2 s (s + 2T)/(s + T)^2.  utils/stats.py:102 and :107."""
    return 0.5 * sigma_sq * (sigma_sq + 2 * t) / (sigma_sq + t).pow(2)


def metric_batch(xt: Tensor, data: Tensor, temp: Tensor, *, regularize: bool = False,
                 sigma_reg_sq_per_point: Optional[Tensor] = None, dtype=torch.float32,
                 return_rows: bool = False):
    """``compute_metric_stats_batch`` for given xt (n_T,B,...).
    utils/stats.py:71-111: Var_w(E/T), optional regularisation floor
    (per-point k-NN sigma :98-103, global 1e-3 fallback :104-108), batch mean."""
    data = data.to(dtype)
    vals, rows = [], []
    for i in range(len(temp)):
        t = temp[i].to(dtype)
        energy = 0.5 * pairwise_sqdist(xt[i].to(dtype), data)
        st = boltzmann_rows(energy, t, aux=sigma_reg_sq_per_point)
        var = st["var_e"]
        if regularize:
            if sigma_reg_sq_per_point is not None:
                var = torch.maximum(var, gaussian_cluster_metric(st["aux_mean"], t))
            else:
                var = torch.maximum(var, gaussian_cluster_metric(torch.tensor(1e-3, dtype=dtype), t))
        rows.append(var)
        vals.append(var.mean().to(torch.float32))
    if return_rows:
        return torch.stack(vals), torch.stack(rows)
    return torch.stack(vals)


def dataset_trace_sigma0(data: Tensor) -> float:
    """Tr Sigma_0 = sum of per-dimension unbiased variances. utils/stats.py:39,176."""
    flat = data.reshape(len(data), -1)
    return torch.var(flat, dim=0).sum().item()


def knn_sigma_reg_sq(data: Tensor, knn_k: int, sigma_reg_scale: float) -> Tensor:
    """Per-point regulariser d_k^2 * scale / D with d_k the distance to the k-th
    neighbour (self excluded).  utils/stats.py:137-146 (sklearn kneighbors with
    n_neighbors=k+1 on the dataset itself, Euclidean).  Brute force in float64
    like sklearn's exact path, cast to float32 like :144."""
    flat = data.reshape(len(data), -1).double()
    d2 = torch.cdist(flat, flat).pow(2)
    kth = torch.topk(d2, knn_k + 1, dim=1, largest=False).values[:, -1]
    return (kth.float() * sigma_reg_scale / float(flat.shape[1]))


# --------------------------------------------------------------------------
# diffusion/scheduler/scheduler.py  (ideal denoiser)
# --------------------------------------------------------------------------

def linear_beta_log_temp(tau: Tensor, min_temp: float, max_temp: float) -> Tensor:
    """LinearBetaScheduler.log_temp_from_tau. diffusion/scheduler/linear.py:5-13."""
    scale = 1 + min_temp
    gamma = math.log((1 + max_temp) / scale)
    return ((tau.pow(2) * gamma).exp() * scale - 1).log()


def posterior_mean_x0(xt: Tensor, alpha_bar: Tensor, data: Tensor, *, dtype=torch.float32) -> Tensor:
    """Ideal denoiser x0_hat = sum_j p_j y_j for VP noise with scalar-like alpha_bar.
    diffusion/scheduler/scheduler.py:60-69: h = 1/2 ||xt - sqrt(ab) y||^2 (the whole
    dataset is scaled, :64), shift by the row min, p = exp(-h/(1-ab)), normalise, p @ data."""
    x = xt.to(dtype)
    y = data.to(dtype)
    ab = alpha_bar.to(dtype).reshape(-1, *[1] * (x.ndim - 1))
    h = 0.5 * pairwise_sqdist(x, ab.sqrt() * y.reshape(len(y), *x.shape[1:]))
    h = h - h.min(1, keepdim=True).values
    p = (-h / (1 - ab).reshape(-1, 1)).exp()
    p = p / p.sum(1, keepdim=True)
    return torch.matmul(p, y.reshape(len(y), -1)).reshape(x.shape)


# --------------------------------------------------------------------------
# utils/metric_utils.py  (Monte-Carlo Fisher metric; y-samples passed in)
# --------------------------------------------------------------------------

def draw_mc_samples(x_samples: Tensor, n_y: int):
    """RNG order of utils/metric_utils.py:24-26 / 83-87 / 172-174: randint then randn."""
    k, d = x_samples.shape
    idx = torch.randint(0, k, (n_y,))
    eps = torch.randn((n_y, d))
    return idx, eps


def metric_scalar_from_samples(y: Tensor, x: Tensor, sigma_sq: Tensor) -> Tensor:
    """utils/metric_utils.py:40-58 for given y samples: softmax over prior samples,
    posterior mean of the per-point lambda-score, D/2 - Var_y."""
    d = x.shape[1]
    y_sq = torch.sum(y ** 2, dim=1, keepdim=True)
    x_sq = torch.sum(x ** 2, dim=1).unsqueeze(0)
    sq = y_sq + x_sq - 2 * torch.mm(y, x.t())
    w = torch.softmax(-0.5 * sq / sigma_sq, dim=1)
    marg = torch.sum(w * (-0.5 * d + 0.5 * sq / sigma_sq), dim=1)
    return 0.5 * d - torch.var(marg)


def diag_marginal_scores(y: Tensor, x: Tensor, sigma_diag: Tensor):
    """Softmax weights under a diagonal covariance and the weighted per-dimension
    second moment  m2[b,i] = sum_k w[b,k] (y[b,i]-x[k,i])^2.
    utils/metric_utils.py:131-137 and :187-192 (materialises (n_y,K,D) there)."""
    diff2 = (y.unsqueeze(1) - x.unsqueeze(0)) ** 2
    w = torch.softmax(-0.5 * torch.sum(diff2 / sigma_diag, dim=2), dim=1)
    return w, torch.sum(w.unsqueeze(2) * diff2, dim=1)


def metric_matrix_from_samples(y: Tensor, x: Tensor, sigma_diag: Tensor) -> Tensor:
    """utils/metric_utils.py:128-151: scores -1/2 + 1/2 diff^2/S_ii, 1/2 - Var_y."""
    diff2 = (y.unsqueeze(1) - x.unsqueeze(0)) ** 2
    w = torch.softmax(-0.5 * torch.sum(diff2 / sigma_diag, dim=2), dim=1)
    marg = torch.sum(w.unsqueeze(2) * (-0.5 + 0.5 * diff2 / sigma_diag), dim=1)
    return 0.5 * torch.ones(x.shape[1]) - torch.var(marg, dim=0)


def rescaled_metric_from_samples(y: Tensor, x: Tensor, sigma_diag: Tensor) -> Tensor:
    """utils/metric_utils.py:186-216: scores -1/2/S + 1/2 diff^2/S^2,
    G = 1/2/S^2 - Var_y, times 4 S^2/(Var(x) + 2 S)."""
    diff2 = (y.unsqueeze(1) - x.unsqueeze(0)) ** 2
    w = torch.softmax(-0.5 * torch.sum(diff2 / sigma_diag, dim=2), dim=1)
    marg = torch.sum(w.unsqueeze(2) * (-0.5 / sigma_diag + 0.5 * diff2 / sigma_diag ** 2), dim=1)
    g = 0.5 / sigma_diag ** 2 - torch.var(marg, dim=0)
    return g * (4 * sigma_diag ** 2 / (torch.var(x, dim=0) + 2 * sigma_diag))


# --------------------------------------------------------------------------
# Independent slow cross-check (pure float64 loops, small cases only)
# --------------------------------------------------------------------------

def boltzmann_rows_loops(xt, data, t):
    """O(B*N*d) python/numpy float64 evaluation of E_min, log l, <e>, <e^2> and the
    posterior mean, using the direct ||x-y||^2 form (no norm expansion)."""
    import numpy as np
    x = np.asarray(xt, dtype=np.float64).reshape(len(xt), -1)
    y = np.asarray(data, dtype=np.float64).reshape(len(data), -1)
    out = {k: np.zeros(len(x)) for k in ("e_min", "log_l", "mean_e", "mean_e2")}
    out["argmin"] = np.zeros(len(x), dtype=np.int64)
    out["mean_y"] = np.zeros_like(x)
    for b in range(len(x)):
        en = 0.5 * ((x[b][None, :] - y) ** 2).sum(1)
        j = int(en.argmin())
        e = (en - en[j]) / float(t)
        w = np.exp(-e)
        l = w.sum()
        out["e_min"][b] = en[j]
        out["argmin"][b] = j
        out["log_l"][b] = np.log(l)
        out["mean_e"][b] = (w * e).sum() / l
        out["mean_e2"][b] = (w * e * e).sum() / l
        out["mean_y"][b] = (w[:, None] * y).sum(0) / l
    return out
