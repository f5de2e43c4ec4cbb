"""Drop-in for the reference's ``utils/metric_utils.py`` (Monte-Carlo Fisher-metric estimators) on the
B200 engine.  Signatures, RNG call order (``randint`` then ``randn`` on ``x_samples.device``) and return
values follow utils/metric_utils.py:4-58, 60-151, 153-216.

All three estimators are a softmax over the K prior samples for each of n_y noised samples -- the same
query-vs-dataset pass as the rest of the hot path -- followed by the variance over y of a posterior
expectation:
  * scalar case: the per-point lambda-score is -D/2 + E_k/T with E_k = |y - x_k|^2/2 and T = sigma^2, so
    its posterior mean is -D/2 + <e> + E_min/T, straight from the fused statistics pass;
  * diagonal case: whiten by 1/sqrt(Sigma_ii) (T = 1); the per-dimension score needs
    sum_k w_k (y_i - x_ki)^2 = y_i^2 - 2 y_i <x_i> + <x_i^2>, i.e. the posterior mean of [x, x^2] (about the
    samples' mean, x^2 as an fp32 pair) -- one pass of the posterior-mean contraction instead of the reference's
    (n_y, K, D) tensor (:131, :187).
"""
from __future__ import annotations

import torch
from torch import Tensor

from pdm_b200 import EmpiricalDataset, PosteriorEngine
from pdm_b200.engine import default_backend


def _marginal_scalar_scores(y_samples: Tensor, x_samples: Tensor, sigma_sq: Tensor) -> Tensor:
    """E_{x|y}[ -D/2 + |y-x|^2/(2 sigma^2) ] for every y (utils/metric_utils.py:40-51)."""
    d = x_samples.shape[1]
    eng = PosteriorEngine(EmpiricalDataset(x_samples, backend=default_backend()))
    t = sigma_sq.to(torch.float32).reshape(1).expand(y_samples.shape[0])
    st = eng.stats(y_samples, t)
    return (-0.5 * d + st["mean_e"] + st["e_min"] / t.to(st["e_min"].device)).to(x_samples.device)


def _diag_second_moments(y_samples: Tensor, x_samples: Tensor, sigma_diag: Tensor) -> Tensor:
    """m2[b, i] = sum_k w[b, k] (y[b, i] - x[k, i])^2 with w the softmax of -1/2 sum_i (.)^2/Sigma_ii.

    Expanded as y^2 - 2 y <x> + <x^2> the three terms cancel to the distance of y from its nearest samples -- at small
    Sigma many orders below |x|^2 -- so the expansion is taken about the samples' mean, <x^2> is carried as an fp32 pair
    (hi + lo = x^2 to 48 bits: a posterior that is a delta on one sample then reproduces (y - x)^2 exactly, the engine
    gathers such rows) and the three terms are combined in float64."""
    s = torch.sqrt(sigma_diag)
    xw, yw = x_samples / s, y_samples / s                       # whitened coordinates, T = 1
    eng = PosteriorEngine(EmpiricalDataset(xw, backend=default_backend()))
    t = torch.ones(yw.shape[0])
    mu = x_samples.mean(dim=0, keepdim=True)
    xc, yc = (x_samples - mu).double(), (y_samples - mu).double()
    sq = xc * xc
    sq_hi = sq.float()
    sq_lo = (sq - sq_hi.double()).float()
    values = torch.cat((xc.float(), sq_hi, sq_lo), dim=1)
    mean = eng.posterior_mean(yw, t, values=values).to(x_samples.device).double()
    d = x_samples.shape[1]
    ex, ex2 = mean[:, :d], mean[:, d:2 * d] + mean[:, 2 * d:]
    return (yc * yc - 2 * yc * ex + ex2).clamp_min(0).to(x_samples.dtype)


def compute_metric_scalar(log_sigma_sq, x_samples, n_y_samples=10000):
    """G(lambda) for Sigma = sigma^2 I, lambda = log sigma^2 (utils/metric_utils.py:4-58)."""
    device = x_samples.device
    K, D = x_samples.shape
    sigma_sq = torch.exp(torch.tensor(log_sigma_sq, device=device))
    sigma = torch.sqrt(sigma_sq)
    indices = torch.randint(0, K, (n_y_samples,), device=device)
    eps = torch.randn((n_y_samples, D), device=device)
    y_samples = x_samples[indices] + sigma * eps
    marginal_scores = _marginal_scalar_scores(y_samples, x_samples, sigma_sq)
    return 0.5 * D - torch.var(marginal_scores)


def compute_metric_matrix(Lambda, x_samples, n_y_samples=10000):
    """Diagonal of the metric for Sigma = exp(Lambda) (utils/metric_utils.py:60-151)."""
    device = x_samples.device
    K, D = x_samples.shape
    evals, evecs = torch.linalg.eigh(Lambda)
    Sigma = evecs @ torch.diag(torch.exp(evals)) @ evecs.t()
    indices = torch.randint(0, K, (n_y_samples,), device=device)
    eps = torch.randn((n_y_samples, D), device=device)
    sqrt_Sigma = evecs @ torch.diag(torch.sqrt(torch.exp(evals))) @ evecs.t()
    y_samples = x_samples[indices] + (sqrt_Sigma @ eps.t()).t()
    Sigma_diag = torch.diag(Sigma)
    m2 = _diag_second_moments(y_samples, x_samples, Sigma_diag)
    marginal_scores = -0.5 + 0.5 * m2 / Sigma_diag
    return 0.5 * torch.ones(D, device=device) - torch.var(marginal_scores, dim=0)


def compute_rescaled_metric_matrix(Sigma, x_samples, n_y_samples=10000):
    """Rescaled metric for the parameterisation theta = Sigma (utils/metric_utils.py:153-216)."""
    device = x_samples.device
    K, D = x_samples.shape
    Sigma_diag = Sigma if Sigma.ndim == 1 else torch.diag(Sigma)
    indices = torch.randint(0, K, (n_y_samples,), device=device)
    eps = torch.randn((n_y_samples, D), device=device)
    if Sigma.ndim == 1:
        y_samples = x_samples[indices] + torch.sqrt(Sigma) * eps
    else:
        evals, evecs = torch.linalg.eigh(Sigma)
        sqrt_Sigma = evecs @ torch.diag(torch.sqrt(torch.clamp(evals, min=1e-10))) @ evecs.t()
        y_samples = x_samples[indices] + (sqrt_Sigma @ eps.t()).t()
    m2 = _diag_second_moments(y_samples, x_samples, Sigma_diag)
    marginal_scores = -0.5 / Sigma_diag + 0.5 * m2 / Sigma_diag ** 2
    G_ii = 0.5 / Sigma_diag ** 2 - torch.var(marginal_scores, dim=0)
    factor = 4 * Sigma_diag ** 2 / (torch.var(x_samples, dim=0) + 2 * Sigma_diag)
    return G_ii * factor
