"""Seeded synthetic inputs for the oracle, the golden fixtures and the CPU baseline.

Test infrastructure only (see ``oracle/__init__.py``).  Generators are CPU torch
with an explicit ``torch.Generator`` so they do not disturb the global RNG
stream the reference's functions consume.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor


def gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def anisotropic_gmm(dim: int, n_components: int, n_samples: int, seed: int) -> Tensor:
    """Mixture of ``n_components`` Gaussians with random means N(0,I) and rotated
    covariances with spectrum 0.01*exp(-linspace(0,5,dim)) -- the recipe of the
    reference's scripts/reproduce_high_dim.py:18-46 (restated with a private
    generator; not bit-identical to that script's global-RNG draw)."""
    g = gen(seed)
    means = torch.randn(n_components, dim, generator=g)
    spectrum = torch.exp(-torch.linspace(0, 5, dim)) * 0.01
    factors = []
    for _ in range(n_components):
        q, _ = torch.linalg.qr(torch.randn(dim, dim, generator=g))
        factors.append(q * spectrum.sqrt())          # Sigma = F F^T
    comp = torch.multinomial(torch.full((n_components,), 1.0 / n_components), n_samples,
                             replacement=True, generator=g)
    z = torch.randn(n_samples, dim, generator=g)
    out = torch.empty(n_samples, dim)
    for c in range(n_components):
        m = comp == c
        out[m] = means[c] + z[m] @ factors[c].T
    return out


def uniform_images(n: int, shape: tuple[int, ...], seed: int) -> Tensor:
    """U[-1,1] 'images' (the normalised pixel range of utils/data.py:52-60)."""
    return torch.rand(n, *shape, generator=gen(seed)) * 2 - 1


def clustered_images(n: int, shape: tuple[int, ...], n_clusters: int, spread: float, seed: int) -> Tensor:
    """Points in tight clusters so that the posterior's transition region sits at
    low temperature (exercises the fp32 cancellation regime)."""
    g = gen(seed)
    centres = torch.rand(n_clusters, *shape, generator=g) * 2 - 1
    which = torch.randint(0, n_clusters, (n,), generator=g)
    return (centres[which] + spread * torch.randn(n, *shape, generator=g)).clamp_(-1, 1)


def hypersphere(d: int, n: int, seed: int) -> Tensor:
    """Uniform on the sphere of radius sqrt(d). utils/synthetic_datasets.py:14-17."""
    s = torch.randn(n, d, generator=gen(seed))
    return s / (torch.norm(s, dim=1, keepdim=True) / math.sqrt(d))


def ddpm_temperatures(n_steps: int, min_temp: float = 1e-4, max_temp: float = 2.478e4) -> Tensor:
    """T_k = exp(log_temp(tau_k)), tau = linspace(0,1,n+1)[1:], linear-beta schedule
    (diffusion/scheduler/linear.py:5-13)."""
    tau = torch.linspace(0, 1, n_steps + 1, dtype=torch.float64)[1:]
    scale = 1 + min_temp
    gamma = math.log((1 + max_temp) / scale)
    return ((tau.pow(2) * gamma).exp() * scale - 1).float()
