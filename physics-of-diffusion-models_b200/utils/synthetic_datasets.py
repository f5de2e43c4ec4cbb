"""Synthetic point sets used as datasets / benchmark inputs (same generators and signatures as the
reference's utils/synthetic_datasets.py:6-35; pure CPU tensor construction, no hot-path arithmetic)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor


def generate_simplex(d: int) -> Tensor:
    """d+1 vertices of a regular simplex: the d basis vectors plus one point on the diagonal."""
    last = torch.full((1, d), (1 - (1 + d) ** 0.5) / d)
    return torch.cat((torch.eye(d), last), dim=0)


def generate_cross_polytope(d: int) -> Tensor:
    """The 2d vertices +-e_i."""
    eye = torch.eye(d)
    return torch.cat((eye, -eye), dim=0)


def sample_on_hypersphere(d: int, n: Optional[int] = None) -> Tensor:
    """n (default 10 d) points uniform on the sphere of radius sqrt(d)."""
    pts = torch.randn(n or 10 * d, d)
    pts /= torch.norm(pts, dim=1, keepdim=True) / d ** 0.5
    return pts


def generate_gaussian(d: int, n: int = 1000) -> Tensor:
    return torch.randn(n, d)


def generate_dataset(name: str = "hypersphere", d: int = 100) -> Tensor:
    makers = {"simplex": generate_simplex, "cross-polytope": generate_cross_polytope,
              "hypersphere": sample_on_hypersphere, "gaussian": generate_gaussian}
    if name not in makers:
        raise ValueError(f"Invalid name: {name}")
    return makers[name](d)
