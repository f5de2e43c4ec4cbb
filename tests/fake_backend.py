"""CPU test double of ``pdm_b200.backend.CudaBackend``.

TEST INFRASTRUCTURE ONLY: it lets the ``-m "not gpu"`` suite exercise the host logic (blocking of the
temperature schedule, RNG order, dataset caching, partial-record merge, gloo sharding) without a GPU.  It
speaks the same tensor-level API and the same 8-float partial-record format as the CUDA library, computing
with torch on the CPU.  The product never imports it.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

PART = 8


def _pack_idx(idx: Tensor) -> Tensor:
    return idx.to(torch.int64).contiguous().view(torch.int32).view(-1, 2).view(torch.float32)


def _unpack_idx(rec: Tensor) -> Tensor:
    return rec[..., 5:7].contiguous().view(torch.int32).view(torch.int64).reshape(rec.shape[:-1])


class FakeBackend:
    name = "fake-cpu"

    def __init__(self, tile: int = 16):
        self.device = torch.device("cpu")
        self.tile = tile
        self.launches = 0
        self.calls: list[str] = []
        self.kernel_events = None
        self.last_plan = (1, 1, 1)

    def supports_tensor_path(self) -> bool:
        return False

    # ---- preparation ---------------------------------------------------------------------------
    def row_norms(self, x: Tensor) -> Tensor:
        self.calls.append("row_norms")
        return (x.double() ** 2).sum(1).float()

    def absmax(self, x: Tensor) -> Tensor:
        return x.abs().max().reshape(1)

    def prepare_rows(self, src, rows, *, noise=None, sigma=None, post=None, fixed_scale=0.0, want_x=False,
                     want_norms=True, want_split=True) -> dict:
        self.calls.append("prepare_rows")
        if noise is not None:
            reps = (rows + src.shape[0] - 1) // src.shape[0]
            v = noise * sigma[:, None] + src.repeat(reps, 1)[:rows]
        else:
            v = src[:rows].clone()
        if post is not None:
            v = v * post[:, None]
        return {"x": v, "norms": (v.double() ** 2).sum(1).float(), "hi": None, "lo": None, "inv_scale": None}

    def column_moments(self, y: Tensor):
        yd = y.double()
        return yd.sum(0), (yd ** 2).sum(0), torch.stack([y.min(), y.max()])

    def transpose_split(self, y, scale):
        raise AssertionError("the fake backend has no tensor path")

    # ---- fused pass ----------------------------------------------------------------------------
    def posterior_stats(self, *, precision, M, N, d, q_norm, y_norm, inv_temp, q=None, y=None, y_aux=None,
                        index_offset=0, n_splits=0, want_partials=True, energy_out=None, energy_mult=1.0, **_kw):
        assert precision == "exact"
        self.calls.append("posterior_stats")
        self.launches += 1
        g = torch.matmul(q, y.t())
        energy = 0.5 * ((q_norm[:, None] - 2 * g) + y_norm[None, :])
        if energy_out is not None:
            energy_out.copy_(energy_mult * energy)
        if not want_partials:
            return None
        n_tiles = (N + self.tile - 1) // self.tile
        s = n_splits if n_splits > 0 else min(3, n_tiles)
        parts = torch.zeros(M, s, PART)
        cols = torch.arange(N)
        for k in range(s):
            sel = ((cols // self.tile) % s) == k
            if not sel.any():
                parts[:, k, 0] = math.inf
                parts[:, k, 5:7] = _pack_idx(torch.full((M,), -1))
                continue
            e = energy[:, sel]
            m, am = e.min(dim=1)
            ee = (e - m[:, None]) * inv_temp[:, None]
            w = torch.exp(-ee)
            parts[:, k, 0] = m
            parts[:, k, 1] = w.sum(1)
            parts[:, k, 2] = (w * ee).sum(1)
            parts[:, k, 3] = (w * ee * ee).sum(1)
            if y_aux is not None:
                parts[:, k, 4] = w @ y_aux[sel]
            parts[:, k, 5:7] = _pack_idx(cols[sel][am] + index_offset)
        return parts

    @staticmethod
    def _combine(parts: Tensor, inv_temp: Tensor):
        """parts (M, R, 8) -> merged (m, l, a1, a2, aux, idx) with the shift rule of SURVEY.md section 5."""
        p = parts.double()
        m_all, l_all = p[..., 0], p[..., 1]
        valid = l_all > 0
        m_eff = torch.where(valid, m_all, torch.full_like(m_all, math.inf))
        m = m_eff.min(dim=1).values
        delta = torch.where(valid, (m_all - m[:, None]) * inv_temp.double()[:, None], torch.zeros_like(m_all))
        c = torch.where(valid, torch.exp(-delta), torch.zeros_like(delta))
        l = (c * l_all).sum(1)
        a1 = (c * (p[..., 2] + delta * l_all)).sum(1)
        a2 = (c * (p[..., 3] + 2 * delta * p[..., 2] + delta * delta * l_all)).sum(1)
        aux = (c * p[..., 4]).sum(1)
        idx_all = _unpack_idx(parts)
        big = torch.iinfo(torch.int64).max
        cand = torch.where(valid & (m_all == m[:, None].float()), idx_all, torch.full_like(idx_all, big))
        return m, l, a1, a2, aux, cand.min(dim=1).values

    def reduce(self, parts: Tensor, inv_temp: Tensor) -> Tensor:
        self.calls.append("reduce")
        m, l, a1, a2, aux, idx = self._combine(parts, inv_temp)
        out = torch.zeros(parts.shape[0], 1, PART)
        for j, v in enumerate((m, l, a1, a2, aux)):
            out[:, 0, j] = v.float()
        out[:, 0, 5:7] = _pack_idx(idx)
        return out

    def merge(self, parts: Tensor, inv_temp: Tensor, n_total: int):
        self.calls.append("merge")
        if parts.dim() == 4:
            parts = parts.permute(1, 0, 2, 3).reshape(parts.shape[1], -1, PART)
        m, l, a1, a2, aux, idx = self._combine(parts, inv_temp)
        mean_e, mean_e2 = a1 / l, a2 / l
        out = torch.stack([m, l.log(), mean_e, mean_e2, (mean_e2 - mean_e ** 2).clamp(min=0), aux / l,
                           l.log() + mean_e - math.log(n_total), l]).float()
        return out, idx

    # ---- posterior mean ------------------------------------------------------------------------
    def weights_from_energy(self, energy, e_min, l, inv_temp, *, split):
        assert not split
        return torch.exp(-(energy - e_min[:, None]) * inv_temp[:, None]) / l[:, None]

    def weighted_mean_exact(self, p, y, out=None, accumulate=False):
        r = p @ y
        if out is None:
            return r
        if accumulate:
            out += r
        else:
            out.copy_(r)
        return out

    def denoiser_backward_weights(self, energy, sdot, e_min, l, inv_temp, s_scale=None, a_in=None):
        e = (energy - e_min[:, None]) * inv_temp[:, None]
        p = torch.exp(-e) / l[:, None]
        s = sdot if s_scale is None else sdot * s_scale[:, None]
        a = (p * s).sum(1) if a_in is None else a_in
        w = p * (s - a[:, None])
        return w, torch.stack([a, (w * e).sum(1)], dim=1)

    def sampler_step(self, x0_hat, xt, noise, c_x0, c_xt, c_noise, out=None):
        r = c_x0 * x0_hat + c_xt * xt
        if noise is not None:
            r = r + c_noise * noise
        if out is None:
            return r
        out.copy_(r)
        return out
