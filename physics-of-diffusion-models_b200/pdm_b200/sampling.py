"""Reverse-diffusion sampling with the closed-form (ideal) denoiser, entirely on the engine.

The reference samples with ``DDPMSampler`` (diffusion/ddpm_sampling.py:14-139) around ``DDPMTrue``: per step
``get_predictions`` (diffusion/ddpm/ddpm.py:33-36), the x0 / eps algebra of ``DDPMPredictions`` (:17-20) and the DDPM
or DDIM update (:94-110) -- a dozen elementwise torch kernels and one host sync (``prev_log_temp > -inf``) per step.
That loop keeps working unchanged on top of the drop-in ``DDPMTrue``; this module is the fused form of the same
recurrence: the schedule is read once on the host, every step is
``PosteriorEngine.posterior_mean`` (VP query x_t / sqrt(ab) at temperature (1 - ab)/ab) followed by ONE kernel
``pdm_sampler_step_f32`` whose three coefficients carry the whole update, and nothing of the loop depends on a device
value.  The RNG stream is the reference's: ``torch.randn`` for the initial state and, for DDPM steps, one
``torch.randn_like`` per step except the last.
"""
from __future__ import annotations

import math
import os
from typing import Iterable, Optional

import torch
from torch import Tensor

from .engine import EmpiricalDataset, PosteriorEngine, EngineConfig, default_backend


def step_coefficients(alpha_bar: float, prev_alpha_bar: float, step_type: str) -> tuple[float, float, float]:
    """(c_x0, c_xt, c_noise) of  x_prev = c_x0 x0_hat + c_xt x_t + c_noise eps'  for one step ab -> ab'.

    DDPM (ddpm_sampling.py:99-107): alpha = ab/ab', beta = 1 - alpha,
        c_x0 = sqrt(ab') beta / (1 - ab),  c_xt = sqrt(alpha)(1 - ab')/(1 - ab),  c_noise = sqrt((1 - ab')/(1 - ab) beta).
    DDIM (:109-110) with eps = (x_t - sqrt(ab) x0_hat)/sqrt(1 - ab) (ddpm.py:19):
        c_x0 = sqrt(ab') - sqrt(1 - ab') sqrt(ab)/sqrt(1 - ab),  c_xt = sqrt(1 - ab')/sqrt(1 - ab),  c_noise = 0."""
    ab, abp = float(alpha_bar), float(prev_alpha_bar)
    if step_type == "ddpm":
        alpha = ab / abp
        beta = 1.0 - alpha
        return (math.sqrt(abp) * beta / (1.0 - ab), math.sqrt(alpha) * (1.0 - abp) / (1.0 - ab),
                math.sqrt(max((1.0 - abp) / (1.0 - ab) * beta, 0.0)))
    if step_type == "ddim":
        r = math.sqrt(max(1.0 - abp, 0.0)) / math.sqrt(1.0 - ab)
        return (math.sqrt(abp) - r * math.sqrt(ab), r, 0.0)
    raise ValueError(f"unknown step type: {step_type}")


class IdealSampler:
    """Sampler for the empirical (ideal) denoiser of a training set resident on the GPU.

    ``log_temp``: ascending log-temperatures of the schedule (what ``DDPMSampler.log_temp`` holds); sampling walks it
    from the last entry down and finishes at the clean state (log_temp = -inf, alpha_bar = 1)."""

    def __init__(self, train_data: Tensor, log_temp: Tensor | Iterable[float], step_type: str = "ddim",
                 engine: Optional[PosteriorEngine] = None, config: Optional[EngineConfig] = None, query_group=None,
                 use_graphs: Optional[bool] = None):
        """``query_group``: torch.distributed group whose ranks share the work of every batch (mode 2:
        the dataset -- 0.6 GB at CIFAR-10 shape -- is replicated, each rank owns a contiguous slice of the trajectories,
        no collective inside the loop; the samples are all-gathered once at the end).  Every rank draws the whole batch's
        noise and keeps its rows, so the result is the one-GPU result for the same seed, whatever the number of ranks."""
        if step_type not in ("ddpm", "ddim"):
            raise ValueError(f"unknown step type: {step_type}")
        # CUDA graphs (PDM_SAMPLER_GRAPHS=0 turns them off): a step never reads anything back from the device, so it is
        # captured once per (batch slice, screening mode) and replayed with five device scalars rewritten in between --
        # the ~60 launches and allocations of a step cost the host one graph launch (the low-noise steps of a trajectory
        # split over 8 GPUs are otherwise bound by the host, not the GPU).
        if use_graphs is None:
            use_graphs = os.environ.get("PDM_SAMPLER_GRAPHS", "1") != "0"
        self.use_graphs = bool(use_graphs)
        self._graphs: dict = {}
        self._graph_seen: dict = {}
        self._graph_pool = None
        self.graph_replays = 0
        self.engine = engine if engine is not None else PosteriorEngine(
            EmpiricalDataset(train_data, backend=default_backend()), config)
        self.backend = self.engine.backend
        self.step_type = step_type
        # An engine over a row-sharded dataset (too large to replicate): the trajectories are split over the SAME group;
        # every step all-gathers the current states (each shard has to see every query) and gets its own slice of the
        # posterior means back from a reduce-scatter (mode 1).
        self.dataset_sharded = self.engine.world > 1
        if self.dataset_sharded:
            if query_group is not None:
                raise ValueError("an engine over a row-sharded dataset splits the trajectories over its own group")
            query_group = self.engine.group
        self.query_group = query_group
        self.q_world, self.q_rank = 1, 0
        if query_group is not None:
            import torch.distributed as dist
            self.q_world, self.q_rank = dist.get_world_size(query_group), dist.get_rank(query_group)
        lt = torch.as_tensor(log_temp, dtype=torch.float64).reshape(-1).cpu()
        self.log_temp = lt.tolist()                                   # read once: the loop never syncs on a device value
        self.alpha_bar = torch.sigmoid(-lt).tolist()                  # alpha_bar_from_log_temp, scheduler.py:24-25
        self.obj_size = tuple(train_data.shape[1:])

    def _my_rows(self, batch_size: int) -> tuple[int, int]:
        per = (batch_size + self.q_world - 1) // self.q_world
        return min(batch_size, self.q_rank * per), min(batch_size, (self.q_rank + 1) * per)

    # ---- one reverse-diffusion step ----------------------------------------------------------------------------------
    def _step_eager(self, flat: Tensor, batch_size: int, lo: int, hi: int, ab: float, abp: float, last: bool) -> None:
        dev, d, rows = self.backend.device, self.engine.ds.d, hi - lo
        t = (1.0 - ab) / ab
        if self.dataset_sharded:
            every = self._gather_rows(flat.view(rows, *self.obj_size), batch_size, copy=False).view(batch_size, d)
            x0_hat = self.engine.posterior_mean(every, torch.full((batch_size,), t, device=dev),
                                                post=torch.full((batch_size,), 1.0 / math.sqrt(ab), device=dev),
                                                temp_bounds=(t, t), scatter=True)[:rows]
        else:
            ones = torch.ones(rows, dtype=torch.float32, device=dev)
            x0_hat = self.engine.posterior_mean(flat, ones * t, post=ones * (1.0 / math.sqrt(ab)), temp_bounds=(t, t))
        c_x0, c_xt, c_noise = step_coefficients(ab, abp, self.step_type)
        noise = None
        if self.step_type == "ddpm" and not last:                 # no draw on the last step (ddpm_sampling.py:107)
            noise = torch.randn(batch_size, *self.obj_size, device=dev)       # = randn_like(xt) of the whole batch
            noise = noise.view(batch_size, d)[lo:hi].contiguous()
        self.backend.sampler_step(x0_hat, flat, noise, c_x0, c_xt, c_noise if noise is not None else 0.0, out=flat)

    def _graph_mode(self, t: float) -> tuple:
        """(screen?, E4M3 stage?) exactly as PosteriorEngine.posterior_mean would decide them from its marks."""
        eng = self.engine
        eng._pm_poll()
        if not eng.screening_usable():
            return (False, False)
        return (t < eng._pm_screen_t, t < eng._pm_f8_t)

    def _step_graphed(self, flat: Tensor, batch_size: int, lo: int, hi: int, ab: float, abp: float, last: bool,
                      scal_row: Tensor) -> bool:
        """Replay (or capture, the second time a configuration is met -- the first time runs eagerly and doubles as the
        warm-up of every lazy initialisation) the CUDA graph of one step.  False: this step has to run eagerly."""
        dev, d, rows = self.backend.device, self.engine.ds.d, hi - lo
        t = (1.0 - ab) / ab
        has_noise = self.step_type == "ddpm" and not last
        if last and self.step_type == "ddpm":
            return False                                  # the one step without a noise draw (ddpm_sampling.py:107): eager, once
        mode = self._graph_mode(t)
        key = (flat.data_ptr(), rows, batch_size, lo, mode, has_noise)
        seen = self._graph_seen.get(key, 0)
        self._graph_seen[key] = seen + 1
        if seen == 0:
            return False
        entry = self._graphs.get(key)
        if entry is None:
            scal = torch.empty(5, dtype=torch.float32, device=dev)
            counts = torch.zeros(2, dtype=torch.int32).pin_memory()
            scal.copy_(scal_row)
            eng = self.engine

            def body():
                temp_rows = scal[0:1].expand(rows)
                post = scal[1:2].expand(rows)
                x0_hat = eng.posterior_mean(flat, temp_rows, post=post, temp_bounds=(t, t), frozen=(mode[0], mode[1], counts))
                noise = None
                if has_noise:
                    noise = torch.randn(batch_size, *self.obj_size, device=dev).view(batch_size, d)[lo:hi].contiguous()
                self.backend.sampler_step(x0_hat, flat, noise, 0.0, 0.0, 0.0, out=flat, coef=scal[2:5])

            graph = torch.cuda.CUDAGraph()
            if self._graph_pool is None:                 # one memory pool for all graphs of this sampler: they replay one
                self._graph_pool = torch.cuda.graph_pool_handle()      # after the other, their scratch can overlap
            try:
                torch.cuda.synchronize(dev)
                with torch.cuda.graph(graph, pool=self._graph_pool):
                    body()
            except Exception as exc:                     # noqa: BLE001  (capture is an optimisation: fall back, say so once)
                import warnings
                warnings.warn(f"pdm_b200.IdealSampler: CUDA-graph capture failed ({type(exc).__name__}: {exc}); "
                              "sampling continues without graphs")
                self.use_graphs = False
                return False
            entry = (graph, scal, counts)
            self._graphs[key] = entry
        graph, scal, counts = entry
        scal.copy_(scal_row, non_blocking=True)           # pinned row of the trajectory's table: a true async copy
        graph.replay()
        self.graph_replays += 1
        if mode[0]:                                       # screening counts of this step -> the engine's lagged feedback
            ev = torch.cuda.Event()
            ev.record()
            rpt = getattr(self.backend, "row_tile", None) or 128 * (self.engine.cfg.cta_group or 2)
            self.engine._pm_pending.append((ev, counts, (rows + rpt - 1) // rpt, rows, t, t))
        return True

    @torch.no_grad()
    def batch_sample(self, batch_size: int, track_states: bool = False, x_init: Optional[Tensor] = None) -> dict[str, Tensor]:
        """``x_init``: start from this state (shape (batch_size, *obj_size), at the noise level of the last entry of the
        schedule) instead of drawing it -- for trajectory slices."""
        dev = self.backend.device
        d = self.engine.ds.d
        lo, hi = self._my_rows(batch_size)
        drawn = x_init is None
        if drawn:
            x_init = torch.randn(batch_size, *self.obj_size, device=dev)       # ddpm_sampling.py:114
        rows = hi - lo
        graphs = self.use_graphs and dev.type == "cuda" and not self.dataset_sharded and rows > 0
        if graphs:
            # the graphs are tied to the address of the state they update: one persistent state buffer per slice shape
            xt = self._graphs.get(("state", rows))
            if xt is None:
                xt = torch.empty(rows, *self.obj_size, dtype=torch.float32, device=dev)
                self._graphs[("state", rows)] = xt
            xt.copy_(x_init.to(device=dev, dtype=torch.float32)[lo:hi])
        else:
            xt = x_init.to(device=dev, dtype=torch.float32)[lo:hi]
            if not drawn or self.q_world > 1:
                xt = xt.clone()                                               # the loop updates its state in place
        states = [] if track_states else None
        flat = xt.view(rows, d)
        table = None
        if graphs:                                        # (t, 1/sqrt(ab), c_x0, c_xt, c_noise) of every step, pinned
            rows_t = []
            for idx in range(len(self.alpha_bar)):
                ab = self.alpha_bar[idx]
                abp = self.alpha_bar[idx - 1] if idx > 0 else 1.0
                c = step_coefficients(ab, abp, self.step_type)
                rows_t.append([(1.0 - ab) / ab, 1.0 / math.sqrt(ab), c[0], c[1],
                               c[2] if (self.step_type == "ddpm" and idx > 0) else 0.0])
            table = torch.tensor(rows_t, dtype=torch.float32).pin_memory()
            self._scal_table = table                      # alive until the copies out of it have run
        for idx in range(len(self.alpha_bar) - 1, -1, -1):
            ab = self.alpha_bar[idx]
            abp = self.alpha_bar[idx - 1] if idx > 0 else 1.0         # clean_log_temp = -inf -> alpha_bar = 1
            last = idx == 0
            if not (graphs and self.use_graphs and self._step_graphed(flat, batch_size, lo, hi, ab, abp, last, table[idx])):
                self._step_eager(flat, batch_size, lo, hi, ab, abp, last)
            if states is not None:
                states.append(self._gather_rows(xt, batch_size, copy=True))
        res = {"x": self._gather_rows(xt, batch_size, copy=graphs)}
        if states is not None:
            res["states"] = torch.stack(states[::-1])
        return res

    def _gather_rows(self, mine: Tensor, batch_size: int, copy: bool) -> Tensor:
        """This rank's trajectories -> the whole batch on every rank (ranks hold ceil(B / world) rows, the last fewer)."""
        if self.q_world == 1:
            return mine.clone() if copy else mine
        import torch.distributed as dist
        per = (batch_size + self.q_world - 1) // self.q_world
        send = torch.zeros(per, *self.obj_size, dtype=mine.dtype, device=mine.device)
        send[:mine.shape[0]] = mine
        recv = torch.empty(self.q_world * per, *self.obj_size, dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(recv, send, group=self.query_group)
        return recv[:batch_size]

    @torch.no_grad()
    def sample(self, n_samples: int, batch_size: int, track_states: bool = False) -> dict[str, Tensor]:
        """``n_samples`` samples in batches of ``batch_size`` (ddpm_sampling.py:131-139); results stay on the device."""
        parts: dict[str, list[Tensor]] = {}
        done = 0
        while done < n_samples:
            b = min(batch_size, n_samples - done)
            for k, v in self.batch_sample(b, track_states).items():
                parts.setdefault(k, []).append(v)
            done += b
        return {k: torch.cat(v, dim=1 if k == "states" else 0) for k, v in parts.items()}
