"""``DDPMTrue``: the closed-form empirical (ideal) denoiser behind the reference's DDPM interface
(diffusion/ddpm/true_model.py:6-12: same constructor, same ``train_data`` buffer, same ``forward(xt, tau)``).

``train_data`` stays a registered buffer, so ``.to(device)`` and ``ddpm.train_data.device``
(scripts/optimize_schedule.py:55) keep working; the engine keeps its own resident copy of it with row norms and operand
splits, built on first use and cached per buffer (diffusion/scheduler/scheduler.py: ``_engine_for_data``).  Beyond the
reference's interface the module hands out that engine and a fused sampler on it."""
from __future__ import annotations

from typing import Iterable, Optional

from torch import Tensor

from .ddpm import DDPM


class DDPMTrue(DDPM):
    def __init__(self, scheduler, parametrization: str, train_data: Tensor):
        DDPM.__init__(self, scheduler, parametrization)
        self.register_buffer("train_data", train_data)

    def forward(self, xt: Tensor, tau: Tensor) -> Tensor:
        """x0_hat(xt, tau): posterior mean of the training set under the VP kernel of noise level tau.  With
        PDM_SHARD_QUERIES=1 under torch.distributed the rows of ``xt`` are split over the ranks (dataset replicated)."""
        denoise = self.scheduler.true_posterior_mean_x0
        return denoise(xt, tau, self.train_data)

    # ---- extensions ---------------------------------------------------------------------------------------------------
    @property
    def engine(self):
        """The ``PosteriorEngine`` that serves this module's buffer (resident dataset, screening marks, statistics)."""
        from ..scheduler.scheduler import _engine_for_data
        return _engine_for_data(self.train_data)

    def sampler(self, log_temp: Tensor | Iterable[float], step_type: str = "ddim", query_group: Optional[object] = None):
        """Fused form of ``DDPMSampler`` around this model (pdm_b200.IdealSampler): one posterior-mean call and one update
        kernel per step, CUDA-graphed, optionally with the trajectories split over ``query_group``."""
        from pdm_b200 import IdealSampler
        return IdealSampler(self.train_data, log_temp, step_type=step_type, engine=self.engine, query_group=query_group)
