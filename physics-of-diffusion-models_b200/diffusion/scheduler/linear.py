"""Linear-beta (DDPM) schedule in the continuous-time parameterisation of the reference
(diffusion/scheduler/linear.py:5-16):  T(tau) = (1 + T_min) * ((1 + T_max)/(1 + T_min))^(tau^2) - 1,
so that T(0) = T_min, T(1) = T_max and alpha_bar = 1/(1 + T) decays like exp(-gamma tau^2) (beta linear in tau).
``temperatures`` / ``alpha_bars`` are host-side tables of a discretised schedule (bench.py, IdealSampler)."""
import math

import torch
from torch import Tensor

from .scheduler import Scheduler


class LinearBetaScheduler(Scheduler):
    def __init__(self, min_temp: float, max_temp: float):
        Scheduler.__init__(self)
        if not (max_temp > min_temp > -1.0):
            raise ValueError(f"need -1 < min_temp < max_temp, got {min_temp}, {max_temp}")
        self.scale = 1 + min_temp                                    # 1 + T(0)
        self.gamma = math.log((1 + max_temp) / self.scale)           # log of the total growth of 1 + T over tau in [0, 1]

    def log_temp_from_tau(self, tau: Tensor) -> Tensor:
        one_plus_t = (tau.pow(2) * self.gamma).exp() * self.scale
        return (one_plus_t - 1).log()

    def tau_from_log_temp(self, log_temp: Tensor) -> Tensor:
        growth = ((log_temp.exp() + 1) / self.scale).log()           # gamma tau^2
        return (growth / self.gamma).sqrt()

    # ---- discretised tables (float64 on the host) --------------------------------------------------------------------
    def temperatures(self, n_steps: int) -> Tensor:
        """T at tau = linspace(0, 1, n_steps + 1)[1:] (the grid of the reference's samplers and statistics scripts)."""
        tau = torch.linspace(0, 1, n_steps + 1, dtype=torch.float64)[1:]
        return (tau.pow(2) * self.gamma).exp() * self.scale - 1

    def alpha_bars(self, n_steps: int) -> Tensor:
        return 1.0 / (1.0 + self.temperatures(n_steps))
