"""``DDPM`` base module and the x0 / eps / score algebra (reference: diffusion/ddpm/ddpm.py:12-45)."""
from __future__ import annotations

from abc import abstractmethod

from torch import Tensor, nn

from ..scheduler import Scheduler, cast_log_temp


class DDPMPredictions:
    """Given one of (x0, eps, score) at noise level alpha_bar, derive the other two from
    xt = sqrt(ab) x0 + sqrt(1 - ab) eps and score = -eps / sqrt(1 - ab)."""

    def __init__(self, pred: Tensor, xt: Tensor, alpha_bar: Tensor, parametrization: str) -> None:
        self.pred = pred
        self.parametrization = parametrization
        sig = (1 - alpha_bar).sqrt()
        if parametrization == "x0":
            self.x0 = pred
            self.eps = (xt - pred * alpha_bar.sqrt()) / sig
            self.score = -self.eps / sig
        elif parametrization == "eps":
            self.x0 = (xt - pred * sig) / alpha_bar.sqrt()
            self.eps = pred
            self.score = -self.eps / sig
        elif parametrization == "score":
            self.x0 = (xt + pred * (1 - alpha_bar)) / alpha_bar.sqrt()
            self.eps = -pred * sig
            self.score = pred
        else:
            raise ValueError(f"unknown parametrization: {parametrization}")


class DDPM(nn.Module):
    def __init__(self, scheduler: Scheduler, parametrization: str):
        super().__init__()
        self.scheduler = scheduler
        self.parametrization = parametrization
        assert self.parametrization in ["x0", "eps", "score"]

    def get_predictions(self, xt: Tensor, log_temp: Tensor) -> DDPMPredictions:
        tau = self.scheduler.tau_from_log_temp(log_temp).clip(0, 1)
        alpha_bar = cast_log_temp(self.scheduler.alpha_bar_from_tau(tau), xt)
        return DDPMPredictions(self(xt, tau), xt, alpha_bar, self.parametrization)

    @abstractmethod
    def forward(self, xt: Tensor, tau: Tensor) -> Tensor:
        pass
