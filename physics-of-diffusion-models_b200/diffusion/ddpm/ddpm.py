"""``DDPM`` base module and the x0 / eps / score algebra of its predictions (reference interface:
diffusion/ddpm/ddpm.py:12-45 -- ``DDPMPredictions(pred, xt, alpha_bar, parametrization)`` with the attributes ``pred``,
``x0``, ``eps``, ``score``; ``DDPM.get_predictions(xt, log_temp)``).

With  xt = sqrt(ab) x0 + sqrt(1 - ab) eps  and  score = -eps / sqrt(1 - ab)  any one of the three quantities fixes the
other two.  Here the conversions are a table of closures evaluated ON DEMAND and cached (the samplers read one or two of
the three; the ideal-denoiser path only ever needs ``x0``), each written in the reference's order of operations so that
the values are the reference's bit for bit.
"""
from __future__ import annotations

from abc import abstractmethod
from typing import Callable, Dict, Tuple

from torch import Tensor, nn

from ..scheduler import Scheduler, cast_log_temp

_PARAMETRIZATIONS = ("x0", "eps", "score")

# (given, wanted) -> f(pred, xt, ab).  The diagonal is the identity; "score" from "x0" goes through eps, as in the reference.
_CONVERT: Dict[Tuple[str, str], Callable[[Tensor, Tensor, Tensor], Tensor]] = {
    ("x0", "eps"): lambda p, xt, ab: (xt - p * ab.sqrt()) / (1 - ab).sqrt(),
    ("eps", "x0"): lambda p, xt, ab: (xt - p * (1 - ab).sqrt()) / ab.sqrt(),
    ("eps", "score"): lambda p, xt, ab: -p / (1 - ab).sqrt(),
    ("score", "x0"): lambda p, xt, ab: (xt + p * (1 - ab)) / ab.sqrt(),
    ("score", "eps"): lambda p, xt, ab: -p * (1 - ab).sqrt(),
}


class DDPMPredictions:
    """One network (or closed-form) output ``pred`` at noise level ``alpha_bar`` seen as x0, eps and score."""

    def __init__(self, pred: Tensor, xt: Tensor, alpha_bar: Tensor, parametrization: str) -> None:
        if parametrization not in _PARAMETRIZATIONS:
            raise ValueError(f"unknown parametrization: {parametrization}")
        self.pred = pred
        self.parametrization = parametrization
        self._xt, self._ab = xt, alpha_bar
        self._cache: Dict[str, Tensor] = {parametrization: pred}

    def _as(self, wanted: str) -> Tensor:
        have = self._cache.get(wanted)
        if have is None:
            given = self.parametrization
            if (given, wanted) in _CONVERT:
                have = _CONVERT[(given, wanted)](self.pred, self._xt, self._ab)
            else:                                   # x0 -> score: by way of eps
                have = _CONVERT[("eps", wanted)](self._as("eps"), self._xt, self._ab)
            self._cache[wanted] = have
        return have

    @property
    def x0(self) -> Tensor:
        return self._as("x0")

    @property
    def eps(self) -> Tensor:
        return self._as("eps")

    @property
    def score(self) -> Tensor:
        return self._as("score")


class DDPM(nn.Module):
    """A denoiser ``forward(xt, tau)`` in one of the three parametrizations, tied to a noise schedule."""

    def __init__(self, scheduler: Scheduler, parametrization: str):
        super().__init__()
        assert parametrization in _PARAMETRIZATIONS
        self.scheduler = scheduler
        self.parametrization = parametrization

    def get_predictions(self, xt: Tensor, log_temp: Tensor) -> DDPMPredictions:
        sched = self.scheduler
        tau = sched.tau_from_log_temp(log_temp).clip(0, 1)
        ab = cast_log_temp(sched.alpha_bar_from_tau(tau), xt)
        return DDPMPredictions(self(xt, tau), xt, ab, self.parametrization)

    @abstractmethod
    def forward(self, xt: Tensor, tau: Tensor) -> Tensor:
        ...
