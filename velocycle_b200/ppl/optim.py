"""``pyro.optim`` subset: ``ClippedAdam`` and ``Adam`` with Pyro's one-optimizer-per-parameter wrapper.

``ClippedAdam`` restates ``pyro/optim/clipped_adam.py`` (Pyro 1.8.6): per step ``lr *= lrd`` first, then the
gradient is clamped elementwise to ``[-clip_norm, clip_norm]`` and a standard bias-corrected Adam update is
applied.  Configured by the tutorials as ``{"lr": 0.03, "lrd": (0.005/0.03)**(1/num_steps), "betas": (0.8, 0.99)}``
(``tutorials/Tutorial_Capolupo_HumanFibroblasts_OneSample.ipynb`` cell 27).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, Union

import torch
from torch.optim import Optimizer

__all__ = ["ClippedAdam", "Adam", "PyroOptim", "TorchClippedAdam"]


class TorchClippedAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_norm=10.0, lrd=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip_norm=clip_norm, lrd=lrd)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            group["lr"] *= group["lrd"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.data
                grad.clamp_(-group["clip_norm"], group["clip_norm"])
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(grad)
                    state["exp_avg_sq"] = torch.zeros_like(grad)
                exp_avg, exp_avg_sq = state["exp_avg"], state["exp_avg_sq"]
                beta1, beta2 = group["betas"]
                state["step"] += 1
                if group["weight_decay"] != 0:
                    grad = grad.add(p.data, alpha=group["weight_decay"])
                exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
                exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                denom = exp_avg_sq.sqrt().add_(group["eps"])
                bias_correction1 = 1 - beta1 ** state["step"]
                bias_correction2 = 1 - beta2 ** state["step"]
                step_size = group["lr"] * math.sqrt(bias_correction2) / bias_correction1
                p.data.addcdiv_(exp_avg, denom, value=-step_size)
        return loss


class PyroOptim:
    """Callable on an iterable of (unconstrained) parameters; lazily creates one torch optimizer per tensor."""

    def __init__(self, optim_constructor: Callable, optim_args: Union[Dict, Callable], clip_args=None):
        self.pt_optim_constructor = optim_constructor
        self.pt_optim_args = optim_args
        self.optim_objs: Dict[torch.Tensor, Optimizer] = {}

    def __call__(self, params: Iterable[torch.Tensor], *args, **kwargs) -> None:
        for p in params:
            if p not in self.optim_objs:
                a = self.pt_optim_args
                a = a() if callable(a) else a
                self.optim_objs[p] = self.pt_optim_constructor([p], **a)
            self.optim_objs[p].step(*args, **kwargs)

    def get_state(self):
        return {i: o.state_dict() for i, o in enumerate(self.optim_objs.values())}


def ClippedAdam(optim_args, clip_args=None) -> PyroOptim:
    return PyroOptim(TorchClippedAdam, optim_args, clip_args)


def Adam(optim_args, clip_args=None) -> PyroOptim:
    return PyroOptim(torch.optim.Adam, optim_args, clip_args)
