"""Optimizer / scheduler pieces the K-Planes training step needs.

``FusedAdam`` has torch.optim.Adam's math (the reference uses ``AdamOptimizerConfig(lr=1e-2, eps=1e-12)``,
NS/configs/method_configs.py:546-557, NS/engine/optimizers.py:74-160) executed by ``kp_adam_step``: one
vectorised pass per parameter over param / grad / exp_avg / exp_avg_sq.  ``cosine_decay_factor`` is the
reference's ``CosineDecayScheduler`` lambda (NS/engine/schedulers.py:126-142).
"""
from __future__ import annotations

from typing import Dict, Iterable

import numpy as np
import torch

from .. import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, use_device_hyper: bool = True):
        """``grad_scale`` multiplies every gradient before use (e.g. 1/GradScaler scale, or 1/world_size).
        ``use_device_hyper``: read lr / bias corrections from the group's ``hyper_dev`` device triple when it has one (the
        graph-replayed step keeps them there); False = the host-side values of this call."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            ps, gs, ms, vs = [], [], [], []
            for p in group["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                g = p.grad
                if g.dtype != torch.float32 or g.stride() != p.stride():
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                ps.append(p), gs.append(g), ms.append(st["exp_avg"]), vs.append(st["exp_avg_sq"])
            if ps:
                steps = {self.state[p]["step"] for p in ps}
                if len(steps) != 1:  # parameters joined the optimizer at different times: per-tensor launches
                    for p, g, m, v in zip(ps, gs, ms, vs):
                        ops.adam_step_(p, g, m, v, group["lr"], b1, b2, group["eps"], group["weight_decay"],
                                       self.state[p]["step"], grad_scale)
                else:
                    ops.adam_multi_(ps, gs, ms, vs, group["lr"], b1, b2, group["eps"], group["weight_decay"], steps.pop(),
                                    grad_scale, hyper_dev=group.get("hyper_dev") if use_device_hyper else None)
        return loss


def cosine_decay_factor(step: int, warm_up_end: int, max_steps: int, learning_rate_alpha: float = 0.0) -> float:
    if step < warm_up_end:
        return step / warm_up_end
    progress = (step - warm_up_end) / (max_steps - warm_up_end)
    return float((np.cos(np.pi * progress) + 1.0) * 0.5 * (1 - learning_rate_alpha) + learning_rate_alpha)


class Optimizers:
    """One FusedAdam + cosine LambdaLR per parameter group, keyed like ``Model.get_param_groups()`` /
    ``TrainerConfig.optimizers`` ("proposal_networks", "fields")."""

    def __init__(self, param_groups: Dict[str, list], lr: float = 1e-2, eps: float = 1e-12, warm_up_end: int = 512,
                 max_steps: int = 30000, learning_rate_alpha: float = 0.0) -> None:
        self.optimizers, self.schedulers, self.parameters = {}, {}, param_groups
        for name, params in param_groups.items():
            opt = FusedAdam(params, lr=lr, eps=eps)
            self.optimizers[name] = opt
            self.schedulers[name] = torch.optim.lr_scheduler.LambdaLR(
                opt, lr_lambda=lambda s: cosine_decay_factor(s, warm_up_end, max_steps, learning_rate_alpha))

    def zero_grad_all(self) -> None:
        for opt in self.optimizers.values():
            opt.zero_grad(set_to_none=True)

    def optimizer_step_all(self, grad_scale: float = 1.0, use_device_hyper: bool = True) -> None:
        for opt in self.optimizers.values():
            opt.step(grad_scale=grad_scale, use_device_hyper=use_device_hyper)

    def scheduler_step_all(self, step: int = 0) -> None:
        for sch in self.schedulers.values():
            sch.step()
