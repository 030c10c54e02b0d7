"""FusedAdam: torch.optim.Adam's update (train.py:39; the reference builds Adam(mod.parameters(), lr) at TwoWL_work.py:100) as ONE
kernel over a flat parameter buffer (twowl_adam_step) instead of the foreach implementation's chain of launches - the TwoWL models
have ~20 small parameter tensors (1e4-1e5 floats in all), so the optimiser step of the small configurations is pure launch
overhead. The parameters are re-pointed at slices of one flat fp32 buffer (their values are kept), the gradients are gathered with
one torch.cat per step, the step count lives in device memory: the whole update is capturable in a CUDA graph.

Same defaults and semantics as torch.optim.Adam(lr, betas, eps, weight_decay) with amsgrad = False; state_dict() is not
interchangeable with torch's (flat buffers).
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch

from . import ops


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError("FusedAdam: fp32 parameters on one CUDA device required (no CPU fallback exists)")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.param_groups = [{"params": self.params, "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}]
        sizes = [p.numel() for p in self.params]
        self.flat = torch.cat([p.detach().reshape(-1) for p in self.params]).contiguous()
        off = 0
        for p, n in zip(self.params, sizes):
            p.data = self.flat[off:off + n].view_as(p)          # the module now reads / the kernel now writes the same memory
            off += n
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.grad_scale = None

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params])
        lr = self.param_groups[0]["lr"]
        ops.adam_step(self.flat, g, self.exp_avg, self.exp_avg_sq, self.step_count, lr, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, self.grad_scale)

    def state_dict(self):
        return {"flat": self.flat, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_count,
                "lr": self.param_groups[0]["lr"], "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.flat.copy_(sd["flat"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count.copy_(sd["step"])
