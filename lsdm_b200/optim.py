"""Fused AdamW for the SDM training step -- the device-side counterpart of ``torch.optim.AdamW`` that run/train_sdm.py:268 builds
(``AdamW(mp_trainer.master_params, lr=...)``) and diffusion/fp16_util.py:198-214 steps.  Same update rule (decoupled weight decay,
bias-corrected moments, no amsgrad), computed by ``lsdm_adamw_step`` (csrc/train_bw.cu); parameters whose ``.grad`` is None are
skipped like torch does.  With ``process_group`` the gradients are summed over the data-parallel ranks with ONE NCCL all-reduce of
the flattened gradient (2.4 M floats) and divided by the world size inside the update kernel.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import _ptr, _stream


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, process_group=None, all_reduce=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.lib = _lib.load()
        self.all_reduce, self.process_group = all_reduce or process_group is not None, process_group

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        todo = [(g, p) for g in self.param_groups for p in g["params"] if p.grad is not None]
        scale = 1.0
        if self.all_reduce and todo:
            import torch.distributed as dist

            flat = torch.cat([p.grad.reshape(-1) for _, p in todo])
            dist.all_reduce(flat, group=self.process_group)           # the step's one gradient exchange
            scale = 1.0 / dist.get_world_size(self.process_group)
            off = 0
            for _, p in todo:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        for g, p in todo:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise ValueError("FusedAdamW needs contiguous float32 CUDA parameters")
            st = self.state[p]
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
            st["step"] += 1
            grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            b1, b2 = g["betas"]
            with torch.cuda.device(p.device):
                _lib.check(self.lib.lsdm_adamw_step(_ptr(p), _ptr(grad), _ptr(st["exp_avg"]), _ptr(st["exp_avg_sq"]), p.numel(), float(g["lr"]),
                                                    float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]), int(st["step"]), float(scale),
                                                    _stream(p.device)))
            _bump_version(p)   # the kernel wrote through the raw pointer: let SceneDiffusionModel see that its weights changed
        return loss


def _bump_version(p):
    try:
        torch.autograd.graph.increment_version(p)
    except AttributeError:  # pragma: no cover - older torch
        p.add_(0.0)
