"""Fused AdamW over the flat gradient arena of `ddp.GradBuckets` (one launch for all parameters; device-resident
hyper-parameters and step counter, so the update is CUDA-graph capturable).  Same update rule as
torch.optim.AdamW (reference main.py:178: AdamW(model.parameters(), lr, weight_decay); step at
trainers/base_trainer.py:36).  The reference's optional `clip_grad_norm_(model.parameters(), args.grad_clip)`
(base_trainer.py:34-35) is `max_grad_norm`: one gt_sumsq launch over the arena, the clip coefficient is applied inside the
optimizer's gradient read (the arena itself stays unscaled)."""
from __future__ import annotations

import torch

from ._lib import call, ptr


class FusedAdamW:
    def __init__(self, buckets, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        self.buckets = buckets
        flat = buckets.flat
        if not flat.is_cuda:
            raise RuntimeError("FusedAdamW runs on CUDA parameters only (no CPU fallback)")
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, weight_decay], dtype=torch.float32, device=flat.device)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=flat.device)
        self.enabled = True
        # {max_norm, sum of squares}: device-resident so that a captured step sees set_max_grad_norm() updates
        self.clip = None
        if max_grad_norm is not None and max_grad_norm > 0:
            self.clip = torch.tensor([float(max_grad_norm), 0.0], dtype=torch.float32, device=flat.device)
            self._clip_scratch = torch.zeros(4 * 148 + 1, dtype=torch.float32, device=flat.device)
        self._desc = self._ptrs = None
        self._blocks = 0

    def _build(self):
        recs, blk = [], 0
        for p, off in zip(self.buckets.params, self.buckets.offsets):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdamW needs contiguous fp32 parameters")
            recs.append([p.data_ptr(), off, p.numel(), blk])
            blk += (p.numel() + 2047) // 2048
        self._blocks = blk
        self._ptrs = tuple(p.data_ptr() for p in self.buckets.params)
        self._desc = torch.tensor(recs, dtype=torch.int64, device=self.buckets.flat.device)

    def set_lr(self, lr: float):
        """device-side write: picked up by the next (replayed) step"""
        self.hyper[0:1].fill_(float(lr))

    @torch.no_grad()
    def step(self):
        if not self.enabled:
            return
        if self._desc is None or self._ptrs != tuple(p.data_ptr() for p in self.buckets.params):
            self._build()
        if self.clip is not None:
            call("gt_sumsq", ptr(self.buckets.flat), self.buckets.flat.numel(), self.clip.data_ptr() + 4,
                 ptr(self._clip_scratch), self._clip_scratch.numel())
        call("gt_adamw_multi", ptr(self._desc), self._desc.shape[0], self._blocks, ptr(self.buckets.flat), ptr(self.m),
             ptr(self.v), ptr(self.hyper), ptr(self.step_count), ptr(self.clip))

    def grad_norm(self):
        """total gradient norm of the last step (device scalar; only tracked with max_grad_norm)"""
        return None if self.clip is None else self.clip[1].sqrt()

    def zero_grad(self):
        self.buckets.zero_grad()

    def state_dict(self):
        return {"m": self.m, "v": self.v, "hyper": self.hyper, "step": self.step_count}

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.hyper.copy_(sd["hyper"])
        self.step_count.copy_(sd["step"])
