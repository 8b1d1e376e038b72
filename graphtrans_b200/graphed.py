"""CUDA-graph replay of one training step (zero_grad -> forward -> loss -> backward).

The hot path launches a few hundred small kernels per step (ideal step times are fractions of a
millisecond), so in eager mode the Python host - not the GPU - sets the pace.  Every
C-ABI export is asynchronous on the caller's stream, allocation-free and sync-free, the dropout RNG
state and the batch plan live on the device, and parameter gradients are accumulated by the kernels
straight into the flat arena of `ddp.GradBuckets`: the whole step is capturable.  `GraphedStep`
captures it once per batch *shape signature* (N, E, B, feature shapes) into a `torch.cuda.CUDAGraph`
with static input buffers and replays it afterwards; a new signature is captured on first use (LRU
cache).  The gradient allreduce stays outside the graph (`buckets.finish()` after the replay).

Replaces the Python-level orchestration of reference trainers/base_trainer.py:28-33
(`optimizer.zero_grad(); pred = model(batch); loss = calc_loss(pred, batch); loss.backward()`).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

from . import _lib, ops


def _signature(batch):
    sig = []
    for k in sorted(batch.__dict__):
        v = getattr(batch, k)
        if torch.is_tensor(v):
            sig.append((k, tuple(v.shape), str(v.dtype)))
        elif k == "max_nodes":   # only its coarse bucket shapes the launches (ops.token_bucket)
            sig.append((k, None if v is None else -(-(int(v) + 1) // 32)))
        else:
            sig.append((k, v if isinstance(v, (int, float, str, type(None))) else None))
    return tuple(sig)


class _Entry:
    __slots__ = ("graph", "static_batch", "loss", "kernels")


class GraphedStep:
    def __init__(self, model, loss_fn, buckets, max_graphs=16, warmup_iters=2, optimizer=None):
        """optimizer: optional graphtrans_b200.optim.FusedAdamW; on one GPU its step is part of the captured graph, with
        several ranks it runs after the gradient allreduce (which stays outside the graph)"""
        self.model, self.loss_fn, self.buckets = model, loss_fn, buckets
        self.optimizer = optimizer
        self.opt_in_graph = optimizer is not None and getattr(buckets, "world", 1) == 1
        if getattr(buckets, "overlap", False):
            raise ValueError("GraphedStep needs GradBuckets(overlap=False): collectives stay outside the graph")
        self.max_graphs, self.warmup_iters = max_graphs, warmup_iters
        self.cache: "OrderedDict[tuple, _Entry]" = OrderedDict()
        self.pool = None
        self.device = buckets.flat.device
        self.last_kernels = 0
        # the step is captured on a HIGH-priority stream: its kernels are the critical path, the weight-gradient and
        # virtual-node branches (default = lowest priority) only fill the SMs it leaves free.  Captured kernel nodes
        # inherit the priority of the stream they were issued on.
        self.capture_stream = None
        if self.device.type == "cuda" and os.environ.get("GT_MAIN_PRIORITY", "1") == "1":
            self.capture_stream = torch.cuda.Stream(device=self.device, priority=-1)

    def _eager(self, b):
        self.buckets.zero_grad()
        loss = self.loss_fn(self.model(b), b)
        loss.backward()
        ops.join_side_streams()       # weight gradients / virtual-node branch issued on side streams (ops.enable_*)
        if self.opt_in_graph:
            self.optimizer.step()
        return loss.detach()

    def _capture(self, batch, sig):
        ent = _Entry()
        ent.static_batch = batch.to(self.device).clone()       # static input buffers of this signature
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        if self.optimizer is not None:
            self.optimizer.enabled = False                      # warm-up must not move the weights
        with torch.cuda.stream(side):                           # warm-up off the capture (cudaFuncSetAttribute,
            for _ in range(self.warmup_iters):                  # allocator warm-up, lazy module state)
                self._eager(ent.static_batch)
        if self.optimizer is not None:
            self.optimizer.enabled = True
            if self.optimizer._desc is None:
                self.optimizer._build()                          # descriptor table allocated outside the capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        ent.graph = torch.cuda.CUDAGraph()
        k0 = _lib.kernel_count
        with torch.cuda.graph(ent.graph, pool=self.pool, stream=self.capture_stream):
            ent.loss = self._eager(ent.static_batch)
        ent.kernels = _lib.kernel_count - k0
        if self.pool is None:
            self.pool = ent.graph.pool()
        self.cache[sig] = ent
        while len(self.cache) > self.max_graphs:
            self.cache.popitem(last=False)
        return ent

    def __call__(self, batch):
        """batch: GraphBatch on the host (pinned) or on the device.  Returns the (static) loss tensor; the
        gradients are in `buckets.flat` / p.grad after the call."""
        sig = _signature(batch)
        ent = self.cache.get(sig)
        if ent is None:
            ent = self._capture(batch, sig)
        else:
            self.cache.move_to_end(sig)
        sb = ent.static_batch
        for k, v in batch.__dict__.items():
            if torch.is_tensor(v):
                getattr(sb, k).copy_(v, non_blocking=True)
        ent.graph.replay()
        self.last_kernels = ent.kernels
        self.buckets.finish()
        if self.optimizer is not None and not self.opt_in_graph:
            self.optimizer.step()
        return ent.loss
