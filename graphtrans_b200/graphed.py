"""CUDA-graph replay of one training step (zero_grad -> forward -> loss -> backward [-> allreduce] [-> AdamW]).

The hot path launches a few hundred small kernels per step (ideal step times are fractions of a
millisecond), so in eager mode the Python host - not the GPU - sets the pace.  Every
C-ABI export is asynchronous on the caller's stream, allocation-free and sync-free, the dropout RNG
state and the batch plan live on the device, and parameter gradients are accumulated by the kernels
straight into the flat arena of `ddp.GradBuckets`: the whole step is capturable.  `GraphedStep`
captures it once per batch *shape signature* into a `torch.cuda.CUDAGraph` with static input buffers
and replays it afterwards.

Real loaders yield a different (N, E) almost every step, so with `bucket=True` a batch is first padded
up to a shape bucket (`loader.pad_to_bucket`: slack nodes / edges that belong to no graph and provably
do not change logits, loss or gradients), gets its int32 CSR built at collate time and travels as one
pinned blob (`loader.prepare`): a handful of signatures then covers an epoch, and a replay refreshes
its static inputs with a single H2D copy.

With several ranks and `GradBuckets(direct=True, overlap=True)` the bucketed `ncclAllReduce`s are part
of the captured graph (forked off the backward where each bucket completes), and the fused AdamW update
follows inside the same graph.

Replaces the Python-level orchestration of reference trainers/base_trainer.py:28-39
(`optimizer.zero_grad(); pred = model(batch); loss = calc_loss(pred, batch); loss.backward();
optimizer.step()`).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

from . import _lib, ops


def _signature(batch):
    sig = []
    for k in sorted(batch.__dict__):
        if k.startswith("_"):
            continue
        v = getattr(batch, k)
        if torch.is_tensor(v):
            sig.append((k, tuple(v.shape), str(v.dtype)))
        elif k == "max_nodes":   # only its coarse bucket shapes the launches (ops.token_bucket: 32 / 64 / 96 / 128 / none)
            sig.append((k, None if v is None else min(-(-(int(v) + 1) // 32), 5)))
        else:
            sig.append((k, v if isinstance(v, (int, float, str, bool, type(None))) else None))
    return tuple(sig)


class _Entry:
    __slots__ = ("graph", "static_batch", "loss", "kernels", "staging", "stage_idx", "consumed", "sig")


class LossHandle:
    """loss of one step on its way to the host: `item()` waits for that step's device->host copy only"""
    __slots__ = ("buf", "ev")

    def __init__(self, buf, ev):
        self.buf, self.ev = buf, ev

    def item(self) -> float:
        self.ev.synchronize()
        return float(self.buf)


class GraphedStep:
    def __init__(self, model, loss_fn, buckets, max_graphs=16, warmup_iters=2, optimizer=None, bucket=False):
        """optimizer: optional graphtrans_b200.optim.FusedAdamW; its step is part of the captured graph whenever the
        gradient reduction is (one GPU, or GradBuckets(direct=True)); otherwise it runs after the host-issued allreduce.
        bucket: pad HOST batches to shape buckets + collate-time CSR + one pinned blob (loader.prepare) before use."""
        self.model, self.loss_fn, self.buckets = model, loss_fn, buckets
        self.optimizer = optimizer
        self.world = getattr(buckets, "world", 1)
        self.comm_in_graph = self.world > 1 and getattr(buckets, "nccl", None) is not None
        self.opt_in_graph = optimizer is not None and (self.world == 1 or self.comm_in_graph)
        if getattr(buckets, "overlap", False) and not self.comm_in_graph:
            raise ValueError("GraphedStep with overlap needs GradBuckets(direct=True): only the direct NCCL binding is "
                             "captured into the graph; torch.distributed collectives stay outside (overlap=False)")
        self.max_graphs, self.warmup_iters = max_graphs, warmup_iters
        self.bucket = bucket                  # False | True (default grid) | loader.Bucketer fitted to the dataset
        self.cache: "OrderedDict[tuple, _Entry]" = OrderedDict()
        self.pool = None
        self.device = buckets.flat.device
        self.last_kernels = 0
        self.captures = 0
        # the step is captured on a HIGH-priority stream: its kernels are the critical path, the weight-gradient and
        # virtual-node branches (default = lowest priority) only fill the SMs it leaves free.  Captured kernel nodes
        # inherit the priority of the stream they were issued on.
        self.capture_stream = None
        if self.device.type == "cuda" and os.environ.get("GT_MAIN_PRIORITY", "1") == "1":
            self.capture_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._staged = {}
        self._loss_ring, self._ring_i = None, 0

    def _eager(self, b):
        self.buckets.zero_grad()
        loss = self.loss_fn(self.model(b), b)
        loss.backward()
        ops.join_side_streams()       # weight gradients / virtual-node branch issued on side streams (ops.enable_*)
        if self.comm_in_graph:
            self.buckets.finish()     # captured: joins the comm stream (buckets the hooks did not launch go last)
        if self.opt_in_graph:
            self.optimizer.step()
        return loss.detach()

    def _snapshot(self):
        """state the warm-up iterations must not move: BatchNorm running statistics / num_batches_tracked and the
        dropout step counter (the optimizer is disabled separately)"""
        bufs = [b for b in self.model.buffers()]
        return [b.clone() for b in bufs], ops.rng_state(self.device).clone()

    def _restore(self, snap):
        saved, rng = snap
        with torch.no_grad():
            for b, s in zip(self.model.buffers(), saved):
                b.copy_(s)
            ops.rng_state(self.device).copy_(rng)

    def _capture(self, batch, sig):
        ent = _Entry()
        ent.staging, ent.stage_idx, ent.consumed, ent.sig = [None, None], 0, [None, None], sig
        ent.static_batch = batch.to(self.device).clone()       # static input buffers of this signature
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        if self.optimizer is not None:
            self.optimizer.enabled = False                      # warm-up must not move the weights
        snap = self._snapshot() if os.environ.get("GT_DBG_NO_SNAPSHOT") != "1" else None
        sync_prev = getattr(self.buckets, "_sync_enabled", True)
        if self.comm_in_graph:
            self.buckets._sync_enabled = False                  # warm-up: local only (no collective outside the graph)
        with torch.cuda.stream(side):                           # warm-up off the capture (cudaFuncSetAttribute,
            for _ in range(self.warmup_iters):                  # allocator warm-up, lazy module state)
                self.buckets.zero_grad()
                loss = self.loss_fn(self.model(ent.static_batch), ent.static_batch)
                loss.backward()
                ops.join_side_streams()
                # drop the autograd graph NOW: a live graph keeps the parameters' AccumulateGrad nodes (created here, on
                # the warm-up stream) alive, the capture would reuse them and the engine's end-of-backward sync with
                # that uncaptured stream invalidates the capture (cudaErrorStreamCaptureIsolation)
                del loss
        if self.comm_in_graph:
            self.buckets._sync_enabled = sync_prev
        torch.cuda.current_stream().wait_stream(side)
        if snap is not None:
            self._restore(snap)                                 # running stats / dropout counter as before the warm-up
        if self.optimizer is not None:
            self.optimizer.enabled = True
            if self.optimizer._desc is None:
                self.optimizer._build()                          # descriptor table allocated outside the capture
        torch.cuda.synchronize(self.device)
        ent.graph = torch.cuda.CUDAGraph()
        k0 = _lib.kernel_count
        with torch.cuda.graph(ent.graph, pool=self.pool, stream=self.capture_stream):
            ent.loss = self._eager(ent.static_batch)
        ent.kernels = _lib.kernel_count - k0
        if self.pool is None:
            self.pool = ent.graph.pool()
        self.cache[sig] = ent
        self.captures += 1
        while len(self.cache) > self.max_graphs:
            self.cache.popitem(last=False)
        return ent

    def prepare(self, batch):
        """host batch -> bucket-padded, CSR-carrying, single-blob batch (idempotent)"""
        if not self.bucket or getattr(batch, "_blob", None) is not None:
            return batch
        if batch.batch.is_cuda:
            return batch
        from . import loader
        return loader.prepare(batch, bucket=self.bucket)

    def prefetch(self, batch):
        """start the H2D copy of a HOST batch on the copy stream (overlaps the step that is running); the next
        `step(batch)` with the same object consumes the staged device copy.  The double-buffered loader of SURVEY §8f
        rank 1: replaces the synchronous `batch.to(device)` of reference trainers/base_trainer.py:23."""
        prepared = self.prepare(batch)
        if prepared.batch.is_cuda:
            return
        ent = self.cache.get(_signature(prepared))
        blob = getattr(prepared, "_blob", None)
        if ent is not None and blob is not None and getattr(ent.static_batch, "_blob", None) is not None:
            # steady state: two persistent device staging blobs per signature (no allocation, no allocator sync)
            k = ent.stage_idx = ent.stage_idx ^ 1
            if ent.staging[k] is None:
                ent.staging[k] = torch.empty_like(ent.static_batch._blob)
            with torch.cuda.stream(self.copy_stream):
                if ent.consumed[k] is not None:
                    self.copy_stream.wait_event(ent.consumed[k])   # the step that read this buffer has copied it out
                ent.staging[k].copy_(blob, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            self._staged[id(batch)] = (prepared, (ent, k), ev)
            return
        with torch.cuda.stream(self.copy_stream):
            dev = prepared.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._staged[id(batch)] = (prepared, dev, ev)

    def __call__(self, batch):
        """batch: GraphBatch on the host (pinned) or on the device.  Returns the (static) loss tensor; the
        gradients are in `buckets.flat` / p.grad after the call."""
        staged = self._staged.pop(id(batch), None)
        if staged is not None and isinstance(staged[1], tuple):    # persistent staging blob of a captured signature
            _, (ent, k), ev = staged
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            ent.static_batch._blob.copy_(ent.staging[k], non_blocking=True)
            ent.consumed[k] = torch.cuda.Event()
            ent.consumed[k].record(cur)
            if self.cache.get(ent.sig) is ent:
                self.cache.move_to_end(ent.sig)
            ent.graph.replay()
            return self._after_replay(ent)
        if staged is not None:
            _, batch, ev = staged                                 # device copy issued earlier on the copy stream
            torch.cuda.current_stream().wait_event(ev)
            for _, v in ([("_blob", batch._blob)] if getattr(batch, "_blob", None) is not None else batch.tensors()):
                v.record_stream(torch.cuda.current_stream())
        else:
            batch = self.prepare(batch)
        sig = _signature(batch)
        ent = self.cache.get(sig)
        if ent is None:
            ent = self._capture(batch, sig)
        else:
            self.cache.move_to_end(sig)
        sb = ent.static_batch
        blob = getattr(batch, "_blob", None)
        if blob is not None and getattr(sb, "_blob", None) is not None:
            sb._blob.copy_(blob, non_blocking=True)              # ONE copy refreshes every static input
        else:
            for k, v in batch.tensors():
                getattr(sb, k).copy_(v, non_blocking=True)
        ent.graph.replay()
        return self._after_replay(ent)

    def step_async(self, batch) -> LossHandle:
        """one step whose loss is copied to pinned host memory asynchronously: the caller reads it (`handle.item()`) after
        launching the NEXT step, so the device never idles on the host round trip that `loss.item()` per step
        (reference trainers/base_trainer.py:41-44) costs; every step's loss still reaches the host"""
        loss = self(batch)
        if self._loss_ring is None:
            self._loss_ring = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(8)]
        buf = self._loss_ring[self._ring_i % len(self._loss_ring)]
        self._ring_i += 1
        buf.copy_(loss.detach().reshape(()).float(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return LossHandle(buf, ev)

    def _after_replay(self, ent):
        self.last_kernels = ent.kernels
        if self.world > 1 and not self.comm_in_graph:
            self.buckets.finish()
        elif self.comm_in_graph:
            self.buckets.reset()      # host-side bookkeeping only (the reductions ran inside the graph)
        if self.optimizer is not None and not self.opt_in_graph:
            self.optimizer.step()
        return ent.loss
