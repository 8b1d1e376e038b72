"""Data-parallel batch sharding: one process per GPU, gradients allreduced over NCCL (NVLink 5 /
NVSwitch) in a few contiguous fp32 buckets, launched on a side stream as soon as the last gradient
of a bucket has been produced by the backward pass (reverse layer order), with the 1/G scale
folded into the reduction (ReduceOp.AVG).  The reference is single-GPU (the only trace is the
commented-out nn.DataParallel at reference main.py:174); per-rank math = the reference on that
rank's shard, BatchNorm statistics stay local (plain DistributedDataParallel behaviour),
SURVEY §8e.  There is no data-path collective: graphs are independent units.

Works with any torch.distributed backend ('nccl' on GPUs; 'gloo' in the CPU tests of the bucket
logic, where AVG is emulated by SUM + scale).  With `direct=True` (CUDA ranks) the reductions are issued through
graphtrans_b200.nccl on the comm stream, which makes them capturable: GraphedStep records one ncclAllReduce node per
bucket inside the CUDA graph of the step, forked off the backward at the point where the bucket's last gradient has
been produced, so the transfer overlaps the rest of the backward on every replay.

Constraints of overlap=True (documented in INTEGRATION.md): exactly ONE backward per zero_grad() - with gradient
accumulation or FLAG's m > 1 ascent steps (reference trainers/flag_trainer.py) wrap all but the last backward in
`no_sync()`, which keeps the buckets un-armed.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, model: torch.nn.Module, n_buckets: int = 4, overlap: bool = True, group=None, direct: bool = False):
        self.group = group
        self._sync_enabled = True
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        # every gradient starts on a 128-byte boundary of the arena: the split-K weight-gradient GEMM accumulates through
        # TMA bulk reductions / 16-byte vector atomics, which a 4-byte-aligned view (anything laid out after GIN's
        # one-element eps or a [5002] head bias) would degrade to one scalar atomic per element
        ALIGN = 32
        total = sum((p.numel() + ALIGN - 1) // ALIGN * ALIGN for p in self.params)
        # one flat fp32 gradient arena; every p.grad is a view into it (no copies around the allreduce)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        # bucket boundaries: parameters() order is forward order, so the LAST bucket completes first
        target = (total + n_buckets - 1) // n_buckets
        self.buckets = []          # (lo, hi, [param indices])
        lo, cur, idxs = 0, 0, []
        off = 0
        self._bucket_of = {}
        self.offsets = []          # arena offset (elements) of every parameter's gradient view
        for i, p in enumerate(self.params):
            self.offsets.append(off)
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            # fused gradient delivery (ops._grad_target): the backward kernels accumulate straight into the arena
            p._gt_main_grad = p.grad
            p._gt_uses = getattr(p, "_gt_uses", 1)
            p._gt_grad_ready = self._make_ready(i)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            cur += p.numel()
            idxs.append(i)
            if cur >= target and len(self.buckets) < n_buckets - 1:
                self.buckets.append((lo, off, idxs))
                lo, cur, idxs = off, 0, []
        if idxs:
            self.buckets.append((lo, off, idxs))
        for b, (_, _, ids) in enumerate(self.buckets):
            for i in ids:
                self._bucket_of[i] = b
        self.overlap = overlap and self.world > 1 and dev.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        # direct NCCL communicator (capturable collectives on an explicit stream)
        self.nccl = None
        if direct and self.world > 1 and dev.type == "cuda":
            from . import nccl
            self.nccl = nccl.Communicator(group)
        self._pending = [0] * len(self.buckets)
        self._work = []
        self._launched = [False] * len(self.buckets)
        self._fused_seen = [False] * len(self.params)     # parameters whose gradients the backward kernels deliver themselves
        if self.overlap:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))
        self.reset()

    # ------------------------------------------------------------------
    def _make_ready(self, i):
        def ready():
            if not self._sync_enabled:
                return
            self._fused_seen[i] = True
            self._uses_left[i] -= 1
            if self._uses_left[i] == 0 and self.overlap:
                self._count(i)
        return ready

    def _count(self, i):
        """parameter i has its complete gradient: counted ONCE per step, whichever notification comes first (the fused
        delivery's `_gt_grad_ready` or autograd's post-accumulate hook - the engine runs the AccumulateGrad node, and with
        it the hook, even when the backward kernels delivered the gradient themselves and returned None)"""
        if self._counted[i]:
            return
        self._counted[i] = True
        b = self._bucket_of[i]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def reset(self):
        self._uses_left = [getattr(p, "_gt_uses", 1) for p in self.params]
        self._pending = [len(ids) for _, _, ids in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._counted = [False] * len(self.params)
        self._work = []

    def zero_grad(self):
        """one memset for every gradient (and re-arm the hooks)"""
        self.flat.zero_()
        self.reset()

    def _make_hook(self, i):
        def hook(_p):
            if not self._sync_enabled:
                return
            if self._uses_left[i] > 0 and getattr(self.params[i], "_gt_main_grad", None) is not None and self._fused_seen[i]:
                return      # fused delivery in progress for this parameter: its own notifications decide
            self._count(i)
        return hook

    def _launch(self, b):
        if self._launched[b] or self.world == 1:
            return
        self._launched[b] = True
        if __import__("os").environ.get("GT_DDP_DEBUG"):
            import threading
            print(f"[ddp] launch bucket {b} thread={threading.current_thread().name} stream={torch.cuda.current_stream().cuda_stream:#x} "
                  f"capturing={torch.cuda.is_current_stream_capturing()} pending={self._pending} "
                  f"undone={[i for i, u in enumerate(self._uses_left) if u > 0][:8]}", flush=True)
        lo, hi, _ = self.buckets[b]
        chunk = self.flat[lo:hi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            # gradients of this bucket may still be queued on the weight-gradient / virtual-node streams (the main
            # stream only joins them after backward()): the reduction has to wait for those as well
            from . import ops
            for side in (ops._wgrad, ops._branch):
                if side["stream"] is not None and side["used"]:
                    self.comm_stream.wait_stream(side["stream"])
            with torch.cuda.stream(self.comm_stream):
                self._allreduce(chunk)
        else:
            self._allreduce(chunk)

    def no_sync(self):
        """context manager: backward passes inside it do not arm / launch the bucket reductions (gradient accumulation,
        FLAG ascent steps); the last backward outside it reduces the accumulated gradients"""
        outer = self

        class _NoSync:
            def __enter__(self):
                self.prev, outer._sync_enabled = outer._sync_enabled, False

            def __exit__(self, *exc):
                outer._sync_enabled = self.prev
                return False
        return _NoSync()

    def _allreduce(self, chunk):
        if self.nccl is not None:
            self.nccl.all_reduce(chunk, avg=True)        # on torch's current stream (= comm stream here); capturable
        elif dist.get_backend(self.group) == "nccl":
            self._work.append(dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            chunk.div_(self.world)

    def finish(self):
        """call after backward(): launches whatever the hooks did not (parameters without a gradient
        this step, or overlap disabled) and makes the compute stream wait for the reductions."""
        if self.world == 1:
            return
        for b in reversed(range(len(self.buckets))):
            self._launch(b)
        for w in self._work:
            w.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.reset()       # re-arm: a CUDA-graph replay of the step does not re-run zero_grad() on the host


def shard_range(n_graphs: int, rank: int, world: int):
    """contiguous graph range of `rank` (B/G graphs per rank, remainder to the first ranks)"""
    base, rem = divmod(n_graphs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
