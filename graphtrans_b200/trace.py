"""In-situ timeline of one CUDA-graph replay of the training step.

A one-thread %globaltimer stamp kernel is captured after every C-ABI call (graphtrans_b200._lib.start_stamps), on the
stream the call was issued on, so the replayed graph reports where the step time goes with warm caches and the real
overlap between the main stream, the weight-gradient stream, the virtual-node branch and (multi-rank) the NCCL branch.
Interval i = stamp[i] - previous stamp on the same stream = kernel duration + dependent-launch gap + the stamp kernel
itself; `cal_us` (two back-to-back stamps) is that fixed overhead.  Used by bench.py for the IN-STEP roofline figures and
by tools/graph_trace.py."""
from __future__ import annotations

import torch

from . import _lib
from .graphed import GraphedStep


def stamped_profile(model, loss_fn, buckets, batch, optimizer=None, reps=5):
    """-> dict(records=[(name, stream_index, mean interval us, call args)], cal_us, span_us, n_calls, streams).
    Captures its own (stamped) graph of the step on `batch` and replays it `reps` times.  Every rank of a multi-rank
    job has to call it (the graph contains the gradient collectives)."""
    g = GraphedStep(model, loss_fn, buckets, warmup_iters=1, optimizer=optimizer)
    orig = g._eager
    box = {}

    def eager_stamped(bb):
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            _lib.start_stamps(device=g.device)
            _lib.stamp("<begin>")
            _lib.stamp("<begin2>")
        out = orig(bb)
        if capturing:
            _lib.stamp("<end>")
            box["rec"], box["buf"] = _lib.stop_stamps()
        return out

    g._eager = eager_stamped
    g(batch)                                   # capture + first replay
    rec = box["rec"]
    n = len(rec)
    streams = {}
    for _, st, _ in rec:
        streams.setdefault(st, len(streams))
    acc = [0.0] * n
    cal = span = 0.0
    for _ in range(reps):
        g(batch)
        torch.cuda.synchronize()
        buf = box["buf"][:n].cpu().tolist()
        t0 = buf[0]
        cal += (buf[1] - buf[0]) / 1e3
        span += (buf[n - 1] - t0) / 1e3
        last, prev_main = {}, t0
        for i, (_, st, _) in enumerate(rec):
            s = streams[st]
            base = last.get(s, prev_main)      # first call on a side stream: measured from its fork point
            acc[i] += (buf[i] - base) / 1e3
            last[s] = buf[i]
            if s == 0:
                prev_main = buf[i]
    records = [(rec[i][0], streams[rec[i][1]], acc[i] / reps, rec[i][2]) for i in range(n)]
    return dict(records=records, cal_us=cal / reps, span_us=span / reps, n_calls=n - 3, streams=len(streams))


def kernel_time(profile, names, net=True):
    """(number of calls, summed in-step microseconds) of the C-ABI calls whose export name is in `names`; net = minus the
    stamp kernel's own cost per interval"""
    cal = profile["cal_us"] if net else 0.0
    sel = [r for r in profile["records"] if r[0] in names]
    return len(sel), sum(max(r[2] - cal, 0.2) for r in sel)
