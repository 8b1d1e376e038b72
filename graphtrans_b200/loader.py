"""Batch collation and host->device staging for the hot path.

Replaces, for the fields the model reads, what the reference gets from torch_geometric's DataLoader
(reference main.py:149-152: `Batch.from_data_list` concatenates node / edge tensors, offsets `edge_index`, builds the
sorted `batch` vector) and `batch.to(device)` (trainers/base_trainer.py:23).  Two additions matter to the kernels:
  * `max_nodes` - the largest graph of the batch, a HOST int known at collate time for free; it selects the tile-local
    attention kernels without a device->host sync (ops.token_bucket);
  * pinned staging + a copy stream, so the H2D copy of batch i+1 overlaps the step of batch i.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Sequence

import torch

from .synth import GraphBatch

_NODE_FIELDS = ("x", "node_depth")
_EDGE_FIELDS = ("edge_attr",)
_GRAPH_FIELDS = ("y", "y_arr")


def _get(g, k):
    return g.get(k) if isinstance(g, dict) else getattr(g, k, None)


def collate(graphs: Sequence) -> GraphBatch:
    """graphs: objects or dicts with `x [n_i, ...]`, `edge_index [2, e_i]` (local node ids) and optionally
    `edge_attr [e_i, ...]`, `node_depth [n_i, 1]`, `y` / `y_arr` with a leading graph dimension of 1."""
    if not graphs:
        raise ValueError("collate: empty list of graphs")
    n = [int(_get(g, "x").shape[0]) for g in graphs]
    out = {}
    for k in _NODE_FIELDS + _EDGE_FIELDS + _GRAPH_FIELDS:
        vals = [_get(g, k) for g in graphs]
        if all(v is None for v in vals):
            if k == "edge_attr":
                out[k] = None
            continue
        if any(v is None for v in vals):
            raise ValueError(f"collate: field {k!r} is missing on some graphs")
        out[k] = torch.cat(list(vals), dim=0)
    offs = torch.tensor([0] + n[:-1], dtype=torch.long).cumsum(0)
    out["edge_index"] = torch.cat([_get(g, "edge_index") + int(o) for g, o in zip(graphs, offs)], dim=1)
    out["batch"] = torch.repeat_interleave(torch.arange(len(graphs), dtype=torch.long), torch.tensor(n, dtype=torch.long))
    out["num_graphs"] = len(graphs)
    out["max_nodes"] = max(n)
    return GraphBatch(**out)


def split(batch: GraphBatch) -> List[dict]:
    """inverse of collate (host tensors): list of per-graph dicts with local edge ids"""
    B = int(batch.num_graphs)
    counts = torch.bincount(batch.batch, minlength=B)
    node_off = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    eg = batch.batch[batch.edge_index[0]]            # graph of every edge (edges never cross graphs)
    graphs = []
    for g in range(B):
        lo, hi = int(node_off[g]), int(node_off[g + 1])
        em = eg == g
        d = {"edge_index": batch.edge_index[:, em] - lo}
        for k in _NODE_FIELDS:
            v = getattr(batch, k, None)
            if v is not None:
                d[k] = v[lo:hi]
        ea = getattr(batch, "edge_attr", None)
        d["edge_attr"] = None if ea is None else ea[em]
        for k in _GRAPH_FIELDS:
            v = getattr(batch, k, None)
            if v is not None:
                d[k] = v[g:g + 1]
        graphs.append(d)
    return graphs


def shard(batch: GraphBatch, lo: int, hi: int) -> GraphBatch:
    """contiguous graph range [lo, hi) of a collated batch (data-parallel sharding, ddp.shard_range): graphs are
    independent units, so a shard is itself a valid batch"""
    B = int(batch.num_graphs)
    if not 0 <= lo < hi <= B:
        raise ValueError(f"shard: bad graph range [{lo}, {hi}) of {B}")
    counts = torch.bincount(batch.batch, minlength=B)
    node_off = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    n_lo, n_hi = int(node_off[lo]), int(node_off[hi])
    src = batch.edge_index[0]
    em = (src >= n_lo) & (src < n_hi)
    out = {"edge_index": batch.edge_index[:, em] - n_lo, "batch": batch.batch[n_lo:n_hi] - lo, "num_graphs": hi - lo,
           "max_nodes": int(counts[lo:hi].max())}
    for k in _NODE_FIELDS:
        v = getattr(batch, k, None)
        if v is not None:
            out[k] = v[n_lo:n_hi]
    ea = getattr(batch, "edge_attr", None)
    out["edge_attr"] = None if ea is None else ea[em]
    for k in _GRAPH_FIELDS:
        v = getattr(batch, k, None)
        if v is not None:
            out[k] = v[lo:hi]
    return GraphBatch(**out)


class DevicePrefetcher:
    """Iterates device-resident batches one copy ahead: batch i+1 is pinned and copied on a side stream while the
    consumer runs the step of batch i.  Replaces the synchronous `batch.to(device)` of trainers/base_trainer.py:23."""

    def __init__(self, batches: Iterable[GraphBatch], device="cuda"):
        self.batches, self.device = batches, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages onto a CUDA device (no CPU fallback)")
        self.stream = torch.cuda.Stream(device=self.device)

    def _stage(self, hb: GraphBatch):
        pinned = hb if all((not torch.is_tensor(v)) or v.is_pinned() for v in hb.__dict__.values()) else hb.pin_memory()
        with torch.cuda.stream(self.stream):
            db = pinned.to(self.device, non_blocking=True)
        return pinned, db            # the pinned source stays alive until the copy has been consumed

    def __iter__(self) -> Iterator[GraphBatch]:
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            _, cur = nxt
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            for v in cur.__dict__.values():
                if torch.is_tensor(v):
                    v.record_stream(torch.cuda.current_stream(self.device))
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            yield cur
