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

import numpy as np
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


# ----------------------------------------------------------------------------- collate-time integer work
_CSR_FIELDS = ("csr_rowptr_dst", "csr_src_by_dst", "csr_eid_by_dst", "csr_rowptr_src", "csr_dst_by_src", "csr_eid_by_src")


def attach_csr(batch: GraphBatch) -> GraphBatch:
    """Build both int32 CSRs of `edge_index` on the HOST (loader worker time, numpy stable counting sort) and attach them
    as `csr_*` fields: bit-identical to what gt_csr_build produces on the device (rows in edge-id order), so the model
    skips the per-step device sort (ops.plan_for).  Returns the same batch object."""
    ei = batch.edge_index.cpu().numpy()
    N = int(batch.batch.numel())
    src, dst = ei[0], ei[1]

    def one(key, other):
        order = np.argsort(key, kind="stable").astype(np.int32)
        rowptr = np.zeros(N + 1, np.int32)
        np.cumsum(np.bincount(key, minlength=N), out=rowptr[1:])
        nbr = other[order].astype(np.int32)
        if order.size == 0:                                   # kernels expect at least one element to point at
            order, nbr = np.zeros(1, np.int32), np.zeros(1, np.int32)
        return torch.from_numpy(rowptr), torch.from_numpy(nbr), torch.from_numpy(order)

    rp_d, s_by_d, e_by_d = one(dst, src)
    rp_s, d_by_s, e_by_s = one(src, dst)
    for k, v in zip(_CSR_FIELDS, (rp_d, s_by_d, e_by_d, rp_s, d_by_s, e_by_s)):
        setattr(batch, k, v)
    return batch


def bucket_size(n: int, log2_steps: int = 6) -> int:
    """smallest bucket > n on a grid whose step is 2^-log2_steps .. 2^-(log2_steps-1) of the size (default 1/64 .. 1/32:
    <= 3 % slack, 1.5 % on average): a few buckets cover the batches of an epoch (N and E of a batch of B iid graphs
    vary by a few percent), so CUDA-graph signatures repeat"""
    n = int(n) + 1
    q = 1 << max(int(n).bit_length() - int(log2_steps), 3)
    return (n + q - 1) // q * q


class Bucketer:
    """Shape-bucket grid for a stream of batches.  The default grid (1/64 .. 1/32 steps) suits batches of many small
    graphs (molecules: N varies by ~1 %); datasets with heavy-tailed graph sizes (Code2: N of a 128-graph batch varies by
    +-20 %) would hit a new bucket almost every step, so `fit` coarsens the grid from a sample of (N, E) pairs until about
    `target` distinct buckets cover the sample - trading a few percent of slack rows for captured graphs that repeat."""

    def __init__(self, log2_steps: int = 6):
        self.log2_steps = int(log2_steps)
        self.fixed = None

    def fit(self, shapes, target: int = 6, single_bucket_spread: float = 0.06):
        """shapes: iterable of (n_nodes, n_edges) of sample batches (collate-time ints).  When the sample's spread is
        small (max / min - 1 <= single_bucket_spread: large batches of iid graphs) ONE bucket just above the largest
        sample serves (nearly) every batch - a few percent of slack rows, no recapture; batches beyond it fall back to
        the grid."""
        shapes = list(shapes)
        n_lo, n_hi = min(s[0] for s in shapes), max(s[0] for s in shapes)
        e_lo, e_hi = min(s[1] for s in shapes), max(s[1] for s in shapes)
        self.fixed = None
        if n_hi <= (1 + single_bucket_spread) * n_lo and e_hi <= (1 + single_bucket_spread) * max(e_lo, 1):
            margin = 1 + single_bucket_spread / 4
            self.fixed = (bucket_size(int(n_hi * margin), 7), bucket_size(int(e_hi * margin), 7))
            return self
        for k in (6, 5, 4, 3):
            self.log2_steps = k
            if len({self.bucket(n, e) for n, e in shapes}) <= target:
                break
        return self

    def bucket(self, n_nodes, n_edges):
        if self.fixed is not None and n_nodes < self.fixed[0] and n_edges <= self.fixed[1]:
            return self.fixed
        return bucket_size(n_nodes, self.log2_steps), bucket_size(n_edges, self.log2_steps)

    def pad(self, batch: "GraphBatch") -> "GraphBatch":
        n, e = self.bucket(int(batch.batch.numel()), int(batch.edge_index.shape[1]))
        return pad_to_bucket(batch, n, e)


def pad_to_bucket(batch: GraphBatch, n_nodes: int = None, n_edges: int = None) -> GraphBatch:
    """Append slack nodes / edges so that the batch has exactly `n_nodes` nodes and `n_edges` edges (default: the next
    buckets, `bucket_size`).  Slack nodes carry batch id == num_graphs, i.e. they belong to NO graph: they get no token
    row, no virtual-node state, stay out of the BatchNorm statistics (GraphPlan.m_valid) and receive exactly zero
    gradients, so logits, loss and parameter gradients equal those of the unpadded batch.  Slack edges are self loops
    spread over the slack nodes.  Host-side (collate time); any `csr_*` fields are rebuilt."""
    if getattr(batch, "slack", False):
        raise ValueError("pad_to_bucket: batch is already padded")
    N, E, B = int(batch.batch.numel()), int(batch.edge_index.shape[1]), int(batch.num_graphs)
    n_nodes = bucket_size(N) if n_nodes is None else int(n_nodes)
    n_edges = bucket_size(E) if n_edges is None else int(n_edges)
    if n_nodes <= N or n_edges < E:
        raise ValueError(f"pad_to_bucket: bucket ({n_nodes}, {n_edges}) does not exceed the batch ({N}, {E})")
    dn, de = n_nodes - N, n_edges - E
    out = {k: v for k, v in batch.__dict__.items() if not k.startswith("csr_") and not k.startswith("_")}

    def grow(t, extra, fill=0):
        return torch.cat([t, t.new_full((extra,) + tuple(t.shape[1:]), fill)], dim=0)

    for k in _NODE_FIELDS:
        v = out.get(k)
        if v is not None:
            out[k] = grow(v, dn)
    out["batch"] = grow(batch.batch, dn, B)
    loops = N + (torch.arange(de, dtype=batch.edge_index.dtype) % dn)
    out["edge_index"] = torch.cat([batch.edge_index, torch.stack([loops, loops])], dim=1)
    if out.get("edge_attr") is not None:
        out["edge_attr"] = grow(batch.edge_attr, de)
    out["slack"] = True
    padded = GraphBatch(**out)
    if getattr(batch, _CSR_FIELDS[0], None) is not None:
        attach_csr(padded)
    return padded


def pack(batch: GraphBatch, pin: bool = True) -> GraphBatch:
    """Lay every tensor of a host batch out in ONE (pinned) byte blob, 256-byte aligned: `packed.to(device)` is then a
    single H2D copy instead of one per field (trainers/base_trainer.py:23 `batch.to(device)` issues 7-9), and a captured
    step refreshes its static inputs with one copy."""
    layout, off = [], 0
    for name, t in batch.tensors():
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        layout.append((name, t.dtype, tuple(t.shape), off, nbytes, t))
        off += (nbytes + 255) // 256 * 256
    blob = torch.empty(max(off, 256), dtype=torch.uint8)
    if pin and torch.cuda.is_available():
        blob = blob.pin_memory()
    out = GraphBatch(**{k: v for k, v in batch.__dict__.items() if not torch.is_tensor(v)})
    for name, dtype, shape, o, nbytes, t in layout:
        blob[o:o + nbytes].copy_(t.view(-1).view(torch.uint8))
        setattr(out, name, blob[o:o + nbytes].view(dtype).view(shape))
    out._blob = blob
    out._layout = [(name, dtype, shape, o, nbytes) for name, dtype, shape, o, nbytes, _ in layout]
    return out


def prepare(batch: GraphBatch, bucket=True, csr: bool = True, blob: bool = True) -> GraphBatch:
    """collate-time pipeline of a host batch: shape-bucket padding -> int32 CSR -> one pinned blob.
    bucket: True (default grid), False, or a `Bucketer` fitted to the dataset"""
    if bucket and not getattr(batch, "slack", False):
        batch = bucket.pad(batch) if isinstance(bucket, Bucketer) else pad_to_bucket(batch)
    if csr and getattr(batch, _CSR_FIELDS[0], None) is None:
        attach_csr(batch)
    return pack(batch) if blob else batch
