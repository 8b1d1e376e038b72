"""Synthetic batched graphs of the shapes BASELINE.json names.

There is no network for OGB / TU datasets, so every measurement and parity test in this repo
runs on seeded synthetic batches that follow the tensor contract the reference's dataset
adapters produce (reference dataset/code.py:117-133, dataset/mol.py:83-85,
dataset/tud.py:65-73, dataset/utils.py:89-141 `augment_edge`).  The batch object is a plain
attribute bag: the model only does attribute access and ``hasattr(batch, "node_depth")``
(reference modules/gnn_module.py:61-62).
"""
from __future__ import annotations

import argparse
from typing import Optional

import numpy as np
import torch

ATOM_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
BOND_DIMS = [5, 6, 2]
CODE2_NUM_NODETYPES = 98
CODE2_NUM_NODEATTRS = 10030
CODE2_MAX_DEPTH = 20
CODE2_NUM_CLASSES = 5002


class GraphBatch:
    """Attribute bag standing in for torch_geometric.data.Batch.

    A *packed* batch (graphtrans_b200.loader.pack) keeps every tensor as a view into ONE byte blob (`_blob`, layout in
    `_layout`): moving it to the device is a single copy of the blob, after which the views are rebuilt."""

    _TENSOR_FIELDS = ("x", "edge_index", "edge_attr", "batch", "node_depth", "y", "y_arr")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def tensors(self):
        """(name, tensor) of the public tensor fields (the blob of a packed batch is not one of them)"""
        return [(k, v) for k, v in self.__dict__.items() if torch.is_tensor(v) and not k.startswith("_")]

    def _map(self, fn):
        blob = self.__dict__.get("_blob")
        out = GraphBatch()
        if blob is not None:
            new_blob = fn(blob)
            for k, v in self.__dict__.items():
                if not torch.is_tensor(v):
                    setattr(out, k, v)
            out._blob = new_blob
            for name, dtype, shape, off, nbytes in self._layout:
                setattr(out, name, new_blob[off:off + nbytes].view(dtype).view(shape))
            return out
        for k, v in self.__dict__.items():
            setattr(out, k, fn(v) if torch.is_tensor(v) else v)
        return out

    def to(self, device, non_blocking=False):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def clone(self):
        return self._map(lambda t: t.clone())

    def nbytes(self):
        blob = self.__dict__.get("_blob")
        if blob is not None:
            return blob.numel()
        return sum(v.numel() * v.element_size() for _, v in self.tensors())


def _undirected_pairs(rng, n_per_graph, pairs_per_graph, offsets, allow_self=True):
    """`pairs_per_graph[i]` random undirected pairs inside graph i, emitted in both directions."""
    total = int(pairs_per_graph.sum())
    gid = np.repeat(np.arange(len(n_per_graph)), pairs_per_graph)
    n = n_per_graph[gid]
    u = (rng.random(total) * n).astype(np.int64)
    v = (rng.random(total) * n).astype(np.int64)
    if not allow_self:
        v = np.where(u == v, (v + 1) % n, v)
    u += offsets[gid]
    v += offsets[gid]
    # per graph: all forward pairs then all reverse pairs would break graph contiguity only in
    # edge order, which the reference does not rely on; keep (u->v, v->u) interleaved.
    src = np.stack([u, v], 1).reshape(-1)
    dst = np.stack([v, u], 1).reshape(-1)
    return src, dst, gid


def _finish(n, **kw):
    B = len(n)
    batch = np.repeat(np.arange(B, dtype=np.int64), n)
    kw["batch"] = torch.from_numpy(batch)
    kw["num_graphs"] = B
    # collate-time metadata (host ints, no device sync needed later): the largest graph decides whether the
    # tile-local attention kernels apply (every graph within one 128-row tile)
    kw["max_nodes"] = int(np.max(n)) if B else 0
    return GraphBatch(**kw)


def gen_nci1(B=32, seed=1):
    """NCI1-like (config 1): n~U{10..50}, floor(1.08 n) undirected pairs, one-hot 37 features."""
    rng = np.random.default_rng(seed)
    n = rng.integers(10, 51, size=B)
    off = np.concatenate([[0], np.cumsum(n)[:-1]])
    src, dst, _ = _undirected_pairs(rng, n, np.floor(1.08 * n).astype(np.int64), off)
    N = int(n.sum())
    x = np.zeros((N, 37), np.float32)
    x[np.arange(N), rng.integers(0, 37, size=N)] = 1.0
    return _finish(
        n, x=torch.from_numpy(x), edge_index=torch.from_numpy(np.stack([src, dst])), edge_attr=None,
        y=torch.from_numpy(rng.integers(0, 2, size=B).astype(np.int64)))


def _mol_features(rng, N, E_pairs):
    x = np.stack([rng.integers(0, d, size=N) for d in ATOM_DIMS], 1).astype(np.int64)
    ea = np.stack([rng.integers(0, d, size=E_pairs) for d in BOND_DIMS], 1).astype(np.int64)
    ea = np.repeat(ea, 2, axis=0)  # same attribute on both directions
    return x, ea


def gen_mol(B=512, seed=0, num_tasks=128, nan_frac=0.6):
    """ogbg-molpcba-like (config 2)."""
    rng = np.random.default_rng(seed)
    n = np.clip(np.rint(rng.normal(26, 6, size=B)), 4, 60).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(n)[:-1]])
    pairs = np.floor(1.08 * n).astype(np.int64)
    src, dst, _ = _undirected_pairs(rng, n, pairs, off)
    x, ea = _mol_features(rng, int(n.sum()), int(pairs.sum()))
    y = rng.integers(0, 2, size=(B, num_tasks)).astype(np.float32)
    y[rng.random((B, num_tasks)) < nan_frac] = np.nan
    return _finish(n, x=torch.from_numpy(x), edge_index=torch.from_numpy(np.stack([src, dst])),
                   edge_attr=torch.from_numpy(ea), y=torch.from_numpy(y))


def gen_syn(B=4096, seed=0, num_tasks=128, nmin=64, nmax=192):
    """Config 4: n~U{64..192}, 2n undirected pairs u!=v, mol-style features."""
    rng = np.random.default_rng(seed)
    n = rng.integers(nmin, nmax + 1, size=B)
    off = np.concatenate([[0], np.cumsum(n)[:-1]])
    pairs = 2 * n
    src, dst, _ = _undirected_pairs(rng, n, pairs, off, allow_self=False)
    x, ea = _mol_features(rng, int(n.sum()), int(pairs.sum()))
    y = rng.integers(0, 2, size=(B, num_tasks)).astype(np.float32)
    y[rng.random((B, num_tasks)) < 0.6] = np.nan
    return _finish(n, x=torch.from_numpy(x), edge_index=torch.from_numpy(np.stack([src, dst])),
                   edge_attr=torch.from_numpy(ea), y=torch.from_numpy(y))


def gen_code2(B=128, seed=0, nmin=8, nmax=2000, mu=4.6, sigma=0.65, max_seq_len=5,
              num_nodetypes=CODE2_NUM_NODETYPES, num_nodeattrs=CODE2_NUM_NODEATTRS,
              num_classes=CODE2_NUM_CLASSES):
    """ogbg-code2-like (configs 3 and 5): random recursive trees + the four `augment_edge` groups."""
    rng = np.random.default_rng(seed)
    n = np.clip(np.rint(rng.lognormal(mu, sigma, size=B)), nmin, nmax).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(n)[:-1]])
    srcs, dsts, attrs = [], [], []
    for i in range(B):
        ni, o = int(n[i]), int(off[i])
        child = np.arange(1, ni, dtype=np.int64)
        parent = np.floor(rng.random(ni - 1) * child).astype(np.int64)
        attributed = np.nonzero(rng.random(ni) < 0.5)[0]
        nt_a, nt_b = attributed[:-1], attributed[1:]
        # order of groups follows reference dataset/utils.py:138-139
        s = np.concatenate([parent, child, nt_a, nt_b]) + o
        d = np.concatenate([child, parent, nt_b, nt_a]) + o
        a = np.concatenate([
            np.tile([0.0, 0.0], (ni - 1, 1)), np.tile([0.0, 1.0], (ni - 1, 1)),
            np.tile([1.0, 0.0], (len(nt_a), 1)), np.tile([1.0, 1.0], (len(nt_a), 1))]).astype(np.float32)
        srcs.append(s), dsts.append(d), attrs.append(a.reshape(-1, 2))
    N = int(n.sum())
    x = np.stack([rng.integers(0, num_nodetypes, size=N), rng.integers(0, num_nodeattrs, size=N)], 1)
    depth = rng.integers(0, 25, size=(N, 1)).astype(np.int64)
    y_arr = rng.integers(0, num_classes, size=(B, max_seq_len)).astype(np.int64)
    return _finish(
        n, x=torch.from_numpy(x.astype(np.int64)),
        edge_index=torch.from_numpy(np.stack([np.concatenate(srcs), np.concatenate(dsts)])),
        edge_attr=torch.from_numpy(np.concatenate(attrs)), node_depth=torch.from_numpy(depth),
        y_arr=torch.from_numpy(y_arr))


def in_degree_histogram(batch: GraphBatch, bins: int) -> torch.Tensor:
    """`deg` histogram as reference dataset/code.py:121-130 / dataset/mol.py:71-79 build it."""
    N = batch.batch.numel()
    d = torch.bincount(batch.edge_index[1], minlength=N)
    return torch.bincount(d, minlength=bins)[:bins].to(torch.long)


# ---------------------------------------------------------------------------------------------
# argparse.Namespace equivalents of the reference's three-stage parse for the BASELINE configs
# ---------------------------------------------------------------------------------------------
_BASE = dict(
    model_type="gnn-transformer", graph_pooling="cls", gnn_type="gcn", gnn_virtual_node=False,
    gnn_dropout=0.0, gnn_num_layer=5, gnn_emb_dim=300, gnn_JK="last", gnn_residual=False,
    d_model=128, nhead=4, dim_feedforward=512, transformer_dropout=0.3, transformer_activation="relu",
    num_encoder_layers=4, max_input_len=1000, transformer_norm_input=True,
    num_encoder_layers_masked=0, transformer_prenorm=False, pos_encoder=False, pretrained_gnn=None,
    freeze_gnn=None, max_seq_len=None,
    aggregators=["mean", "max", "min", "std"], scalers=["identity", "amplification", "attenuation"],
    post_layers=1, add_edge="none", deg=None)

CONFIGS = {
    # reference configs/NCI1/gnn-transformer/no-virtual/gd=128+gdp=0.1+tdp=0.1+l=3+cosine.yml
    "nci1": dict(dataset="nci1", gnn_type="gcn", gnn_emb_dim=128, d_model=128, dim_feedforward=256,
                 num_encoder_layers=3, transformer_dropout=0.1, gnn_dropout=0.1, num_tasks=2, batch_size=32),
    # reference configs/molpcba/gnn-transformer/JK=cat/pooling=cls+gin+norm_input.yml
    "molpcba": dict(dataset="mol", gnn_type="gin", gnn_virtual_node=True, gnn_JK="cat", gnn_dropout=0.3,
                    num_tasks=128, batch_size=512),
    # reference configs/code2/gnn-transformer/JK=cat/pooling=cls+norm_input.yml (d_model=256 per BASELINE)
    "code2": dict(dataset="code2", gnn_type="gcn", gnn_virtual_node=True, gnn_JK="cat", d_model=256,
                  max_seq_len=5, num_tasks=CODE2_NUM_CLASSES, batch_size=128),
    # BASELINE.json configs[3]: synthetic 4 GIN + 4 Tx layers, d=256
    "syn": dict(dataset="syn", gnn_type="gin", gnn_num_layer=4, gnn_emb_dim=256, d_model=256,
                num_tasks=128, batch_size=4096),
    # reference configs/code2/pna-transformer/pooling=cls+norm_input.yml
    "code2-pna": dict(dataset="code2", model_type="pna-transformer", gnn_emb_dim=272, gnn_num_layer=4,
                      gnn_residual=True, gnn_dropout=0.0, max_seq_len=5, num_tasks=CODE2_NUM_CLASSES,
                      batch_size=128),
}


def make_args(name: str, **overrides) -> argparse.Namespace:
    d = dict(_BASE)
    d.update(CONFIGS[name])
    d.update(overrides)
    return argparse.Namespace(**d)


def make_batch(args, B: Optional[int] = None, seed: int = 0) -> GraphBatch:
    B = args.batch_size if B is None else B
    ds = args.dataset
    if ds == "nci1":
        return gen_nci1(B, seed)
    if ds == "mol":
        return gen_mol(B, seed, num_tasks=args.num_tasks)
    if ds == "syn":
        return gen_syn(B, seed, num_tasks=args.num_tasks)
    if ds == "code2":
        return gen_code2(B, seed, max_seq_len=args.max_seq_len, num_classes=args.num_tasks)
    raise ValueError(ds)
