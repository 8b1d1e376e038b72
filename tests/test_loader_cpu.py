"""Collation (graphtrans_b200/loader.py) against the synthetic generators: split -> collate is the identity on every
field the model reads, shards are valid batches, and the collate-time metadata (num_graphs, max_nodes) is right."""
import pytest
import torch

from graphtrans_b200 import loader, synth
from graphtrans_b200.ddp import shard_range


def _same(a, b):
    for k in ("x", "edge_index", "edge_attr", "batch", "node_depth", "y", "y_arr"):
        va, vb = getattr(a, k, None), getattr(b, k, None)
        assert (va is None) == (vb is None), k
        if va is not None:
            assert va.dtype == vb.dtype and torch.equal(va, vb) or (va.dtype.is_floating_point and torch.equal(torch.nan_to_num(va), torch.nan_to_num(vb))), k
    assert a.num_graphs == b.num_graphs and a.max_nodes == b.max_nodes


@pytest.mark.parametrize("cfg", ["molpcba", "code2", "nci1", "syn"])
def test_split_then_collate_is_identity(cfg):
    args = synth.make_args(cfg)
    b = synth.make_batch(args, B=9, seed=3)
    # the generators emit edges graph by graph, which is also collate's order
    eg = b.batch[b.edge_index[0]]
    assert bool((eg[1:] >= eg[:-1]).all())
    graphs = loader.split(b)
    assert len(graphs) == 9 and sum(g["x"].shape[0] for g in graphs) == b.batch.numel()
    assert all(int(g["edge_index"].max()) < g["x"].shape[0] for g in graphs if g["edge_index"].numel())
    _same(loader.collate(graphs), b)
    assert b.max_nodes == int(torch.bincount(b.batch).max())


def test_shards_partition_the_batch():
    args = synth.make_args("molpcba")
    b = synth.make_batch(args, B=13, seed=1)
    parts = [loader.shard(b, *shard_range(13, r, 4)) for r in range(4)]
    assert sum(p.num_graphs for p in parts) == 13
    assert sum(p.batch.numel() for p in parts) == b.batch.numel()
    assert sum(p.edge_index.shape[1] for p in parts) == b.edge_index.shape[1]
    _same(loader.collate([g for p in parts for g in loader.split(p)]), b)
    for p in parts:
        assert int(p.batch[0]) == 0 and int(p.batch[-1]) == p.num_graphs - 1
        assert int(p.edge_index.min()) >= 0 and int(p.edge_index.max()) < p.batch.numel()
    with pytest.raises(ValueError):
        loader.shard(b, 5, 5)


def test_collate_rejects_inconsistent_graphs():
    with pytest.raises(ValueError):
        loader.collate([])
    g1 = {"x": torch.zeros(2, 3), "edge_index": torch.zeros(2, 0, dtype=torch.long), "edge_attr": torch.zeros(0, 2)}
    g2 = {"x": torch.zeros(1, 3), "edge_index": torch.zeros(2, 0, dtype=torch.long), "edge_attr": None}
    with pytest.raises(ValueError):
        loader.collate([g1, g2])


def test_prefetcher_needs_cuda():
    with pytest.raises(RuntimeError):
        loader.DevicePrefetcher([], device="cpu")
