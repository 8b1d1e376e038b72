"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo processes run the bucketed
gradient allreduce of graphtrans_b200/ddp.py on a small torch model (the CUDA kernels are not involved)
and must end with identical, averaged gradients; shard_range partitions graphs contiguously."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphtrans_b200.ddp import GradBuckets, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(7, 13), torch.nn.ReLU(), torch.nn.Linear(13, 5), torch.nn.ReLU(),
                               torch.nn.Linear(5, 3))


def _worker(rank, world, port, n_buckets, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _model()
    buckets = GradBuckets(model, n_buckets=n_buckets)
    torch.manual_seed(100 + rank)
    for step in range(2):                     # second step checks zero_grad()/re-arming
        x = torch.randn(11, 7)
        buckets.zero_grad()
        model(x).square().sum().backward()
        buckets.finish()
    flat = buckets.flat.clone()
    grads = torch.cat([p.grad.flatten() for p in buckets.params])      # the (128-byte aligned) views, without the padding
    aligned = all((p.grad.data_ptr() - buckets.flat.data_ptr()) % 128 == 0 for p in buckets.params)
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save(dict(flat=grads, aligned=aligned, same=all(torch.equal(g, gathered[0]) for g in gathered),
                        n_buckets=len(buckets.buckets), views=all(p.grad.data_ptr() >= buckets.flat.data_ptr()
                                                                  for p in buckets.params)), out)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_buckets", [1, 4])
def test_bucketed_allreduce_gloo_world2(tmp_path, n_buckets):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_buckets, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["same"] and res["views"] and res["aligned"] and res["n_buckets"] <= n_buckets
    # expected: mean over ranks of the per-rank gradient of the SECOND step
    exp = []
    for rank in range(2):
        model = _model()
        torch.manual_seed(100 + rank)
        torch.randn(11, 7)
        x = torch.randn(11, 7)
        model(x).square().sum().backward()
        exp.append(torch.cat([p.grad.flatten() for p in model.parameters()]))
    assert torch.allclose(res["flat"], (exp[0] + exp[1]) / 2, rtol=1e-6, atol=1e-7)


def test_shard_range_partitions_contiguously():
    for n, w in [(4096, 8), (13, 4), (2, 2), (7, 1)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
