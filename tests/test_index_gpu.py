"""Bit-exact integer work on the GPU: CSR build / degrees, batch plan, pad_batch layout and mask
(reference modules/utils.py:5-29, modules/conv.py:57) against numpy / the oracle's closed form,
which tests/test_oracle_golden.py and tests/test_pad_cpu.py pin to the reference's Python loop."""
import numpy as np
import pytest
import torch

from graphtrans_b200 import ops, synth
from graphtrans_b200.modules.utils import pad_batch, unpad_batch
from oracle import graphtrans_oracle as O

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("gen,B", [("nci1", 32), ("mol", 64), ("code2", 12), ("syn", 16)])
def test_csr_build_bit_exact(gen, B):
    batch = getattr(synth, "gen_" + gen)(B, seed=3)
    ei = batch.edge_index
    N = batch.batch.numel()
    plan = ops.GraphPlan(ei.cuda(), batch.batch.cuda(), B)
    src, dst = ei[0].numpy(), ei[1].numpy()
    E = src.size
    for key, nbr, rp, nb, eid in (("dst", src, plan.rowptr_dst, plan.src_by_dst, plan.eid_by_dst),
                                  ("src", dst, plan.rowptr_src, plan.dst_by_src, plan.eid_by_src)):
        k = dst if key == "dst" else src
        order = np.argsort(k, kind="stable")
        cnt = np.bincount(k, minlength=N)
        ref_rp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        assert np.array_equal(_np(rp), ref_rp)
        assert np.array_equal(_np(eid)[:E], order.astype(np.int32))
        assert np.array_equal(_np(nb)[:E], nbr[order].astype(np.int32))


@pytest.mark.parametrize("L", [1, 3, 7, 1000])
@pytest.mark.parametrize("cls", [True, False])
def test_batch_plan_bit_exact(L, cls):
    batch = synth.gen_code2(9, seed=4, nmin=1, nmax=40, mu=2.5, sigma=0.8)
    bidx = batch.batch
    N, B = bidx.numel(), 9
    plan = ops.GraphPlan(batch.edge_index.cuda(), bidx.cuda(), B, L, cls=cls)
    n, off, k, S = O.pad_plan(bidx, L)
    c = 1 if cls else 0
    assert plan.S == S
    assert np.array_equal(_np(plan.node_off), np.concatenate([off.numpy(), [N]]))
    assert np.array_equal(_np(plan.kept), k.numpy())
    tok_off = np.concatenate([[0], np.cumsum(k.numpy() + c)])
    assert np.array_equal(_np(plan.tok_off), tok_off)
    t2n = np.full(N + B, -2, np.int64)
    n2t = np.full(N, -1, np.int64)
    pooled = np.zeros(B, np.int64)
    for g in range(B):
        kept_nodes = np.arange(off[g] + n[g] - k[g], off[g] + n[g])
        rows = tok_off[g] + np.arange(k[g])
        t2n[rows] = kept_nodes
        n2t[kept_nodes] = rows
        if cls:
            t2n[tok_off[g] + k[g]] = -1
            pooled[g] = tok_off[g] + k[g]
        else:
            pooled[g] = tok_off[g] + k[g] - 1
    assert np.array_equal(_np(plan.tok2node), t2n)
    assert np.array_equal(_np(plan.node2tok), n2t)
    assert np.array_equal(_np(plan.cls_rows), pooled)
    tg = _np(plan.tok_graph)
    for g in range(B):
        assert (tg[tok_off[g]:tok_off[g + 1]] == g).all()
    assert (tg[tok_off[B]:] == -1).all()


@pytest.mark.parametrize("L", [1, 3, 7, 1000])
@pytest.mark.parametrize("d", [8, 36, 300])
def test_pad_batch_bit_exact(L, d):
    batch = synth.gen_code2(7, seed=5, nmin=1, nmax=30, mu=2.3, sigma=0.8)
    bidx = batch.batch
    h = torch.randn(bidx.numel(), d)
    ref_p, ref_m = O.pad_batch(h, bidx, L)
    padded, mask, num_nodes, masks, S = pad_batch(h.cuda(), bidx.cuda(), L, get_mask=True)
    assert padded.shape == ref_p.shape and mask.dtype == torch.bool
    assert torch.equal(padded.cpu(), ref_p)          # pure copies: bit-exact
    assert torch.equal(mask.cpu(), ref_m)
    assert S == ref_p.shape[0] and len(num_nodes) == 7 and len(masks) == 7
    assert [int(v) for v in num_nodes] == torch.bincount(bidx).tolist()
    assert torch.equal(masks[2].cpu(), bidx.eq(2))
    p2, m2 = pad_batch(h.cuda(), bidx.cuda(), L)
    assert torch.equal(p2.cpu(), ref_p) and torch.equal(m2.cpu(), ref_m)
    # backward = inverse gather: only kept nodes receive gradient, exactly
    hh = h.cuda().requires_grad_(True)
    pp, _ = pad_batch(hh, bidx.cuda(), L)
    w = torch.randn_like(pp)
    (pp * w).sum().backward()
    hr = h.clone().requires_grad_(True)
    rp, _ = O.pad_batch(hr, bidx, L)
    (rp * w.cpu()).sum().backward()
    assert torch.equal(hh.grad.cpu(), hr.grad)
    # unpad_batch restores kept rows, truncated rows keep prev (reference modules/utils.py:32-53)
    prev = torch.randn_like(h).cuda()
    un = unpad_batch(padded, prev, num_nodes, masks, S)
    n, off, k, _ = O.pad_plan(bidx, L)
    for g in range(7):
        lo, hi = int(off[g]), int(off[g] + n[g])
        keep_lo = hi - int(k[g])
        assert torch.equal(un[keep_lo:hi].cpu(), h[keep_lo:hi])
        assert torch.equal(un[lo:keep_lo], prev[lo:keep_lo])


def test_empty_edge_list_and_single_node_graphs():
    bidx = torch.tensor([0, 1, 1, 2])
    plan = ops.GraphPlan(torch.zeros(2, 0, dtype=torch.long).cuda(), bidx.cuda(), 3, 1000)
    assert _np(plan.rowptr_dst).tolist() == [0, 0, 0, 0, 0]
    assert _np(plan.tok_off).tolist() == [0, 2, 5, 7]
    assert plan.S == 2


@pytest.mark.parametrize("B,ntypes", [(64, 60), (7, 3), (300, 1000)])
def test_edges_by_type_is_a_sorted_permutation(B, ntypes):
    """gt_edges_by_type: counting sort of the edges by combined edge type - run boundaries equal the histogram scan,
    types are sorted, and the (src, dst, type) multiset is preserved (order inside a run is unspecified)"""
    batch = synth.gen_mol(B, seed=5)
    ei = batch.edge_index.cuda()
    E = ei.shape[1]
    g = torch.Generator().manual_seed(1)
    et = torch.randint(0, ntypes, (E,), generator=g, dtype=torch.int32)
    plan = ops.GraphPlan(ei, batch.batch.cuda(), B)
    src_t, dst_t, type_t, type_ptr = plan.edges_by_type(ei, et.cuda(), ntypes)
    cnt = np.bincount(et.numpy(), minlength=ntypes)
    assert np.array_equal(_np(type_ptr), np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32))
    ty = _np(type_t)[:E]
    assert np.all(np.diff(ty) >= 0)
    got = np.stack([ty, _np(src_t)[:E], _np(dst_t)[:E]], 1)
    ref = np.stack([et.numpy(), ei[0].cpu().numpy().astype(np.int32), ei[1].cpu().numpy().astype(np.int32)], 1)
    assert np.array_equal(got[np.lexsort(got.T[::-1])], ref[np.lexsort(ref.T[::-1])])
