"""Host-side logic that needs no GPU: token buckets of the tile-local attention vs the CUDA-graph signature, gradient
arena alignment, the C-ABI kernel-count table."""
import types

import torch

from graphtrans_b200 import _lib, ops
from graphtrans_b200.ddp import GradBuckets
from graphtrans_b200.graphed import _signature


def test_token_bucket_is_a_function_of_the_signature_field():
    """two batches with the same graph signature must select the same attention path / launch bounds: the signature
    carries ceil((max_nodes + 1) / 32) and ops.token_bucket depends on max_nodes only through that value"""
    seen = {}
    for mn in range(0, 300):
        b = types.SimpleNamespace(max_nodes=mn)
        sig = dict(_signature(b))["max_nodes"]
        for L in (1000, 50):
            for cls in (True, False):
                key = (sig, L, cls)
                bucket = ops.token_bucket(mn, L, cls)
                if L >= mn:                       # truncation can only merge buckets, never split a signature class
                    assert seen.setdefault(key, bucket) == bucket, (mn, L, cls)
                if bucket is not None:
                    assert min(mn, L) + (1 if cls else 0) <= bucket <= 128
    assert ops.token_bucket(None) is None and ops.token_bucket(127) == 128 and ops.token_bucket(128) is None
    assert dict(_signature(types.SimpleNamespace(max_nodes=None)))["max_nodes"] is None


def test_gradient_arena_views_are_128_byte_aligned():
    m = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Linear(5, 1), torch.nn.Linear(1, 7))   # odd sizes, 1-element bias
    gb = GradBuckets(m, n_buckets=2, overlap=False)
    base = gb.flat.data_ptr()
    for p, off in zip(gb.params, gb.offsets):
        assert p.grad.data_ptr() == base + 4 * off and (4 * off) % 128 == 0
        assert p.grad.shape == p.shape
    assert gb.flat.numel() >= sum(p.numel() for p in m.parameters())
    gb.flat.fill_(1.0)
    gb.zero_grad()
    assert float(gb.flat.abs().sum()) == 0.0


def test_every_signature_names_an_export_of_the_header():
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(_lib.__file__), "..", "include", "graphtrans_b200.h")).read()
    declared = set(re.findall(r"\b(gt_[a-z0-9_]+)\s*\(", hdr))
    for name in _lib.SIGNATURES:
        assert name in declared, name
    for name in _lib.KERNELS_PER_CALL:
        assert name in _lib.SIGNATURES


def test_bf16_relu_threshold_identity():
    """The packed-mask GIN adjoint (csrc/aggregate.cu k_agg_bwd3p) replaces the fp32 predicate  x + e > 0  (x a bf16
    activation, e an fp32 edge-table entry; reference modules/conv.py:32 `F.relu(x_j + edge_attr)`) by the bf16 compare
    x > th with th = round-DOWN-to-bf16(-e) (k_edge_table_thresholds: truncate, +1 ulp of magnitude when negative and
    inexact).  Checked here for EVERY finite bf16 x against random and adversarial e (ties, one fp32 ulp either side of a
    bf16 value, tiny / huge magnitudes): the two predicates agree exactly."""
    import numpy as np
    import torch

    def thresholds(e):                      # bit-level restatement of k_edge_table_thresholds
        v = (-e).astype(np.float32).view(np.uint32)
        b = (v >> 16).astype(np.uint32)
        b = np.where(((v & 0xFFFF) != 0) & ((v >> 31) != 0), b + 1, b)
        return (b.astype(np.uint32) << 16).view(np.float32)      # the bf16 value, widened

    bits = np.arange(1 << 16, dtype=np.uint32)
    xs = (bits << 16).view(np.float32)
    xs = xs[np.isfinite(xs)]
    rng = np.random.default_rng(0)
    es = [rng.standard_normal(64).astype(np.float32), (rng.standard_normal(16) * 1e-30).astype(np.float32),
          (rng.standard_normal(16) * 1e30).astype(np.float32), np.array([0.0, -0.0, 1.0, -1.0], np.float32)]
    near = xs[rng.integers(0, xs.size, 256)]
    near = near[np.abs(near) < 1e30]
    es += [-near, -np.nextafter(near, np.float32(np.inf)), -np.nextafter(near, np.float32(-np.inf))]
    e = np.concatenate(es).astype(np.float32)
    th = thresholds(e)
    x = torch.from_numpy(xs)[:, None]
    ref = (x + torch.from_numpy(e)[None, :]) > 0                 # fp32 add + compare, as the fp32-mask kernels do
    got = x > torch.from_numpy(th)[None, :]
    assert torch.equal(ref, got)
