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
