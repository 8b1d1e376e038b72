"""Pins the oracle's closed-form pad_batch (integer/index work) bit-exactly against outputs of the
reference's own Python loop (tests/golden/_pad_batch_ref.pt, written by oracle/gen_golden.py from
reference modules/utils.py:5-53): ragged batches, 1-node graphs, truncation L in {1,3,7,1000}."""
import os

import pytest
import torch

from oracle import graphtrans_oracle as O
from tests.helpers import GOLDEN_DIR

CASES = torch.load(os.path.join(GOLDEN_DIR, "_pad_batch_ref.pt"), weights_only=False)


def test_fixture_covers_truncation_and_ragged():
    assert len(CASES) == 40
    assert any(max(c["sizes"]) > c["L"] for c in CASES) and any(1 in c["sizes"] for c in CASES)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_pad_batch_bit_exact(i):
    c = CASES[i]
    padded, mask = O.pad_batch(c["h"], c["batch"], c["L"])
    assert torch.equal(padded, c["padded"])
    assert torch.equal(mask, c["mask"])
    n, off, k, S = O.pad_plan(c["batch"], c["L"])
    assert S == c["max_num_nodes"] and n.tolist() == c["num_nodes"]
