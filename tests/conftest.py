import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the warp-per-graph pooled-query kernels (csrc/attn_cls.cu) are selected by batch shape (>= 592 short graphs); the tests
# use handfuls of graphs, so they force them for every eligible layout (d = 256, bf16) - the warp-per-(graph, head)
# kernels stay covered by the fp32 and d = 128 cases
os.environ.setdefault("GT_CLS_WIDE", "2")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
