"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/graphtrans_b200.h declares (and nothing in the ctypes table is stale), and the product
package never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "graphtrans_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from graphtrans_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", ROOT, "-j8"], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_the_hot_path():
    syms = header_symbols()
    for must in ("gt_csr_build", "gt_aggregate_fwd", "gt_aggregate_bwd", "gt_batch_plan", "gt_pad_batch_fwd",
                 "gt_mha_fwd", "gt_mha_bwd", "gt_gemm", "gt_pna_reduce_fwd", "gt_pna_reduce_bwd", "gt_layernorm_fwd"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header(lib):
    from graphtrans_b200 import _lib
    declared = set(header_symbols()) - {"gt_version", "gt_last_error"}
    assert set(_lib.SIGNATURES) == declared
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\bint\s+" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        n_params = len([p for p in m.group(1).split(",") if p.strip()])
        assert n_params == len(argtypes) + 0, f"{name}: header has {n_params} params, ctypes table {len(argtypes)}"


def test_version_and_error_string(lib):
    lib.gt_version.restype = ctypes.c_int
    lib.gt_last_error.restype = ctypes.c_char_p
    assert lib.gt_version() >= 100
    assert isinstance(lib.gt_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu(lib):
    """shape validation happens before any CUDA call: a bad call returns < 0 and sets the message"""
    lib.gt_last_error.restype = ctypes.c_char_p
    lib.gt_dropout.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float,
                               ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
    assert lib.gt_dropout(0, None, 3, None, 0.5, None, 0, None) < 0
    assert b"multiple of 4" in lib.gt_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "graphtrans_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", txt, flags=re.S).replace("# oracle", ""), f


def test_no_fallback_without_cuda():
    """the product path fails loudly on CPU tensors instead of falling back"""
    import torch
    from graphtrans_b200 import ops
    with pytest.raises(RuntimeError):
        ops.GraphPlan(torch.zeros(2, 0, dtype=torch.long), torch.zeros(3, dtype=torch.long), 1)
