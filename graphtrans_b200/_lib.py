"""ctypes binding of the C ABI declared in include/graphtrans_b200.h.

There is NO fallback: if the shared library is missing or a call fails the product path raises
(`RuntimeError`), it never routes through PyTorch eager ops or the CPU oracle.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgraphtrans_b200.so")

GT_F32, GT_BF16 = 0, 1
EDGE_NONE, EDGE_LINEAR, EDGE_TABLE = 0, 1, 2
CONV_GCN, CONV_GIN = 0, 1
EPI_RELU, EPI_ACCUM, EPI_OUT_F32, EPI_RESID_F32 = 1, 2, 4, 8

P, I, L, F = c_void_p, c_int, c_int64, c_float
I32 = c_int32
U64 = c_uint64

# name -> argtypes; every export returns int (0 ok). Kept in the order of include/graphtrans_b200.h
SIGNATURES = {
    "gt_csr_build": [P, L, L, P, P, P, P, P, P, P, P],
    "gt_edge_type": [P, L, I32, P, P, P],
    "gt_rng_advance": [P, P],
    "gt_dropout": [I, P, L, P, F, P, U64, P],
    "gt_batch_plan": [P, L, L, L, I32, P, P, P, P, P, P, P, P, P, P],
    "gt_embed_sum_fwd": [I, P, L, I32, I32, I32, P, P, P, P, P],
    "gt_embed_sum_bwd": [I, P, L, I32, I32, I32, P, P, P, P, P],
    "gt_onehot": [L, I32, P, P, P, P, I32, P, P],
    "gt_embed_unpack": [P, I32, I32, I32, I32, P, P, P, P],
    "gt_aggregate_fwd": [I, I, P, P, L, I32, I32, P, P, P, P, I, P, I32, P, P, P, P, P, P, P, P, P],
    "gt_aggregate_bwd": [I, I, P, P, P, L, I32, I32, P, P, P, P, I, P, I32, P, P, P, P, I32, P, P, P, P, P, P, P, P, P, P, P],
    "gt_edge_slots": [P, P, P, P, L, L, P, P, I32, P, P, P, P],
    "gt_edges_by_type": [P, P, L, I32, P, P, P, P, P, P],
    "gt_aggregate_table_grad": [I, I, P, P, L, I32, I32, P, L, P, P, P, P, I32, P, P],
    "gt_segment_sum": [I, P, P, L, I32, P, P],
    "gt_segment_sum_sorted": [I, P, P, L, I32, P, P, P],
    "gt_add_graph_vec": [I, P, P, P, L, I32, P, P],
    "gt_colstats": [I, P, L, I32, P, P, P],
    "gt_bn_finalize": [P, L, I32, I32, P, P, P, P, P, F, F, I, P, P],
    "gt_bn_apply_fwd": [I, P, L, I32, I32, P, I, P, P, P, P, F, P, U64, P],
    "gt_bn_norm_fwd": [I, P, L, I32, I32, P, P, P, P, P, P, F, F, I, I, P, P, P, P, P, F, P, U64, P, P],
    "gt_bn_bwd_reduce": [I, P, P, L, I32, I32, P, I, P, F, P, U64, P, P],
    "gt_bn_bwd_apply": [I, P, P, L, I32, I32, P, P, I, I, P, P, P, P, F, P, U64, P, P],
    "gt_gemm": [I, P, I, L, P, I, L, P, L, L, L, L, L, P, P, L, I, F, P, U64, I, P],
    "gt_gemm_stats": [I, P, I, L, P, I, L, P, L, L, L, L, L, P, P, L, I, F, P, U64, I, P, P, P],
    "gt_split3": [P, L, L, L, P, L, P],
    "gt_relu_bwd": [I, P, P, L, P, F, P],
    "gt_relu_bwd_colsum": [I, P, P, L, L, L, P, F, P, P],
    "gt_colsum": [I, P, L, L, L, P, P],
    "gt_cast_multi": [P, I32, L, P],
    "gt_add_blocks": [P, I32, I32, P, P, P, P, P, P, P],
    "gt_cast_pad": [I, P, L, L, L, I, P, L, L, L, P],
    "gt_layernorm_fwd": [I, P, P, P, P, L, I32, P, P, F, P, P, P, F, P, U64, P],
    "gt_layernorm_bwd": [I, P, P, P, P, L, I32, P, P, P, P, P, P, F, P, U64, P],
    "gt_gather_rows": [I, P, P, P, L, I32, P, P],
    "gt_scatter_rows": [I, P, P, L, I32, P, P, P],
    "gt_pad_batch_fwd": [I, P, P, L, L, I32, P, P, P],
    "gt_pad_batch_bwd": [I, P, P, P, L, L, L, I32, P, P],
    "gt_mha_meta": [P, P, L, L, P, P, P],
    "gt_mha_fwd": [I, P, P, P, P, P, P, L, L, I32, I32, F, P, P, F, P, U64, I, P],
    "gt_mha_bwd": [I, P, P, P, P, P, P, P, P, P, L, L, I32, I32, F, P, P, F, P, U64, I, P],
    "gt_mha_local_tiles": [P, L, L, P, P, P],
    "gt_mha_local_fwd": [I, P, P, P, P, L, L, I32, I32, F, P, P, F, P, U64, P],
    "gt_mha_local_bwd": [I, P, P, P, P, P, P, L, L, I32, I32, F, P, F, P, U64, P],
    "gt_bce_masked_fwd": [P, P, L, I32, L, L, P, P, P],
    "gt_bce_masked_bwd": [P, P, L, I32, L, L, P, P, P, L, I32, P],
    "gt_ce_fwd": [P, P, L, L, I32, L, P, P, P, P],
    "gt_ce_bwd": [P, P, L, L, I32, L, P, P, P, L, I32, P],
    "gt_mha_cls_fwd": [I, P, P, P, P, L, L, I32, I32, F, P, P, F, P, U64, P],
    "gt_mha_cls_bwd": [I, P, P, P, P, P, P, P, L, L, I32, I32, F, P, P, F, P, U64, P],
    "gt_adamw_multi": [P, I32, L, P, P, P, P, P, P, P],
    "gt_sumsq": [P, L, P, P, I32, P],
    "gt_segment_pool_fwd": [I, I, P, P, L, I32, P, P, P],
    "gt_segment_pool_bwd": [I, I, P, P, P, P, L, I32, P, P],
    "gt_argmax_rows": [P, L, I32, L, P, P],
    "gt_pna_reduce_fwd": [I, P, P, P, L, I32, I32, I32, P, P, F, P, I32, P, P, P],
    "gt_pna_reduce_bwd": [I, P, P, P, L, I32, I32, I32, I32, P, P, F, P, P, P, P, P, P],
}

# kernels launched per export (everything not listed launches exactly one); cudaMemsetAsync nodes
# are not counted
KERNELS_PER_CALL = {"gt_csr_build": 4, "gt_batch_plan": 3, "gt_mha_bwd": 3, "gt_edges_by_type": 3, "gt_adamw_multi": 2}

_lib = None
launch_count = 0   # C-ABI calls issued
kernel_count = 0   # kernels those calls launched (bench.py reports it as gpu_launches)
_profile = None    # when a list: (name, start_event, end_event, info) per call (bench.py roofline pass)
_stamps = None     # when a dict: device timestamp after every call (tools/graph_trace.py; survives graph capture)


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"graphtrans_b200: {LIB_PATH} is missing - build it with `make` (or "
            "`python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.gt_version.restype = c_int
    lib.gt_last_error.restype = c_char_p
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


def ptr(t):
    return None if t is None else t.data_ptr()


def dt_of(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return GT_F32
    if t.dtype == torch.bfloat16:
        return GT_BF16
    raise TypeError(f"graphtrans_b200: unsupported activation dtype {t.dtype}")


def stream():
    return torch.cuda.current_stream().cuda_stream   # follows torch.cuda.stream(...) contexts (side streams, capture)


def call(name, *args):
    """Invoke an export on torch's current CUDA stream; non-zero return -> RuntimeError."""
    global launch_count, kernel_count
    lib = load()
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream())
        e1.record()
        _profile.append((name, e0, e1, args))
    else:
        rc = getattr(lib, name)(*args, stream())
    if _stamps is not None:
        stamp(name, args)
    launch_count += 1
    kernel_count += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise RuntimeError(f"{name} failed (rc={rc}): {lib.gt_last_error().decode()}")


def start_stamps(capacity=8192, device="cuda"):
    """record a %globaltimer stamp on the launching stream after every C-ABI call (and wherever stamp() is called)"""
    global _stamps
    _stamps = {"buf": torch.zeros(capacity, dtype=torch.int64, device=device), "rec": []}


def stamp(name, args=()):
    st = _stamps
    idx = len(st["rec"])
    load().gtdbg_stamp(ctypes.c_void_p(st["buf"].data_ptr()), idx, ctypes.c_void_p(stream()))
    st["rec"].append((name, stream(), args))


def stop_stamps():
    """-> (records [(name, stream, args)], int64 tensor of ns stamps); the buffer keeps being rewritten by replays"""
    global _stamps
    st, _stamps = _stamps, None
    return st["rec"], st["buf"]


def start_profile():
    global _profile
    _profile = []


def stop_profile():
    """-> list of (export name, milliseconds, call args); synchronises the device"""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    return [(n, e0.elapsed_time(e1), a) for n, e0, e1, a in rec]


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("graphtrans_b200 runs on CUDA tensors only (no CPU fallback)")
