"""Direct NCCL binding (ctypes on the libnccl that torch already loaded) for the gradient allreduce.

`ncclAllReduce` is issued on an explicit CUDA stream, so it can be captured INSIDE the CUDA graph of the training step
as a forked branch (one node per gradient bucket, launched as soon as the backward has produced the bucket's last
gradient) and overlaps the rest of the backward on replay.  torch.distributed is still what bootstraps the ranks
(the 128-byte NCCL unique id travels through the default process group); the data path is NCCL over NVLink / NVSwitch
only.  The reference is single-GPU: this is the one collective of the data-parallel design."""
from __future__ import annotations

import ctypes
import glob
import os

import torch
import torch.distributed as dist

NCCL_FLOAT32, NCCL_BFLOAT16 = 7, 9
NCCL_SUM, NCCL_AVG = 0, 4

_lib = None


class _UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * 128)]


def load():
    global _lib
    if _lib is not None:
        return _lib
    cands = ["libnccl.so.2"]
    base = os.path.dirname(os.path.dirname(torch.__file__))
    cands += sorted(glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*")))
    err = None
    for c in cands:
        try:
            lib = ctypes.CDLL(c)
            break
        except OSError as e:      # noqa: PERF203
            err = e
    else:
        raise RuntimeError(f"graphtrans_b200.nccl: libnccl not found ({err})")
    lib.ncclGetErrorString.restype = ctypes.c_char_p
    lib.ncclGetErrorString.argtypes = [ctypes.c_int]
    lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(_UniqueId)]
    lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
    lib.ncclAllReduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p]
    lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    lib.ncclGetVersion.argtypes = [ctypes.POINTER(ctypes.c_int)]
    for fn in ("ncclGetUniqueId", "ncclCommInitRank", "ncclAllReduce", "ncclCommDestroy", "ncclGetVersion"):
        getattr(lib, fn).restype = ctypes.c_int
    _lib = lib
    return lib


def _check(lib, rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib.ncclGetErrorString(rc).decode()}")


def version() -> int:
    lib = load()
    v = ctypes.c_int(0)
    _check(lib, lib.ncclGetVersion(ctypes.byref(v)), "ncclGetVersion")
    return v.value


class Communicator:
    """one NCCL communicator over the ranks of `group` (default: the world), created on the current CUDA device"""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("graphtrans_b200.nccl.Communicator needs an initialised torch.distributed process group")
        lib = load()
        self.lib = lib
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        uid = _UniqueId()
        if self.rank == 0:
            _check(lib, lib.ncclGetUniqueId(ctypes.byref(uid)), "ncclGetUniqueId")
        box = [bytes(uid.internal)]
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(box, src=src, group=group)
        ctypes.memmove(ctypes.byref(uid), box[0], 128)
        self.comm = ctypes.c_void_p()
        _check(lib, lib.ncclCommInitRank(ctypes.byref(self.comm), self.world, uid, self.rank), "ncclCommInitRank")
        self.device = torch.cuda.current_device()

    def all_reduce(self, t: torch.Tensor, avg: bool = True, stream=None):
        """in-place sum (or average) of a contiguous fp32 / bf16 CUDA tensor across the ranks, asynchronous on `stream`
        (default: torch's current stream); capturable"""
        if not t.is_cuda or not t.is_contiguous():
            raise RuntimeError("nccl.all_reduce: needs a contiguous CUDA tensor")
        dt = {torch.float32: NCCL_FLOAT32, torch.bfloat16: NCCL_BFLOAT16}.get(t.dtype)
        if dt is None:
            raise TypeError(f"nccl.all_reduce: unsupported dtype {t.dtype}")
        st = (stream or torch.cuda.current_stream()).cuda_stream
        _check(self.lib, self.lib.ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), dt, NCCL_AVG if avg else NCCL_SUM,
                                                self.comm, ctypes.c_void_p(st)), "ncclAllReduce")

    def destroy(self):
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = ctypes.c_void_p()
