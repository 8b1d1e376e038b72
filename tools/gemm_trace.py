"""Phase breakdown of one CTA of the tcgen05 GEMM (clock64 stamps, GT_GEMM_TRACE=1):
0 entry | 1 setup done (barriers, TMEM alloc) | 2 first TMA issued | 3 first stage landed | 4 tile-0 MMAs committed |
5 epilogue sees tile 0 | 6 epilogue tile 0 issued its stores | 7 epilogue loop done | 8 stores drained | 9 exit."""
import ctypes
import os
import sys

os.environ["GT_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200 import _lib  # noqa: E402
from graphtrans_b200._lib import EPI_ACCUM, EPI_OUT_F32, EPI_RELU, call, ptr  # noqa: E402


def trace(name, M, N, K, a_mn, b_mn, flags=0):
    lda = ((M if a_mn else K) + 7) // 8 * 8
    ldb = ((N if b_mn else K) + 7) // 8 * 8
    A = torch.randn((K if a_mn else M), lda, device="cuda").bfloat16()
    B = torch.randn((K if b_mn else N), ldb, device="cuda").bfloat16()
    out_f32 = bool(flags & EPI_OUT_F32)
    ldc = (N + 7) // 8 * 8
    C = torch.zeros(M, ldc, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    bias = torch.randn(N, device="cuda") if not (flags & EPI_ACCUM) else None

    def run():
        call("gt_gemm", 1, ptr(A), a_mn, lda, ptr(B), b_mn, ldb, ptr(C), ldc, M, N, K, ldc, ptr(bias), None, 0, flags, 0.0, None, 0, 2)

    for _ in range(4):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    lib = _lib.load()
    assert lib.gtdbg_gemm_trace_read(buf) == 0
    t = [buf[i] for i in range(10)]
    ns = [(x - t[0]) / 1.965 for x in t]
    print(f"{name:18s} M={M:6d} N={N:5d} K={K:6d} event {e0.elapsed_time(e1) * 1e3:6.1f} us | stamps (ns from entry): "
          + " ".join(f"{i}:{v:7.0f}" for i, v in enumerate(ns)))


if __name__ == "__main__":
    Nn, T, B = 13510, 14022, 512
    trace("vn mlp fwd", B, 600, 300, 0, 0, EPI_RELU)
    trace("1 tile k=64", 128, 128, 64, 0, 0, EPI_RELU)
    trace("1 tile k=512", 128, 256, 512, 0, 0, EPI_RELU)
    trace("gin mlp.0 fwd", Nn, 600, 300, 0, 0, EPI_RELU)
    trace("gin mlp.3 fwd", Nn, 300, 600, 0, 0, EPI_RELU)
    trace("ffn1 fwd", T, 512, 128, 0, 0, EPI_RELU)
    trace("out_proj fwd", T, 128, 128, 0, 0, EPI_RELU)
    trace("gin mlp.0 dX", Nn, 300, 600, 0, 1)
    trace("gin mlp.0 dW", 600, 300, Nn, 1, 1, EPI_ACCUM | EPI_OUT_F32)
    trace("out_proj dW", 128, 128, T, 1, 1, EPI_ACCUM | EPI_OUT_F32)
