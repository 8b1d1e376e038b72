"""Per-shape timing of the tcgen05 GEMM (CUDA events, 50 back-to-back launches, warm L2) for the
contraction shapes of the molpcba / code2 steps.  python tools/gemm_bench.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphtrans_b200._lib import call, ptr, EPI_ACCUM, EPI_OUT_F32, EPI_RELU

ONLY = os.environ.get("GEMM_ONLY")


def bench(name, M, N, K, a_mn, b_mn, flags=0, reps=50):
    if ONLY and ONLY != name:
        return
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
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): run()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    fl = 2.0 * M * N * K
    byt = (M * K + N * K) * 2 + M * N * (4 if out_f32 else 2)
    print(f"{name:34s} M={M:6d} N={N:5d} K={K:6d} amn={a_mn} bmn={b_mn} {us:8.2f} us  {fl/us/1e6:8.1f} TF/s  {byt/us/1e3:8.1f} GB/s(min traffic)")

if __name__ == "__main__":
    Nn, T, B = 13510, 14022, 512
    for name, M, N, K in [("gin mlp.0 fwd", Nn, 600, 300), ("gin mlp.3 fwd", Nn, 300, 600), ("in_proj fwd", T, 384, 128),
                          ("out_proj fwd", T, 128, 128), ("ffn1 fwd", T, 512, 128), ("ffn2 fwd", T, 128, 512),
                          ("g2t fwd", Nn, 128, 300), ("vn mlp fwd", B, 600, 300), ("code2 head", 128, 5002, 256),
                          ("code2 qkv", 15970, 768, 256)]:
        bench(name, M, N, K, 0, 0, EPI_RELU)
    for name, M, N, K in [("gin mlp.0 dX", Nn, 300, 600), ("gin mlp.3 dX", Nn, 600, 300), ("in_proj dX", T, 128, 384),
                          ("ffn1 dX", T, 128, 512), ("ffn2 dX", T, 512, 128)]:
        bench(name, M, N, K, 0, 1)
    for name, M, N, K in [("gin mlp.0 dW", 600, 300, Nn), ("gin mlp.3 dW", 300, 600, Nn), ("in_proj dW", 384, 128, T),
                          ("ffn1 dW", 512, 128, T), ("ffn2 dW", 128, 512, T), ("out_proj dW", 128, 128, T),
                          ("code2 head dW", 5002, 256, 128)]:
        bench(name, M, N, K, 1, 1, EPI_ACCUM | EPI_OUT_F32)
