"""Does alternating the 226 KB-shared-memory tcgen05 GEMM with small elementwise kernels cost more than the sum of the
two kernels timed alone (shared-memory carve-out reconfiguration / cold descriptors)?  CUDA-graph replays of
50 x [A], 50 x [B] and 50 x [A, B].  python tools/transition_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200._lib import EPI_ACCUM, EPI_OUT_F32, call, ptr  # noqa: E402


def graph_time(fns, reps=50):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns:
                f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def mk_gemm(M, N, K, a_mn, b_mn, flags):
    lda = ((M if a_mn else K) + 7) // 8 * 8
    ldb = ((N if b_mn else K) + 7) // 8 * 8
    A = torch.randn((K if a_mn else M), lda, device="cuda").bfloat16()
    B = torch.randn((K if b_mn else N), ldb, device="cuda").bfloat16()
    ldc = (N + 7) // 8 * 8
    C = torch.zeros(M, ldc, device="cuda", dtype=torch.float32 if flags & EPI_OUT_F32 else torch.bfloat16)
    return lambda: call("gt_gemm", 1, ptr(A), a_mn, lda, ptr(B), b_mn, ldb, ptr(C), ldc, M, N, K, ldc, None, None, 0, flags,
                        0.0, None, 0, 2)


T = 13745
x = torch.randn(T, 512, device="cuda").bfloat16()
y = torch.empty_like(x)
g2 = torch.randn(T, 512, device="cuda").bfloat16()
small = torch.randn(512, 128, device="cuda")
small_o = torch.empty(512, 128, device="cuda").bfloat16()


def relu_bwd():
    call("gt_relu_bwd", 1, ptr(g2), ptr(x), x.numel(), ptr(y), 1.0)


def tiny():
    call("gt_cast_pad", 0, ptr(small), 512, 128, 128, 1, ptr(small_o), 512, 128, 128)


cases = {
    "dW 128x128 K=13745": mk_gemm(128, 128, T, 1, 1, EPI_ACCUM | EPI_OUT_F32),
    "dW 600x300 K=13233": mk_gemm(600, 300, 13233, 1, 1, EPI_ACCUM | EPI_OUT_F32),
    "fwd 13745x512 K=128": mk_gemm(T, 512, 128, 0, 0, 0),
    "dX 13745x128 K=512": mk_gemm(T, 128, 512, 0, 1, 0),
}
t_relu, t_tiny = graph_time([relu_bwd]), graph_time([tiny])
print(f"relu_bwd [13745x512] alone {t_relu:.2f} us; tiny cast alone {t_tiny:.2f} us")
for name, fn in cases.items():
    a = graph_time([fn])
    ab = graph_time([fn, relu_bwd])
    at = graph_time([fn, tiny])
    print(f"{name:24s} alone {a:6.2f}  +relu_bwd pair {ab:6.2f} (sum {a + t_relu:6.2f})  +tiny pair {at:6.2f} (sum {a + t_tiny:6.2f})")
fns = list(cases.values())
print(f"4 different GEMMs in sequence: {graph_time(fns):.2f} us per round (sum alone {sum(graph_time([f]) for f in fns):.2f})")
