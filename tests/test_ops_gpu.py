"""Operator-level GPU tests through the C ABI against plain fp32/fp64 torch references of the same
op (dense contraction layouts and epilogues, packed masked attention fwd/bwd, dropout masks)."""
import math

import pytest
import torch

from graphtrans_b200 import _lib, ops
from graphtrans_b200._lib import EPI_ACCUM, EPI_OUT_F32, EPI_RELU, call, dt_of, ptr
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(70, 36, 300), (129, 600, 304), (513, 128, 64), (5, 5002, 128)])
def test_gemm_layouts(dtype, a_mn, b_mn, M, N, K):
    torch.manual_seed(0)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(dtype)
    Bm = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(dtype)
    bias = torch.randn(N, device="cuda")
    ldc = (N + 7) // 8 * 8
    C = torch.full((M, ldc), 7.0, device="cuda", dtype=dtype)
    call("gt_gemm", dt_of(A), ptr(A), a_mn, A.shape[1], ptr(Bm), b_mn, Bm.shape[1], ptr(C), ldc, M, N, K, ldc,
         ptr(bias), None, 0, EPI_RELU, 0.0, None, 0, ops.GEMM_IMPL)
    Af = (A.t() if a_mn else A).double()
    Bf = (Bm.t() if b_mn else Bm).double()
    ref = torch.relu(Af @ Bf.t() + bias.double())
    assert rel_l2(C[:, :N], ref) < (1e-5 if dtype == torch.float32 else 6e-3)
    assert (C[:, N:] == 0).all()            # pad columns are kept zero


def test_gemm_splitk_accumulate():
    torch.manual_seed(1)
    M, N, K = 40, 24, 5000
    A = torch.randn(K, M, device="cuda")
    Bm = torch.randn(K, N, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    call("gt_gemm", 0, ptr(A), 1, M, ptr(Bm), 1, N, ptr(C), N, M, N, K, N, None, None, 0, EPI_ACCUM | EPI_OUT_F32,
         0.0, None, 0, ops.GEMM_IMPL)
    assert rel_l2(C, A.double().t() @ Bm.double()) < 1e-5


def _ref_attention(qkv, tok_off, nhead):
    n, d3 = qkv.shape
    d = d3 // 3
    dh = d // nhead
    out = torch.zeros(n, d, dtype=qkv.dtype, device=qkv.device)
    for g in range(len(tok_off) - 1):
        lo, hi = tok_off[g], tok_off[g + 1]
        q, k, v = qkv[lo:hi].split(d, dim=1)
        q = q.view(-1, nhead, dh).transpose(0, 1) * dh ** -0.5
        k = k.view(-1, nhead, dh).transpose(0, 1)
        v = v.view(-1, nhead, dh).transpose(0, 1)
        p = torch.softmax(q @ k.transpose(1, 2), -1)
        out[lo:hi] = (p @ v).transpose(0, 1).reshape(hi - lo, d)
    return out


class _Plan:
    pass


def _packed_plan(lens, extra=3):
    p = _Plan()
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    p.B = len(lens)
    p.tok_off = torch.tensor(off, dtype=torch.int32, device="cuda")
    tg = sum(([g] * n for g, n in enumerate(lens)), []) + [-1] * extra
    p.tok_graph = torch.tensor(tg, dtype=torch.int32, device="cuda")
    return p, off


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("nhead,dh", [(4, 32), (4, 64), (2, 8)])
def test_mha_packed_fwd_bwd(dtype, nhead, dh):
    torch.manual_seed(0)
    lens = [1, 2, 33, 70, 129, 5]
    plan, off = _packed_plan(lens)
    n = off[-1] + 3
    d = nhead * dh
    qkv = torch.randn(n, 3 * d, device="cuda").to(dtype).requires_grad_(True)
    out = ops.mha_packed(qkv, plan, nhead)
    w = torch.randn(n, d, device="cuda").to(dtype)
    w[off[-1]:] = 0
    (out.float() * w.float()).sum().backward()
    q64 = qkv.detach().double().requires_grad_(True)
    ref = _ref_attention(q64, off, nhead)
    (ref * w.double()).sum().backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel_l2(out[:off[-1]].detach(), ref[:off[-1]].detach()) < tol
    assert (out[off[-1]:] == 0).all()
    assert rel_l2(qkv.grad[:off[-1]], q64.grad[:off[-1]]) < tol
    assert (qkv.grad[off[-1]:] == 0).all()


def test_mha_key_start_dense_layout():
    """left-padded dense layout of the public TransformerNodeEncoder API: rows before key_start are
    never keys (the reference's -inf key_padding_mask) and receive zero K/V gradient"""
    torch.manual_seed(3)
    nhead, dh, T = 4, 16, 12
    lens_valid = [12, 5, 1]
    plan, off = _packed_plan([T] * 3, extra=0)
    ks = torch.tensor([off[g] + T - lens_valid[g] for g in range(3)], dtype=torch.int32, device="cuda")
    d = nhead * dh
    qkv = torch.randn(3 * T, 3 * d, device="cuda", requires_grad=True)
    out = ops.mha_packed(qkv, plan, nhead, key_start=ks)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    q64 = qkv.detach().double().requires_grad_(True)
    outs = []
    for g in range(3):
        blk = q64[off[g]:off[g + 1]]
        q, k, v = blk.split(d, 1)
        q = q.view(T, nhead, dh).transpose(0, 1) * dh ** -0.5
        k = k.view(T, nhead, dh).transpose(0, 1)
        v = v.view(T, nhead, dh).transpose(0, 1)
        s = q @ k.transpose(1, 2)
        s[:, :, :T - lens_valid[g]] = float("-inf")
        outs.append((torch.softmax(s, -1) @ v).transpose(0, 1).reshape(T, d))
    ref = torch.cat(outs)
    (ref * w.double()).sum().backward()
    assert rel_l2(out.detach(), ref.detach()) < 1e-5
    assert rel_l2(qkv.grad, q64.grad) < 1e-5


@pytest.mark.parametrize("p", [0.1, 0.3, 0.5])
def test_dropout_kernel(p):
    ops.manual_seed(123)
    ops.begin_step("cuda")
    x = torch.ones(1 << 20, 8, device="cuda", requires_grad=True)
    y = ops.dropout(x, p)
    keep = (y != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 3e-3
    assert torch.allclose(y[y != 0], torch.tensor(1 / (1 - p), device="cuda"))
    y.sum().backward()
    assert torch.equal(x.grad, y.detach())            # the backward re-creates the same mask
    y2 = ops.dropout(x, p)                            # next call site: a different mask
    assert not torch.equal(y2, y)
    ops.begin_step("cuda")                            # next step: different again
    y3 = ops.dropout(x, p)
    assert not torch.equal(y3, y)
    # rows / columns are not correlated (hash quality smoke test)
    m = (y.detach() != 0).float()
    assert abs(m.mean(0) - (1 - p)).max() < 5e-3


def test_mha_dropout_matches_masked_reference():
    """the kernel's keep mask is recovered from a run with V = identity blocks, then the dropped
    attention is checked against softmax * mask / (1-p) in fp64, forward and backward"""
    torch.manual_seed(4)
    nhead, dh, n = 1, 64, 48
    p = 0.3
    plan, off = _packed_plan([n], extra=0)
    d = nhead * dh
    ops.manual_seed(9)
    ops.begin_step("cuda")
    qkv = torch.randn(n, 3 * d, device="cuda")
    # pass 1: Q = K = 0 -> uniform probabilities 1/n; V = I (n <= dh) -> out[i, j] = mask[i, j] / (n (1-p))
    probe = torch.zeros_like(qkv)
    probe[:, 2 * d:2 * d + n] = torch.eye(n, device="cuda")
    salt = 77
    out = ops._MHAFn.apply(probe, plan, nhead, None, p, salt, None)
    mask = (out[:, :n] > 0).double()
    assert abs(mask.mean().item() - (1 - p)) < 0.05
    qkv.requires_grad_(True)
    out = ops._MHAFn.apply(qkv, plan, nhead, None, p, salt, None)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    q64 = qkv.detach().double().requires_grad_(True)
    q, k, v = q64.split(d, 1)
    pr = torch.softmax((q * dh ** -0.5) @ k.t(), -1) * mask / (1 - p)
    ref = pr @ v
    (ref * w.double()).sum().backward()
    assert rel_l2(out.detach(), ref.detach()) < 1e-5
    assert rel_l2(qkv.grad, q64.grad) < 1e-5


def test_errors_are_loud():
    x = torch.zeros(4, 6, device="cuda")
    with pytest.raises(RuntimeError, match="gt_dropout"):
        call("gt_dropout", 0, ptr(x), 3, ptr(x), 0.5, None, 0)
    with pytest.raises(TypeError):
        _lib.dt_of(torch.zeros(1, dtype=torch.float16))


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(136, 304, 1000), (128, 64, 64), (1000, 608, 304), (264, 5008, 256), (72, 200, 40)])
@pytest.mark.parametrize("out_f32", [False, True])
def test_gemm_tcgen05_forced(a_mn, b_mn, M, N, K, out_f32):
    """impl=2: the tcgen05/TMEM/TMA kernel only (an ineligible call would raise), every operand
    major-ness, ragged tiles in M, N and K, bias + residual + ReLU epilogue, bf16 and fp32 outputs"""
    torch.manual_seed(0)
    A = (torch.randn((K, M) if a_mn else (M, K), device="cuda")).bfloat16()
    Bm = (torch.randn((K, N) if b_mn else (N, K), device="cuda")).bfloat16()
    bias = torch.randn(N, device="cuda")
    odt = torch.float32 if out_f32 else torch.bfloat16
    resid = torch.randn(M, N, device="cuda").to(odt)
    C = torch.full((M, N), 7.0, device="cuda", dtype=odt)
    flags = EPI_RELU | (EPI_OUT_F32 | _lib.EPI_RESID_F32 if out_f32 else 0)
    call("gt_gemm", 1, ptr(A), a_mn, A.shape[1], ptr(Bm), b_mn, Bm.shape[1], ptr(C), N, M, N, K, N, ptr(bias),
         ptr(resid), N, flags, 0.0, None, 0, 2)
    Af = (A.t() if a_mn else A).double()
    Bf = (Bm.t() if b_mn else Bm).double()
    ref = torch.relu(Af @ Bf.t() + bias.double() + resid.double())
    assert rel_l2(C, ref) < (5e-3 if not out_f32 else 1e-5 * K ** 0.5 + 1e-6)


@pytest.mark.parametrize("M,N,K", [(600, 304, 13512), (128, 608, 530000), (5008, 256, 128)])
def test_gemm_tcgen05_splitk_weight_gradient(M, N, K):
    """dW[n,k] = sum_rows dY[row,n] X[row,k]: both operands MN-major, split-K with vector fp32 reductions"""
    torch.manual_seed(1)
    dY = torch.randn(K, M, device="cuda").bfloat16()
    X = torch.randn(K, N, device="cuda").bfloat16()
    C = torch.zeros(M, N, device="cuda")
    call("gt_gemm", 1, ptr(dY), 1, M, ptr(X), 1, N, ptr(C), N, M, N, K, N, None, None, 0, EPI_ACCUM | EPI_OUT_F32, 0.0, None, 0, 2)
    ref = dY.double().t() @ X.double()
    assert rel_l2(C, ref) < 1e-4


def _mha_raw(qkv, plan, nhead, impl, drop_p=0.0, salt=0):
    n, d3 = qkv.shape
    d = d3 // 3
    dh = d // nhead
    out = torch.empty(n, d, dtype=qkv.dtype, device="cuda")
    lse = torch.empty(nhead * n, dtype=torch.float32, device="cuda")
    call("gt_mha_fwd", dt_of(qkv), ptr(qkv), ptr(plan.tok_graph), ptr(plan.tok_off), None, None, None, n, plan.B, nhead, dh,
         dh ** -0.5, ptr(out), ptr(lse), drop_p, ptr(ops.rng_state("cuda")) if drop_p else None, salt, impl)
    return out, lse


@pytest.mark.parametrize("nhead,dh", [(4, 32), (4, 64)])
@pytest.mark.parametrize("lens", [[1, 2, 33, 70, 129, 5], [27] * 40, [753, 300, 8, 1001], [128, 128, 256]])
def test_mha_tcgen05_forward(nhead, dh, lens):
    """impl=2 (tcgen05/TMEM/TMA kernel only) against the fp64 reference and the CUDA-core kernel:
    block-diagonal packing of short graphs, long graphs spanning many key tiles, tile-aligned lengths"""
    torch.manual_seed(0)
    plan, off = _packed_plan(lens, extra=5)
    n = off[-1] + 5
    d = nhead * dh
    qkv = torch.randn(n, 3 * d, device="cuda").bfloat16()
    out, lse = _mha_raw(qkv, plan, nhead, 2)
    ref = _ref_attention(qkv.double(), off, nhead)
    assert rel_l2(out[:off[-1]], ref[:off[-1]]) < 1e-2
    assert (out[off[-1]:] == 0).all()
    out1, lse1 = _mha_raw(qkv, plan, nhead, 1)
    assert rel_l2(out[:off[-1]], out1[:off[-1]].double()) < 1e-2
    lse, lse1 = lse.view(nhead, n)[:, :off[-1]], lse1.view(nhead, n)[:, :off[-1]]
    assert (lse - lse1).abs().max() < 2e-2


def test_mha_tcgen05_dropout_mask_matches_cuda_core_kernel():
    """both kernels derive the keep mask from the same counter-based hash of (head, query row, key row)"""
    torch.manual_seed(0)
    nhead, dh = 4, 32
    plan, off = _packed_plan([50, 90, 200], extra=0)
    qkv = torch.randn(off[-1], 3 * nhead * dh, device="cuda").bfloat16()
    ops.manual_seed(3)
    ops.begin_step("cuda")
    o2, _ = _mha_raw(qkv, plan, nhead, 2, 0.3, 11)
    o1, _ = _mha_raw(qkv, plan, nhead, 1, 0.3, 11)
    assert rel_l2(o2, o1.double()) < 1.5e-2
    o3, _ = _mha_raw(qkv, plan, nhead, 2, 0.3, 12)
    assert rel_l2(o3, o1.double()) > 0.2


def _mha_bwd_raw(qkv, out, dout, lse, plan, nhead, impl, drop_p=0.0, salt=0):
    n, d3 = qkv.shape
    d = d3 // 3
    dh = d // nhead
    dqkv = torch.full_like(qkv, float("nan"))
    delta = torch.empty(nhead * n, dtype=torch.float32, device="cuda")
    call("gt_mha_bwd", dt_of(qkv), ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(plan.tok_graph), ptr(plan.tok_off), None,
         None, None, n, plan.B, nhead, dh, dh ** -0.5, ptr(dqkv), ptr(delta), drop_p,
         ptr(ops.rng_state("cuda")) if drop_p else None, salt, impl)
    return dqkv


@pytest.mark.parametrize("nhead,dh", [(4, 32), (4, 64)])
@pytest.mark.parametrize("lens", [[1, 2, 33, 70, 129, 5], [27] * 40, [753, 300, 8, 1001], [128, 128, 256]])
@pytest.mark.parametrize("drop_p", [0.0, 0.3])
def test_mha_tcgen05_backward(nhead, dh, lens, drop_p):
    """impl=2 backward (dQ kernel + dK/dV kernel on tcgen05) against fp64 autograd (p = 0) and against the
    CUDA-core backward under the same dropout mask (p > 0)"""
    torch.manual_seed(0)
    plan, off = _packed_plan(lens, extra=5)
    n = off[-1] + 5
    d = nhead * dh
    qkv = torch.randn(n, 3 * d, device="cuda").bfloat16()
    dout = torch.randn(n, d, device="cuda").bfloat16()
    dout[off[-1]:] = 0
    ops.manual_seed(5)
    ops.begin_step("cuda")
    out, lse = _mha_raw(qkv, plan, nhead, 2, drop_p, 21)
    dq2 = _mha_bwd_raw(qkv, out, dout, lse, plan, nhead, 2, drop_p, 21)
    assert torch.isfinite(dq2.float()).all()
    assert (dq2[off[-1]:] == 0).all()
    dq1 = _mha_bwd_raw(qkv, out, dout, lse, plan, nhead, 1, drop_p, 21)
    for blk in range(3):
        a, b = dq2[:off[-1], blk * d:(blk + 1) * d], dq1[:off[-1], blk * d:(blk + 1) * d]
        assert rel_l2(a, b.double()) < 2e-2, blk
    if drop_p == 0.0:
        q64 = qkv.double().requires_grad_(True)
        ref = _ref_attention(q64, off, nhead)
        (ref * dout.double()).sum().backward()
        for blk in range(3):
            a, b = dq2[:off[-1], blk * d:(blk + 1) * d], q64.grad[:off[-1], blk * d:(blk + 1) * d]
            assert rel_l2(a, b) < 2e-2, blk


@pytest.mark.parametrize("impl,dtype", [(1, torch.float32), (1, torch.bfloat16), (2, torch.bfloat16)])
def test_gemm_epilogue_dropout_matches_standalone_mask(impl, dtype):
    """drop(relu(x W^T + b)) in the GEMM epilogue uses the same (vector index -> keep) map as gt_dropout on the
    output matrix, on the CUDA-core and on both tcgen05 epilogue paths; linear() backward re-derives it from y == 0"""
    torch.manual_seed(0)
    M, N, K, p = 300, 136, 64, 0.3
    ops.manual_seed(17)
    ops.begin_step("cuda")
    x = torch.randn(M, K, device="cuda").to(dtype)
    w = torch.randn(N, K, device="cuda").to(dtype)
    b = torch.randn(N, device="cuda")
    y_ref = torch.empty(M, N, device="cuda", dtype=dtype)
    call("gt_gemm", dt_of(x), ptr(x), 0, K, ptr(w), 0, K, ptr(y_ref), N, M, N, K, N, ptr(b), None, 0, EPI_RELU, 0.0, None, 0, impl)
    want = torch.empty_like(y_ref)
    call("gt_dropout", dt_of(y_ref), ptr(y_ref), y_ref.numel(), ptr(want), p, ptr(ops.rng_state("cuda")), 5)
    got = torch.empty_like(y_ref)
    call("gt_gemm", dt_of(x), ptr(x), 0, K, ptr(w), 0, K, ptr(got), N, M, N, K, N, ptr(b), None, 0, EPI_RELU, p,
         ptr(ops.rng_state("cuda")), 5, impl)
    assert rel_l2(got, want.double()) < (1e-6 if dtype == torch.float32 else 1e-2)
    assert ((got == 0) == (want == 0)).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_relu_dropout_and_layernorm_dropout_autograd(dtype):
    """fused dropout sites of the transformer layer against torch autograd with the recovered masks"""
    torch.manual_seed(1)
    ops.set_precision("fp32" if dtype == torch.float32 else "bf16")
    try:
        M, d, f, p = 200, 64, 128, 0.25
        ops.manual_seed(3)
        ops.begin_step("cuda")
        x = torch.randn(M, d, device="cuda").to(dtype).requires_grad_(True)
        w1 = torch.randn(f, d, device="cuda", requires_grad=True)
        b1 = torch.randn(f, device="cuda", requires_grad=True)
        ln = torch.nn.LayerNorm(d).cuda()
        a = torch.randn(M, d, device="cuda").to(dtype).requires_grad_(True)
        h = ops.linear(x, w1, b1, relu=True, drop_p=p)
        y = ops.layer_norm(a, ln, resid=x, drop_p=p)
        gh, gy = torch.randn_like(h), torch.randn_like(y)
        (h.float() * gh.float()).sum().backward(retain_graph=True)
        gx_h, gw1 = x.grad.clone(), w1.grad.clone()
        x.grad = None
        (y.float() * gy.float()).sum().backward()
        # reference with masks recovered from the outputs
        xr = x.detach().double().requires_grad_(True)
        w1r, b1r = w1.detach().double().requires_grad_(True), b1.detach().double()
        pre = torch.relu(xr @ (w1r.to(dtype).double() if dtype != torch.float32 else w1r).t() + b1r)
        mask_h = (h.detach() != 0) | (pre.detach() == 0)
        href = pre * mask_h / (1 - p)
        (href * gh.double()).sum().backward()
        tol = 1e-5 if dtype == torch.float32 else 2e-2
        assert rel_l2(h.detach(), href.detach()) < tol
        assert rel_l2(gx_h, xr.grad) < tol and rel_l2(gw1, w1r.grad) < tol
        pos = pre.detach() > 0
        keep = abs((h.detach() != 0)[pos].double().mean().item() - (1 - p))
        assert keep < 0.03
        # layer norm: recover the mask of `a` from the gradient (d a = d presum * mask / (1-p))
        ar = a.detach().double().requires_grad_(True)
        xr2 = x.detach().double().requires_grad_(True)
        ratio = (a.grad.double() / x.grad.double())
        mask_a = ratio.abs() > 0.5
        yref = torch.nn.functional.layer_norm(ar * mask_a / (1 - p) + xr2, (d,), ln.weight.double(), ln.bias.double())
        (yref * gy.double()).sum().backward()
        assert rel_l2(y.detach(), yref.detach()) < tol
        assert rel_l2(a.grad, ar.grad) < tol and rel_l2(x.grad, xr2.grad) < tol
        assert abs(mask_a.double().mean().item() - (1 - p)) < 0.03
    finally:
        ops.set_precision("fp32")


def test_fused_losses_match_torch():
    torch.manual_seed(0)
    B, C = 37, 5002
    base = torch.randn(B, 5008, device="cuda", requires_grad=True)
    pred = base[:, :C]                                   # row-strided view like the model's padded logits
    tgt = torch.randint(0, C, (B, 5), device="cuda")
    loss = ops.cross_entropy_mean(pred, tgt[:, 2])
    loss.backward()
    ref_in = base.detach().double().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in[:, :C], tgt[:, 2])
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5
    assert rel_l2(base.grad, ref_in.grad) < 1e-5
    x = torch.randn(64, 128, device="cuda", requires_grad=True)
    y = torch.randint(0, 2, (64, 128), device="cuda").float()
    y[torch.rand(64, 128, device="cuda") < 0.6] = float("nan")
    loss = ops.bce_with_logits_masked_mean(x, y) / 2.0
    loss.backward()
    xr = x.detach().double().requires_grad_(True)
    lab = y == y
    ref = torch.nn.functional.binary_cross_entropy_with_logits(xr[lab], y.double()[lab]) / 2.0
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-6
    assert rel_l2(x.grad, xr.grad) < 1e-5


@pytest.mark.parametrize("conv", ["gin", "gcn"])
@pytest.mark.parametrize("dtype,d", [(torch.float32, 300), (torch.bfloat16, 72), (torch.float32, 512)])
def test_aggregate_table_gradient_split_kernel(conv, dtype, d):
    """gt_aggregate_table_grad (edge-table gradient over type-sorted edges) against the dense formula
    d_table[t] = sum_{e: type t} norm_e * dout[dst_e] * 1[x[src_e] + table[t] > 0] in fp64, and gt_aggregate_bwd with
    d_table = NULL leaves dx unchanged"""
    from graphtrans_b200 import synth
    from graphtrans_b200._lib import CONV_GCN, CONV_GIN, EDGE_TABLE
    batch = synth.gen_mol(40, seed=9)
    ei = batch.edge_index.cuda()
    N, E, ntypes = batch.batch.numel(), ei.shape[1], 60
    plan = ops.GraphPlan(ei, batch.batch.cuda(), 40)
    ld = ops.ldp(d)
    torch.manual_seed(3)
    x = torch.zeros(N, ld, device="cuda")
    x[:, :d] = torch.randn(N, d, device="cuda")
    x = x.to(dtype).requires_grad_(True)
    table = torch.zeros(ntypes, ld, device="cuda")
    table[:, :d] = torch.randn(ntypes, d, device="cuda") * 0.5
    table.requires_grad_(True)
    etype = torch.randint(0, ntypes, (E,), device="cuda", dtype=torch.int32)
    kind = CONV_GIN if conv == "gin" else CONV_GCN
    sp = (torch.zeros(1, device="cuda") if conv == "gin" else torch.randn(d, device="cuda")).requires_grad_(True)
    y = ops.aggregate(x, plan, kind, d, sp, edge_kind=EDGE_TABLE, etype=etype, table=table)
    gy = torch.zeros(N, ld, device="cuda")
    gy[:, :d] = torch.randn(N, d, device="cuda")
    gy = gy.to(dtype)
    dx, dtab = torch.autograd.grad(y, (x, table), gy)
    src, dst = ei[0], ei[1]
    xd, gd, td = x.detach().double()[:, :d], gy.double()[:, :d], table.detach().double()[:, :d]
    nrm = torch.ones(E, device="cuda", dtype=torch.float64)
    if conv == "gcn":
        deg = torch.bincount(src, minlength=N).double() + 1
        nrm = deg[src].rsqrt() * deg[dst].rsqrt()
    mask = (xd[src] + td[etype.long()]) > 0
    gm = nrm[:, None] * gd[dst] * mask
    ref_tab = torch.zeros(ntypes, d, device="cuda", dtype=torch.float64).index_add_(0, etype.long(), gm)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert (dtab[:, :d].double() - ref_tab).norm() / ref_tab.norm() < tol
    assert float(dtab[:, d:].abs().max()) == 0.0 if ld > d else True
    ref_dx = torch.zeros(N, d, device="cuda", dtype=torch.float64).index_add_(0, src, gm)
    if conv == "gin":
        ref_dx += gd
    else:
        ref_dx += gd * ((xd + sp.detach().double()) > 0) / deg[:, None]
    assert (dx[:, :d].double() - ref_dx).norm() / ref_dx.norm() < (1e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("dtype,N", [(torch.bfloat16, 5000), (torch.bfloat16, 300), (torch.float32, 5000)])
def test_embed_sum_gradient_onehot_contraction(dtype, N):
    """embedding-sum backward: small tables go through gt_onehot + gt_gemm + gt_embed_unpack in bf16 mode (N >= 2048),
    a big table (Code2 attribute vocabulary) and every other case through the atomics kernel; all against index_add_ in
    fp64, bit-for-bit integer semantics (clamped indices, strided index columns)"""
    torch.manual_seed(5)
    d = 300
    dims = [119, 4, 12, 3000, 21]
    clamps = [118, 3, 11, 2999, 20]
    xi = torch.stack([torch.randint(0, v + (5 if i == 4 else 0), (N,)) for i, v in enumerate(dims)], 1).cuda()  # col 4 exceeds its clamp
    tabs = [torch.randn(v, d, device="cuda", requires_grad=True) for v in dims]
    out = ops.embed_sum([xi[:, c] for c in range(len(dims))], tabs, clamps=clamps, dtype=dtype)
    ref = sum(t.detach().double()[xi[:, c].clamp(max=clamps[c])] for c, t in enumerate(tabs))
    assert rel_l2(out[:, :d].double(), ref) < (1e-6 if dtype == torch.float32 else 1e-2)
    ld = out.shape[1]
    g = torch.zeros(N, ld, device="cuda")
    g[:, :d] = torch.randn(N, d, device="cuda")
    g = g.to(dtype)
    grads = torch.autograd.grad(out, tabs, g)
    for c, (t, gt_) in enumerate(zip(tabs, grads)):
        r = torch.zeros(dims[c], d, device="cuda", dtype=torch.float64).index_add_(0, xi[:, c].clamp(max=clamps[c]), g[:, :d].double())
        assert gt_.shape == t.shape
        assert rel_l2(gt_.double(), r) < 1e-5, (c, rel_l2(gt_.double(), r))


def test_segment_sum_sorted_long_and_empty_graphs():
    """global_add_pool over contiguous graph ranges: block-per-(graph, chunk) kernel with graphs of 0, 1 and 3000 rows"""
    import types
    sizes = [3, 0, 3000, 1, 0, 257]
    off = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32, device="cuda")
    N, ld = int(off[-1]), 304
    torch.manual_seed(0)
    for dtype in (torch.float32, torch.bfloat16):
        x = torch.randn(N, ld, device="cuda").to(dtype)
        init = torch.randn(len(sizes), ld, device="cuda")
        plan = types.SimpleNamespace(node_off=off, B=len(sizes))
        out = ops.segment_sum(x, plan, init=init)
        ref = init.double().clone()
        for g in range(len(sizes)):
            ref[g] += x[int(off[g]):int(off[g + 1])].double().sum(0)
        assert (out.double() - ref).abs().max() < (1e-3 if dtype == torch.float32 else 1e-3) * max(1.0, float(ref.abs().max()))


def _small_graph_plan(lens, max_nodes):
    import numpy as np
    batch = torch.from_numpy(np.repeat(np.arange(len(lens)), lens)).cuda()
    ei = torch.zeros(2, 0, dtype=torch.long, device="cuda")
    return ops.GraphPlan(ei, batch, len(lens), 1000, cls=True, max_nodes=max_nodes)


@pytest.mark.parametrize("nhead,dh", [(4, 32), (4, 64), (2, 64)])
@pytest.mark.parametrize("drop_p", [0.0, 0.3])
@pytest.mark.parametrize("lens", [[5, 60, 1, 33, 127, 2, 64, 64, 17], [26] * 40, [127, 127, 1, 1, 1, 126]])
def test_mha_tile_local_matches_streamed_kernels(nhead, dh, drop_p, lens):
    """graph-aligned single-tile attention (gt_mha_local_*) against the streamed tcgen05 kernels and the fp32 CUDA-core
    kernels on the same packed tokens: forward, lse and all three input gradients, with the shared dropout hash"""
    torch.manual_seed(11)
    plan_loc = _small_graph_plan(lens, max(lens))
    plan_gen = _small_graph_plan(lens, None)
    assert plan_loc.loc_tiles is not None and plan_gen.loc_tiles is None
    tiles = plan_loc.loc_tiles.view(-1, 2).cpu()
    cnt = int(plan_loc.loc_count)
    tok_off = plan_loc.tok_off.cpu()
    # tiles partition the token rows at graph boundaries, each within 128 rows
    assert cnt <= plan_loc.loc_max_tiles and int(tiles[:cnt, 1].sum()) == int(tok_off[-1])
    assert int(tiles[:cnt, 1].max()) <= 128 and all(int(t) in set(tok_off.tolist()) for t in tiles[:cnt, 0])
    assert bool((tiles[cnt:] == 0).all())
    d = nhead * dh
    n_rows = plan_loc.n_rows
    qkv = (torch.randn(n_rows, 3 * d, device="cuda") * 0.7).bfloat16().requires_grad_(True)
    go = torch.randn(n_rows, d, device="cuda").bfloat16()
    ops.manual_seed(77)
    ops.begin_step("cuda")
    res = []
    for plan, impl in ((plan_loc, None), (plan_gen, None)):
        ops._salt[0] = 0
        o = ops._MHAFn.apply(qkv, plan, nhead, None, drop_p, 5 if drop_p else 0, impl)
        (g,) = torch.autograd.grad(o, qkv, go)
        res.append((o.float(), g.float()))
    n_tok = int(tok_off[-1])
    assert rel_l2(res[0][0][:n_tok], res[1][0][:n_tok]) < 1e-2
    assert rel_l2(res[0][1][:n_tok], res[1][1][:n_tok]) < 2e-2
    # and against the exact fp32 CUDA-core kernels
    q32 = qkv.detach().float().requires_grad_(True)
    o32 = ops._MHAFn.apply(q32, plan_gen, nhead, None, drop_p, 5 if drop_p else 0, 1)
    (g32,) = torch.autograd.grad(o32, q32, go.float())
    assert rel_l2(res[0][0][:n_tok], o32[:n_tok]) < 1.5e-2
    assert rel_l2(res[0][1][:n_tok], g32[:n_tok]) < 3e-2


@pytest.mark.parametrize("H,N,K,M", [(5, 5002, 256, 128), (3, 50, 64, 7)])
def test_stacked_heads_match_per_head_linears(H, N, K, M):
    """all prediction heads as one contraction over the stacked operand copy (+ fused CE over the stacked logits) against
    H separate torch Linears + F.cross_entropy: logits, loss and the gradients of x, every weight and every bias"""
    import torch.nn.functional as F
    from graphtrans_b200 import factory
    torch.manual_seed(2)
    heads = torch.nn.ModuleList(torch.nn.Linear(K, N) for _ in range(H)).cuda()
    ops.set_precision("bf16")
    try:
        reg = ops.W16Registry()
        reg.register_heads(heads)
        reg.register(heads)
        assert not reg.params                                   # stacked weights are not copied twice
        ops.begin_step("cuda", registry=reg)
        x = torch.randn(M, K, device="cuda").bfloat16().requires_grad_(True)
        y_arr = torch.randint(0, N, (M, H), device="cuda")
        st = ops.stacked_heads(x, heads)
        assert st is not None
        y, rp = st
        v = y.view(M, H, rp)
        preds = ops.PredList(v[:, h, :N] for h in range(H))
        preds.stacked = (y, rp, N)
        import types
        args = types.SimpleNamespace(dataset="code2")
        loss = factory.loss_fn(args)(preds, types.SimpleNamespace(y_arr=y_arr))
        params = [p for l in heads for p in (l.weight, l.bias)]
        grads = torch.autograd.grad(loss, [x] + params)
        ops.join_side_streams()
        # reference in fp32 on the bf16-rounded operands
        xr = x.detach().float().requires_grad_(True)
        ref_loss = 0
        for h, l in enumerate(heads):
            wq = l.weight.detach().bfloat16().float().requires_grad_(True)
            bq = l.bias.detach().clone().requires_grad_(True)
            lg = xr @ wq.t() + bq
            assert rel_l2(preds[h].float(), lg.detach()) < 1e-2
            ref_loss = ref_loss + F.cross_entropy(lg, y_arr[:, h])
            l._ref = (wq, bq)
        ref_loss = ref_loss / H
        assert abs(float(loss) - float(ref_loss)) < 2e-2 * max(1.0, abs(float(ref_loss)))
        rg = torch.autograd.grad(ref_loss, [xr] + [t for l in heads for t in l._ref])
        assert float(y.view(M, H, rp)[:, :, N:].abs().max()) == 0.0 if rp > N else True
        for a, b in zip(grads, rg):
            assert rel_l2(a.float(), b) < 3e-2, (a.shape, rel_l2(a.float(), b))
    finally:
        ops.set_precision("fp32")


@pytest.mark.parametrize("M,N,K", [(13233, 600, 300), (512, 300, 600), (129, 72, 64), (70, 36, 40)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_linear_with_fused_column_statistics(M, N, K, dtype):
    """ops.linear(col_stats=True): the BatchNorm column sums / sums of squares taken in the GEMM epilogue (tcgen05 TMA-store
    path; other paths run gt_colstats inside gt_gemm_stats) equal fp64 sums over the STORED output, and batch_norm on
    that output matches batch_norm with its own statistics pass"""
    torch.manual_seed(1)
    ops.set_precision("bf16" if dtype == torch.bfloat16 else "fp32")
    try:
        ops.begin_step("cuda")
        lin = torch.nn.Linear(K, N).cuda()
        ld_in = ops.ldp(K)
        x = torch.zeros(M, ld_in, device="cuda")
        x[:, :K] = torch.randn(M, K, device="cuda")
        x = x.to(dtype)
        y = ops.linear(x, lin.weight, lin.bias, col_stats=True)
        st = y._gt_colstats.clone()
        ld = y.shape[1]
        yd = y.detach().double()
        assert torch.allclose(st[:ld], yd.sum(0), rtol=1e-6, atol=1e-3 * max(1.0, float(yd.abs().max())))
        assert torch.allclose(st[ld:], (yd * yd).sum(0), rtol=1e-6, atol=1e-3 * max(1.0, float((yd * yd).max())))
        y0 = ops.linear(x, lin.weight, lin.bias)
        assert torch.equal(y0, y.detach())
        bn = torch.nn.BatchNorm1d(N).cuda().train()
        bn2 = torch.nn.BatchNorm1d(N).cuda().train()
        a = ops.batch_norm(y, bn, relu=True)
        b = ops.batch_norm(y0, bn2, relu=True)
        assert rel_l2(a.float(), b.float()) < 1e-5
        assert torch.allclose(bn.running_var, bn2.running_var, rtol=1e-5, atol=1e-7)
    finally:
        ops.set_precision("fp32")


def test_mha_tile_local_fails_loudly_on_a_wrong_size_hint():
    """a batch whose largest graph exceeds the host-side max_nodes bound must not silently skip graphs: the tile builder
    flags it on the device and the forward writes NaN"""
    lens = [20, 200, 30]
    plan = _small_graph_plan(lens, 30)          # lie: the second graph has 200 nodes
    assert plan.loc_tiles is not None and int(plan.loc_count) < 0
    qkv = torch.randn(plan.n_rows, 3 * 128, device="cuda").bfloat16()
    o = ops.mha_packed(qkv, plan, 4)
    assert torch.isnan(o.float()).any()
