"""GPU parity tests proper: the CUDA path (through the C ABI) against
 (a) the committed golden fixtures produced by the UNMODIFIED reference in fp64, and
 (b) the CPU oracle on seeded inputs the oracle finishes in seconds.
fp32 mode contract (BASELINE.json north_star / SURVEY §8c): logits <= 1e-3 rel-L2 (expected
~1e-5), gradients <= 1e-3 global rel-L2, per parameter <= 1e-3 relative to
max(|g_p|, 1e-3 |g|_global); integer pad/index/mask work bit-exact (tests/test_index_gpu.py)."""
import copy

import pytest
import torch

from graphtrans_b200 import factory, ops, synth
from oracle import graphtrans_oracle as O
from tests.helpers import GOLDEN_CASES, as_list, grad_report, load_golden, rel_l2

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3
GRAD_TOL = 1e-3
WORST_TOL = 5e-3


def run_product(args, batch, init_sd, precision="fp32", train=True, buckets=False, wgrad_stream=False):
    ops.set_precision(precision)
    model = factory.build_model(args).cuda()
    model.load_state_dict(init_sd, strict=True)
    model.train(train)
    b = batch.clone().to("cuda")
    if buckets:   # fused gradient delivery into the flat DDP arena (graphtrans_b200.ddp.GradBuckets)
        from graphtrans_b200.ddp import GradBuckets
        gb = GradBuckets(model, n_buckets=3)
        gb.flat.fill_(123.0)      # zero_grad() must clear it
        gb.zero_grad()
    else:
        model.zero_grad()
    if wgrad_stream:
        ops.enable_wgrad_stream(True)
        ops.enable_branch_stream(True)
    try:
        pred = model(b)
        loss = factory.loss_fn(args)(pred, b)
        loss.backward()
        ops.join_side_streams()
    finally:
        ops.enable_wgrad_stream(False)
        ops.enable_branch_stream(False)
    torch.cuda.synchronize()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
             for k, p in model.named_parameters()}
    bufs = {k: v.detach().clone() for k, v in model.named_buffers()}
    return model, pred, loss.detach(), grads, bufs


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_fwd_bwd_fp32(name):
    fx = load_golden(name)
    model, pred, loss, grads, bufs = run_product(fx["args"], fx["batch"], fx["init_sd"])
    for a, b in zip(as_list(pred), as_list(fx["logits"])):
        assert a.dtype == torch.float32 and a.shape == b.shape
        assert rel_l2(a.detach(), b) < LOGIT_TOL
    assert abs(float(loss) - float(fx["loss"])) < 1e-4 * max(1.0, abs(float(fx["loss"])))
    glob, worst, key = grad_report(grads, fx["grads"])
    assert glob < GRAD_TOL, (glob, worst, key)
    assert worst < WORST_TOL, (glob, worst, key)
    for k, v in fx["buffers"].items():           # BatchNorm running stats + num_batches_tracked
        assert rel_l2(bufs[k].double(), v.double()) < 1e-4, k
    # eval mode (running statistics, no dropout) against the reference's eval logits
    model.eval()
    with torch.no_grad():
        pe = model(fx["batch"].clone().to("cuda"))
    for a, b in zip(as_list(pe), as_list(fx["logits_eval"])):
        assert rel_l2(a, b) < LOGIT_TOL


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_fwd_bwd_fp32_on_tensor_cores(name):
    """the fp32 contract with EVERY contraction on the tcgen05 kernel (ops.GEMM_TC_PARITY: three-term bf16 split, six
    products accumulated in fp32 in TMEM) - the tensor-core GEMM itself meets the <= 1e-3 bar (SURVEY 7 'fp32 parity on
    tensor cores'), not only its CUDA-core twin"""
    fx = load_golden(name)
    ops.GEMM_TC_PARITY = 1
    try:
        k0 = ops._lib.kernel_count
        model, pred, loss, grads, bufs = run_product(fx["args"], fx["batch"], fx["init_sd"])
        assert ops._lib.kernel_count > k0
    finally:
        ops.GEMM_TC_PARITY = 0
    for a, b in zip(as_list(pred), as_list(fx["logits"])):
        assert rel_l2(a.detach(), b) < LOGIT_TOL
    assert abs(float(loss) - float(fx["loss"])) < 1e-4 * max(1.0, abs(float(fx["loss"])))
    glob, worst, key = grad_report(grads, fx["grads"])
    assert glob < GRAD_TOL and worst < WORST_TOL, (glob, worst, key)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_grads_with_wgrad_side_stream(name):
    """weight / bias / embedding gradients issued on the second stream (ops.enable_wgrad_stream) and the virtual-node
    update on the parallel branch stream (ops.enable_branch_stream): same parity bar"""
    fx = load_golden(name)
    _, pred, loss, grads, _ = run_product(fx["args"], fx["batch"], fx["init_sd"], buckets=True, wgrad_stream=True)
    for a, b in zip(as_list(pred), as_list(fx["logits"])):
        assert rel_l2(a.detach(), b) < LOGIT_TOL
    glob, worst, key = grad_report(grads, fx["grads"])
    assert glob < GRAD_TOL and worst < WORST_TOL, (glob, worst, key)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_grads_delivered_into_flat_arena(name):
    """same parity bar when the backward kernels accumulate parameter gradients straight into the
    flat allreduce arena (no autograd accumulation kernels)"""
    fx = load_golden(name)
    model, pred, loss, grads, bufs = run_product(fx["args"], fx["batch"], fx["init_sd"], buckets=True)
    glob, worst, key = grad_report(grads, fx["grads"])
    assert glob < GRAD_TOL and worst < WORST_TOL, (glob, worst, key)
    for k, p in model.named_parameters():
        assert p.grad.data_ptr() == p._gt_main_grad.data_ptr(), k


@pytest.mark.parametrize("name", ["gcn_virtual_cat_code2", "gin_virtual_cat_mol", "pna_code2"])
def test_golden_bf16_mode_is_close(name):
    """bf16 throughput mode: not covered by the 1e-3 contract (the reference under bf16 autocast is
    itself 1.4e-2 / 0.25 off); sanity bound only."""
    fx = load_golden(name)
    _, pred, loss, grads, _ = run_product(fx["args"], fx["batch"], fx["init_sd"], precision="bf16")
    ops.set_precision("fp32")
    for a, b in zip(as_list(pred), as_list(fx["logits"])):
        assert rel_l2(a.detach(), b) < 0.1
    glob, worst, key = grad_report(grads, fx["grads"])
    assert glob < 0.5, (glob, worst, key)


@pytest.mark.parametrize("cfg,B", [("nci1", 32), ("molpcba", 48), ("code2", 6), ("syn", 8), ("code2-pna", 6)])
def test_oracle_parity_at_config_shapes(cfg, B):
    """BASELINE configs at their real widths (d_g=300/272/256, d=128/256, 5+4 layers) on a reduced
    number of graphs so the fp64 CPU oracle finishes in seconds; dropout 0 (RNG cannot match).
    molpcba uses 48 graphs: with 24 the train-mode BatchNorm of the virtual-node MLP (statistics over B rows) is so
    ill-conditioned that the fp32 ORACLE itself is 1.0e-3 away from its fp64 run (1.8e-4 at B=48)."""
    kw = dict(gnn_dropout=0.0, transformer_dropout=0.0)
    if cfg in ("code2", "code2-pna"):
        kw.update(num_tasks=500)         # 5 heads x 500 classes keep the oracle fast
    args = synth.make_args(cfg, **kw)
    batch = synth.make_batch(args, B=B, seed=11)
    if cfg == "code2-pna":
        args.deg = synth.in_degree_histogram(batch, 800)
    if cfg in ("code2", "code2-pna"):
        batch.y_arr = batch.y_arr % 500
    torch.manual_seed(5)
    init = copy.deepcopy(factory.build_model(args).state_dict())
    _, pred, loss, grads, bufs = run_product(args, batch, init)
    opred, oloss, ograds, obufs = O.fwd_bwd(init, args, batch, dtype=torch.float64)
    for a, b in zip(as_list(pred), as_list(opred)):
        assert rel_l2(a.detach(), b.detach()) < LOGIT_TOL
    assert abs(float(loss) - float(oloss)) < 1e-4 * max(1.0, abs(float(oloss)))
    glob, worst, key = grad_report(grads, ograds)
    assert glob < GRAD_TOL, (glob, worst, key)
    assert worst < WORST_TOL, (glob, worst, key)
    for k, v in obufs.items():
        assert rel_l2(bufs[k].double(), v.double()) < 1e-4, k


def test_truncation_keeps_last_nodes():
    """graphs longer than max_input_len: the packed path must equal the oracle's left-padded,
    keep-the-LAST-L-nodes behaviour (reference modules/utils.py:16,22-24)."""
    args = synth.make_args("code2", gnn_dropout=0.0, transformer_dropout=0.0, num_tasks=40, max_input_len=16,
                           gnn_emb_dim=64, d_model=64, gnn_num_layer=2, num_encoder_layers=2)
    batch = synth.gen_code2(5, seed=3, nmin=4, nmax=60, mu=3.2, sigma=0.7, num_classes=40)
    torch.manual_seed(1)
    init = copy.deepcopy(factory.build_model(args).state_dict())
    _, pred, loss, grads, _ = run_product(args, batch, init)
    opred, oloss, ograds, _ = O.fwd_bwd(init, args, batch, dtype=torch.float64)
    for a, b in zip(as_list(pred), as_list(opred)):
        assert rel_l2(a.detach(), b.detach()) < LOGIT_TOL
    glob, worst, key = grad_report(grads, ograds)
    assert glob < GRAD_TOL and worst < WORST_TOL, (glob, worst, key)


def test_pooling_last():
    args = synth.make_args("nci1", gnn_dropout=0.0, transformer_dropout=0.0, graph_pooling="last", gnn_emb_dim=32,
                           d_model=32, dim_feedforward=64)
    batch = synth.gen_nci1(7, seed=2)
    torch.manual_seed(2)
    init = copy.deepcopy(factory.build_model(args).state_dict())
    _, pred, loss, grads, _ = run_product(args, batch, init)
    opred, oloss, ograds, _ = O.fwd_bwd(init, args, batch, dtype=torch.float64)
    assert rel_l2(pred.detach(), opred.detach()) < LOGIT_TOL
    glob, worst, key = grad_report(grads, ograds)
    assert glob < GRAD_TOL and worst < WORST_TOL, (glob, worst, key)


def test_dropout_training_step_is_finite_and_seeded():
    """with the configured dropouts the step runs, is reproducible under the same seed/step and
    differs between steps; E[output] sanity is covered in tests/test_ops_gpu.py"""
    args = synth.make_args("molpcba")
    batch = synth.make_batch(args, B=16, seed=1).to("cuda")
    torch.manual_seed(0)
    model = factory.build_model(args).cuda().train()
    lossf = factory.loss_fn(args)

    def step(seed):
        ops.manual_seed(seed)
        model.zero_grad()
        loss = lossf(model(batch), batch)
        loss.backward()
        return float(loss.detach()), model.gnn2transformer.weight.grad.clone()

    l1, g1 = step(7)
    l2, g2 = step(7)
    l3, g3 = step(8)
    # same seed -> same masks (the only run-to-run noise left is fp32 atomic summation order)
    # (atomics reorder fp32 sums; train-mode BatchNorm over the 16 graphs of the virtual-node MLP amplifies that noise)
    assert abs(l1 - l2) < 1e-5 and rel_l2(g1, g2) < 2e-3
    assert abs(l1 - l3) > 1e-4 and rel_l2(g3, g1) > 1e-2
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
