"""GPU tests of the round-2 additions, all through the C ABI:
  * pooled-query last encoder layer == full last layer on everything the loss depends on (incl. dropout masks)
  * shape-bucket padding (slack nodes / edges) leaves logits, loss and every gradient unchanged
  * collate-time CSR == device CSR (bit-exact); single-blob batches; bucketed CUDA-graph capture covers many batches
  * global mean / max pooling, row argmax, gradient clipping inside the fused AdamW
  * eval path: BatchNorm folding on / off agree, argmax read-out
"""
import copy

import pytest
import torch

from graphtrans_b200 import _lib, factory, loader, ops, synth
from graphtrans_b200.ddp import GradBuckets
from graphtrans_b200.graphed import GraphedStep
from graphtrans_b200.modules import transformer_encoder as te
from tests.helpers import as_list, grad_report, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _small(cfg, **kw):
    base = dict(gnn_dropout=0.0, transformer_dropout=0.0)
    if cfg in ("code2", "code2-pna"):
        base["num_tasks"] = 300
    base.update(kw)
    args = synth.make_args(cfg, **base)
    return args


def _batch(args, B, seed):
    b = synth.make_batch(args, B=B, seed=seed)
    if args.dataset == "code2":
        b.y_arr = b.y_arr % args.num_tasks
    if args.model_type in ("pna-transformer", "pna") and args.deg is None:
        args.deg = synth.in_degree_histogram(b, 800)
    return b


def _run(model, lossf, b):
    model.zero_grad(set_to_none=True)
    pred = model(b)
    loss = lossf(pred, b)
    loss.backward()
    ops.join_side_streams()
    torch.cuda.synchronize()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    return [t.detach().clone() for t in as_list(pred)], float(loss.detach()), grads


# ----------------------------------------------------------------------------- pooled-query last layer
@pytest.mark.parametrize("nhead,dh", [(4, 32), (4, 64), (8, 32)])      # d = 256 in bf16: the warp-per-graph kernels
@pytest.mark.parametrize("drop_p", [0.0, 0.3])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pooled_query_attention_matches_full_attention_rows(nhead, dh, drop_p, dtype):
    """gt_mha_cls_* against the CUDA-core full attention restricted to the pooled rows: output, dq, dk, dv - with the
    SAME dropout mask (row id = packed row of the query, column = key row).  tests/conftest.py sets GT_CLS_WIDE=2, so the
    d = 256 bf16 cases run the warp-per-graph kernels although the batch has only 8 graphs."""
    import numpy as np
    torch.manual_seed(5)
    lens = [5, 60, 1, 33, 200, 2, 64, 17]
    batch = torch.from_numpy(np.repeat(np.arange(len(lens)), lens)).cuda()
    # max_input_len = 150: the 200-node graph is truncated, so the static row bound N + B leaves unused tail rows
    plan = ops.GraphPlan(torch.zeros(2, 0, dtype=torch.long, device="cuda"), batch, len(lens), 150, cls=True)
    d, n_rows, B = nhead * dh, plan.n_rows, len(lens)
    qkv32 = (torch.randn(n_rows, 3 * d, device="cuda") * 0.7)
    go_full = torch.zeros(n_rows, d, device="cuda")
    rows = plan.cls_rows.long()
    go_cls = torch.randn(B, d, device="cuda")
    go_full[rows] = go_cls
    ops.manual_seed(9)
    ops.begin_step("cuda")
    ref_in = qkv32.clone().requires_grad_(True)
    o_full = ops._MHAFn.apply(ref_in, plan, nhead, None, drop_p, 7 if drop_p else 0, 1)      # exact fp32 kernels
    (g_full,) = torch.autograd.grad(o_full, ref_in, go_full)
    x = qkv32.to(dtype)
    q = x[rows, :d].contiguous().requires_grad_(True)
    kv = x[:, d:].contiguous().requires_grad_(True)
    o = ops._MHAClsFn.apply(q, kv, plan, nhead, drop_p, 7 if drop_p else 0)
    dq, dkv = torch.autograd.grad(o, (q, kv), go_cls.to(dtype))
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    n_tok = int(plan.tok_off[-1])
    assert n_tok < n_rows
    assert rel_l2(o.float(), o_full[rows]) < tol
    assert rel_l2(dq.float(), g_full[rows, :d]) < tol
    assert rel_l2(dkv.float()[:n_tok], g_full[:n_tok, d:]) < tol
    assert float(dkv.float()[n_tok:].abs().max()) == 0.0        # unused tail rows are cleared


@pytest.mark.parametrize("name", ["gcn_virtual_cat_code2", "gin_virtual_cat_mol", "gcn_plain_nci1", "pna_code2"])
def test_pooled_last_layer_equals_full_last_layer(name):
    fx = load_golden(name)
    ops.set_precision("fp32")
    res = []
    for flag in (1, 0):
        te.POOLED_LAST = flag
        try:
            model = factory.build_model(fx["args"]).cuda().train()
            model.load_state_dict(fx["init_sd"])
            res.append(_run(model, factory.loss_fn(fx["args"]), fx["batch"].clone().to("cuda")))
        finally:
            te.POOLED_LAST = 1
    for a, b in zip(res[0][0], res[1][0]):
        assert rel_l2(a, b) < 1e-5
    glob, worst, key = grad_report(res[0][2], {k: v.cpu() for k, v in res[1][2].items()})
    assert glob < 1e-4 and worst < 1e-3, (glob, worst, key)


# ----------------------------------------------------------------------------- shape buckets / collate-time work
def test_collate_time_csr_is_bit_exact():
    args = _small("code2")
    hb = _batch(args, 12, seed=4)
    loader.attach_csr(hb)
    db = hb.to("cuda")
    plan_dev = ops.GraphPlan(db.edge_index, db.batch, db.num_graphs)
    plan_pre = ops.plan_for(db)
    for k in ("rowptr_dst", "src_by_dst", "eid_by_dst", "rowptr_src", "dst_by_src", "eid_by_src"):
        assert torch.equal(getattr(plan_dev, k), getattr(plan_pre, k)), k
    assert plan_pre.rowptr_dst.data_ptr() == db.csr_rowptr_dst.data_ptr()       # no rebuild, no copy


def test_packed_batch_is_one_blob():
    args = _small("molpcba")
    hb = loader.prepare(_batch(args, 16, seed=2))
    assert hb._blob.is_pinned()
    db = hb.to("cuda", non_blocking=True)
    lo, hi = db._blob.data_ptr(), db._blob.data_ptr() + db._blob.numel()
    for k, v in db.tensors():
        assert lo <= v.data_ptr() < hi, k
        assert torch.equal(v.cpu().contiguous().view(-1).view(torch.uint8),
                           getattr(hb, k).contiguous().view(-1).view(torch.uint8)), k      # bytes (labels hold NaNs)
    assert db.slack and db.num_graphs == 16 and int(db.batch.max()) == 16


@pytest.mark.parametrize("cfg,B", [("molpcba", 24), ("code2", 5), ("syn", 6), ("code2-pna", 5), ("nci1", 16)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_bucket_padding_changes_nothing(cfg, B, precision):
    """slack nodes / edges belong to no graph: logits, loss and every parameter gradient of the padded batch equal those
    of the original batch (train-mode BatchNorm statistics included)"""
    ops.set_precision(precision)
    try:
        args = _small(cfg)
        hb = _batch(args, B, seed=3)
        torch.manual_seed(0)
        model = factory.build_model(args).cuda().train()
        lossf = factory.loss_fn(args)
        init = copy.deepcopy(model.state_dict())
        p0, l0, g0 = _run(model, lossf, hb.clone().to("cuda"))
        bufs0 = {k: v.clone() for k, v in model.named_buffers()}
        model.load_state_dict(init)
        pb = loader.prepare(hb.clone())
        assert pb.batch.numel() > hb.batch.numel() and pb.edge_index.shape[1] >= hb.edge_index.shape[1]
        p1, l1, g1 = _run(model, lossf, pb.to("cuda"))
        tol = 2e-5 if precision == "fp32" else 2e-2
        for a, b in zip(p1, p0):
            assert a.shape == b.shape and rel_l2(a, b) < tol
        assert abs(l1 - l0) < tol * max(1.0, abs(l0))
        glob, worst, key = grad_report(g1, {k: v.cpu() for k, v in g0.items()})
        # rounding noise (fp32: different partial-sum partitions; bf16: different tile composition) amplified by the
        # train-mode BatchNorm of the virtual-node MLP over a few dozen graphs
        assert glob < (5e-3 if precision == "fp32" else 0.3), (glob, worst, key)
        for k, v in model.named_buffers():       # running statistics see the real rows only
            assert rel_l2(v.double(), bufs0[k].double()) < (1e-5 if precision == "fp32" else 1e-2), k
    finally:
        ops.set_precision("fp32")


def test_bucketed_capture_covers_many_batches():
    """200 molpcba-shaped batches with fresh sizes every step: shape buckets make the captured graphs repeat (the
    reference loop trainers/base_trainer.py:22-33 yields a new (N, E) almost every step), and every replay returns the
    loss of the eager step on the same batch"""
    ops.set_precision("fp32")
    args = _small("molpcba")
    torch.manual_seed(0)
    model = factory.build_model(args).cuda().train()
    lossf = factory.loss_fn(args)
    buckets = GradBuckets(model, n_buckets=2, overlap=False)
    step = GraphedStep(model, lossf, buckets, max_graphs=16, bucket=True)
    ref = factory.build_model(args).cuda().train()
    rb = GradBuckets(ref, n_buckets=1, overlap=False)
    sigs = set()
    for i in range(200):
        hb = _batch(args, 256, seed=100 + i)
        sigs.add((hb.batch.numel(), hb.edge_index.shape[1]))
        loss = float(step(hb))
        if i % 20 == 0:                      # eager check on a model with the same weights and buffers
            ref.load_state_dict(model.state_dict())
            rb.zero_grad()
            le = lossf(ref(hb.to("cuda")), hb.to("cuda"))
            le.backward()
            assert abs(loss - float(le.detach())) < 2e-4 * max(1.0, abs(float(le.detach()))), i
            # padded vs unpadded differ by fp32 rounding (different partial-sum partitions of the BatchNorm statistics);
            # the train-mode BatchNorm of the virtual-node MLP amplifies that to 1e-4 .. 3e-3 in its weight gradients
            # (the fp32 oracle itself is 4e-4 away from its fp64 run)
            assert rel_l2(buckets.flat, rb.flat) < 1e-2, i
    assert len(sigs) > 100                   # the raw shapes almost never repeat ...
    assert step.captures <= 16, step.captures     # ... the bucketed ones do


def test_capture_warmup_does_not_move_bn_buffers_or_rng():
    ops.set_precision("fp32")
    args = synth.make_args("molpcba")
    torch.manual_seed(0)
    model = factory.build_model(args).cuda().train()
    buckets = GradBuckets(model, n_buckets=2, overlap=False)
    step = GraphedStep(model, factory.loss_fn(args), buckets, warmup_iters=2)
    bn = model.gnn_node.batch_norms[0]
    rng0 = ops.rng_state("cuda").clone()
    step(_batch(args, 8, seed=1).to("cuda"))            # 2 warm-ups + capture + ONE replay
    assert int(bn.num_batches_tracked) == 1
    assert int(ops.rng_state("cuda")[1]) == int(rng0[1]) + 1


# ----------------------------------------------------------------------------- read-outs / optimizer
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kind", ["mean", "max", "sum"])
def test_segment_pool_matches_torch(kind, dtype):
    import numpy as np
    torch.manual_seed(1)
    lens = [3, 1, 40, 7, 300, 2]
    batch = torch.from_numpy(np.repeat(np.arange(len(lens)), lens)).cuda()
    plan = ops.GraphPlan(torch.zeros(2, 0, dtype=torch.long, device="cuda"), batch, len(lens))
    x = torch.randn(sum(lens), 136, device="cuda").to(dtype).requires_grad_(True)
    go = torch.randn(len(lens), 136, device="cuda")
    out = ops.segment_pool(x, plan, kind)
    (gx,) = torch.autograd.grad(out, x, go)
    xr = x.detach().double().requires_grad_(True)
    parts = torch.split(xr, lens)
    ref = torch.stack([p.sum(0) if kind == "sum" else p.mean(0) if kind == "mean" else p.max(0).values for p in parts])
    (gr,) = torch.autograd.grad(ref, xr, go.double())
    assert rel_l2(out, ref) < (1e-6 if dtype == torch.float32 else 1e-5)
    assert rel_l2(gx.float(), gr) < (1e-6 if dtype == torch.float32 else 5e-3)


def test_argmax_rows_first_maximum():
    torch.manual_seed(3)
    x = torch.randn(37, 5008, device="cuda")
    x[5, 100] = x[5, 4000] = 50.0            # tie: first index wins (torch.argmax on CUDA returns the first as well)
    got = ops.argmax_rows(x, n_cols=5002)
    assert torch.equal(got, x[:, :5002].argmax(1))
    assert int(got[5]) == 100


def test_grad_clip_inside_fused_adamw():
    """clip_grad_norm_(max_norm) + AdamW of reference trainers/base_trainer.py:34-36 against gt_sumsq + gt_adamw_multi"""
    from graphtrans_b200.optim import FusedAdamW
    ops.set_precision("fp32")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(40, 64), torch.nn.ReLU(), torch.nn.Linear(64, 7)).cuda()
    ref = copy.deepcopy(net)
    buckets = GradBuckets(net, n_buckets=2, overlap=False)
    opt = FusedAdamW(buckets, lr=1e-2, weight_decay=1e-2, max_grad_norm=0.05)
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=1e-2)
    for it in range(3):
        x = torch.randn(32, 40, device="cuda")
        buckets.zero_grad()
        net(x).pow(2).mean().backward()       # plain autograd accumulation into the arena views
        opt.step()
        ropt.zero_grad()
        ref(x).pow(2).mean().backward()
        total = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        ropt.step()
        assert abs(float(opt.grad_norm()) - float(total)) < 1e-4 * float(total)
        assert float(total) > 0.05            # the clip is active
    for a, b in zip(net.parameters(), ref.parameters()):
        assert rel_l2(a, b) < 1e-5


# ----------------------------------------------------------------------------- eval path
@pytest.mark.parametrize("name", ["gin_virtual_cat_mol", "gin_plain_sum_syn", "zgnn_gin_virtual_mean_mol"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_eval_bn_folding_matches_unfolded(name, precision):
    fx = load_golden(name)
    ops.set_precision(precision)
    try:
        model = factory.build_model(fx["args"]).cuda()
        sd = dict(fx["init_sd"])
        sd.update(fx["buffers"])
        model.load_state_dict(sd)
        model.eval()
        b = fx["batch"].clone().to("cuda")
        outs = []
        for flag in (1, 0):
            ops.EVAL_FOLD_BN = flag
            with torch.no_grad():
                model(b)                                   # the first eval forward folds (and casts) the weights
                k0 = ops._lib.kernel_count
                outs.append(([t.clone() for t in as_list(model(b))], ops._lib.kernel_count - k0))
        ops.EVAL_FOLD_BN = 1
        for a, c in zip(outs[0][0], outs[1][0]):
            assert rel_l2(a, c) < (1e-5 if precision == "fp32" else 2e-2)
        assert outs[0][1] < outs[1][1]                       # fewer launches with the BatchNorms folded away
        if precision == "fp32":
            for a, c in zip(outs[0][0], as_list(fx["logits_eval"])):
                assert rel_l2(a, c) < 1e-3
        # a training step in between invalidates the folded weights
        model.train()
        loss = factory.loss_fn(fx["args"])(model(b), b)
        loss.backward()
        with torch.no_grad():
            for p in model.parameters():
                p.add_(0.01 * torch.randn_like(p))
        model.eval()
        with torch.no_grad():
            a1 = [t.clone() for t in as_list(model(b))]
            ops.EVAL_FOLD_BN = 0
            a0 = [t.clone() for t in as_list(model(b))]
            ops.EVAL_FOLD_BN = 1
        for a, c in zip(a1, a0):
            assert rel_l2(a, c) < (1e-5 if precision == "fp32" else 2e-2)
    finally:
        ops.EVAL_FOLD_BN = 1
        ops.set_precision("fp32")


def test_eval_argmax_readout_code2():
    fx = load_golden("gcn_virtual_cat_code2")
    ops.set_precision("bf16")
    try:
        model = factory.build_model(fx["args"]).cuda().eval()
        model.load_state_dict(fx["init_sd"])
        b = fx["batch"].clone().to("cuda")
        with torch.no_grad():
            pred = model(b)
            ids = factory.predict_fn(fx["args"])(pred)
        assert ids.shape == (b.num_graphs, fx["args"].max_seq_len) and ids.dtype == torch.int64
        for h, p in enumerate(pred):
            assert torch.equal(ids[:, h], p.argmax(1))
    finally:
        ops.set_precision("fp32")


def test_prefetch_and_async_loss_match_plain_steps():
    """the pipelined loop (H2D of batch i+1 under step i into persistent staging blobs, loss read one step late) returns
    exactly the losses / gradients of plain synchronous steps on the same batches"""
    ops.set_precision("fp32")
    args = _small("molpcba")
    torch.manual_seed(0)
    model = factory.build_model(args).cuda().train()
    lossf = factory.loss_fn(args)
    buckets = GradBuckets(model, n_buckets=2, overlap=False)
    step = GraphedStep(model, lossf, buckets, bucket=True)
    batches = [_batch(args, 48, seed=300 + i) for i in range(10)]
    plain, grads = [], []
    for hb in batches:
        plain.append(float(step(hb)))
        grads.append(buckets.flat.clone())
    got, pending = [], None
    step.prefetch(batches[0])
    for i, hb in enumerate(batches):
        if i + 1 < len(batches):
            step.prefetch(batches[i + 1])
        h = step.step_async(hb)
        if i == 4:
            torch.cuda.synchronize()
            assert rel_l2(buckets.flat, grads[4]) < 1e-5         # same captured graph, same inputs (atomics order aside)
        if pending is not None:
            got.append(pending.item())
        pending = h
    got.append(pending.item())
    assert len(got) == len(plain) and all(abs(a - b) < 1e-5 * max(1.0, abs(b)) for a, b in zip(got, plain)), (got, plain)


@pytest.mark.parametrize("conv", ["gin", "gcn"])
@pytest.mark.parametrize("d,n_graphs", [(256, 300), (250, 40), (128, 40), (256, 1), (256, -512), (256, -65)])
def test_aggregate_table_gradient_fused_in_adjoint(conv, d, n_graphs):
    """k_agg_bwd4: the edge-table gradient is contracted on the tensor cores inside the adjoint (per-edge gradients staged in
    shared memory, one-hot edge types as the second operand).  Checked against the fp64 dense formula
    d_table[t] = sum_{e: type t} norm_e * dout[dst_e] * 1[x[src_e] + table[t] > 0] and against the separate-GEMM path
    (GT_AGG_TABLE_FUSED=0 semantics via ops.TABLE_GRAD_FUSED); dx and the self-parameter gradient ride along."""
    from graphtrans_b200 import synth
    from graphtrans_b200._lib import CONV_GCN, CONV_GIN, EDGE_TABLE
    if n_graphs < 0:            # config-4 graphs: ~150 nodes and ~20 staging tiles per block (buffer reuse)
        n_graphs = -n_graphs
        batch = synth.gen_syn(n_graphs, seed=11)
        if n_graphs % 2:        # hubs: three nodes with ~1000 self loops each (what the slack nodes of a shape bucket look
            n_all = batch.batch.numel()    # like): the other workers' consecutive rows lie dozens of staging tiles apart
            hubs = torch.tensor([n_all // 3, n_all // 2, n_all - 1]).repeat_interleave(1000)
            batch.edge_index = torch.cat([batch.edge_index, torch.stack([hubs, hubs])], dim=1)
    else:
        batch = synth.gen_mol(n_graphs, seed=11)
    ei = batch.edge_index.cuda()
    N, E, ntypes = batch.batch.numel(), ei.shape[1], 60
    ld = ops.ldp(d)
    torch.manual_seed(5)
    x0 = torch.zeros(N, ld, device="cuda")
    x0[:, :d] = torch.randn(N, d, device="cuda")
    table0 = torch.zeros(ntypes, ld, device="cuda")
    table0[:, :d] = torch.randn(ntypes, d, device="cuda") * 0.5
    etype = torch.randint(0, ntypes, (E,), device="cuda", dtype=torch.int32)
    kind = CONV_GIN if conv == "gin" else CONV_GCN
    sp0 = torch.full((1,), 0.25, device="cuda") if conv == "gin" else torch.randn(d, device="cuda")
    gy = torch.zeros(N, ld, device="cuda")
    gy[:, :d] = torch.randn(N, d, device="cuda")
    gy = gy.bfloat16()

    def run(fused):
        prev = ops.TABLE_GRAD_FUSED
        ops.TABLE_GRAD_FUSED = fused
        try:
            plan = ops.GraphPlan(ei, batch.batch.cuda(), n_graphs)
            x = x0.bfloat16().requires_grad_(True)
            table = table0.clone().requires_grad_(True)
            sp = sp0.clone().requires_grad_(True)
            y = ops.aggregate(x, plan, kind, d, sp, edge_kind=EDGE_TABLE, etype=etype, table=table)
            return torch.autograd.grad(y, (x, table, sp), gy)
        finally:
            ops.TABLE_GRAD_FUSED = prev

    k0 = _lib.kernel_count
    dx, dtab, dsp = run(1)
    n_fused = _lib.kernel_count - k0
    k0 = _lib.kernel_count
    dx_g, dtab_g, dsp_g = run(0)
    assert n_fused < _lib.kernel_count - k0          # no one-hot / GEMM launches on the fused path
    src, dst = ei[0], ei[1]
    xd, gd, td = x0.bfloat16().double()[:, :d], gy.double()[:, :d], table0.double()[:, :d]
    nrm = torch.ones(E, device="cuda", dtype=torch.float64)
    if conv == "gcn":
        deg = torch.bincount(src, minlength=N).double() + 1
        nrm = deg[src].rsqrt() * deg[dst].rsqrt()
    gm = nrm[:, None] * gd[dst] * ((xd[src] + td[etype.long()]) > 0)
    ref_tab = torch.zeros(ntypes, d, device="cuda", dtype=torch.float64).index_add_(0, etype.long(), gm)
    assert (dtab[:, :d].double() - ref_tab).norm() / ref_tab.norm() < 1e-2
    assert (dtab.double() - dtab_g.double()).norm() / dtab_g.double().norm() < 2e-3
    if ld > d:
        assert float(dtab[:, d:].abs().max()) == 0.0
    assert torch.equal(dx, dx_g)                     # same per-edge arithmetic and summation order
    assert (dsp.double() - dsp_g.double()).norm() <= 1e-3 * dsp_g.double().norm() + 1e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,ld", [(5000, 512, 512), (4097, 300, 304), (70, 8, 8)])
def test_relu_bwd_colsum(dtype, M, N, ld):
    """gt_relu_bwd_colsum == gt_relu_bwd followed by gt_colsum (bias gradient of a Linear with a fused ReLU epilogue)"""
    torch.manual_seed(2)
    dy = torch.randn(M, ld, device="cuda").to(dtype)
    y = torch.relu(torch.randn(M, ld, device="cuda")).to(dtype)
    y[:, N:] = 0
    dz = torch.empty_like(dy)
    cs = torch.full((N,), 0.5, device="cuda")
    ops.call("gt_relu_bwd_colsum", ops.dt_of(dy), ops.ptr(dy), ops.ptr(y), M, N, ld, ops.ptr(dz), 1.25, ops.ptr(cs))
    ref32 = torch.where(y > 0, dy.float() * 1.25, torch.zeros((), device="cuda"))
    assert torch.equal(dz, ref32.to(dtype))
    want = ref32.double()[:, :N].sum(0) + 0.5                     # fp32 products summed before the store rounds them;
                                                                  # accumulates into the target
    assert (cs.double() - want).abs().max() <= 1e-4 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("d,n_graphs", [(256, 300), (300, 40), (72, 40), (256, 1)])
def test_gin_adjoint_packed_mask_is_exact(d, n_graphs):
    """k_agg_bwd3p forms the ReLU mask with packed bf16 compares against round-down-to-bf16(-table) thresholds: for a
    bf16 x that is the SAME predicate as x + e > 0 in fp32, so dx and the per-edge gradients (hence d_table) must equal
    the fp32-mask kernel bit for bit.  The table holds values that are not bf16-representable, exact ties x == -e
    (mask false on both paths) and entries just above / below a bf16 value."""
    from graphtrans_b200._lib import CONV_GIN, EDGE_TABLE
    batch = synth.gen_mol(n_graphs, seed=13)
    ei = batch.edge_index.cuda()
    N, E, ntypes = batch.batch.numel(), ei.shape[1], 60
    ld = ops.ldp(d)
    torch.manual_seed(6)
    x0 = torch.zeros(N, ld, device="cuda")
    x0[:, :d] = torch.randn(N, d, device="cuda")
    xb = x0.bfloat16()
    table0 = torch.zeros(ntypes, ld, device="cuda")
    table0[:, :d] = torch.randn(ntypes, d, device="cuda") * 0.5
    # ties and near-ties: -table equal to / one fp32 ulp above / below bf16 values that occur in x
    vals = xb[torch.arange(ntypes, device="cuda") % N, :d].float()
    table0[:, 0:d:7] = -vals[:, 0:d:7]
    table0[:, 1:d:7] = -torch.nextafter(vals[:, 1:d:7], torch.full_like(vals[:, 1:d:7], 1e9))
    table0[:, 2:d:7] = -torch.nextafter(vals[:, 2:d:7], torch.full_like(vals[:, 2:d:7], -1e9))
    etype = torch.randint(0, ntypes, (E,), device="cuda", dtype=torch.int32)
    src = ei[0]
    etype[: min(E, N)] = (src[: min(E, N)] % ntypes).int()      # many edges whose source row hits its own tie row
    gy = torch.zeros(N, ld, device="cuda")
    gy[:, :d] = torch.randn(N, d, device="cuda")
    gy = gy.bfloat16()

    def run(packed):
        prev = ops.AGG_PACKED
        ops.AGG_PACKED = packed
        try:
            plan = ops.GraphPlan(ei, batch.batch.cuda(), n_graphs)
            x = xb.clone().requires_grad_(True)
            table = table0.clone().requires_grad_(True)
            sp = torch.full((1,), 0.25, device="cuda").requires_grad_(True)
            y = ops.aggregate(x, plan, CONV_GIN, d, sp, edge_kind=EDGE_TABLE, etype=etype, table=table)
            return torch.autograd.grad(y, (x, table, sp), gy)
        finally:
            ops.AGG_PACKED = prev

    dx1, dt1, ds1 = run(1)
    dx0, dt0, ds0 = run(0)
    assert torch.equal(dx1, dx0)
    assert (dt1.double() - dt0.double()).norm() <= 1e-5 * dt0.double().norm()      # same gm, split-K reduction order only
    assert (ds1.double() - ds0.double()).abs().max() <= 1e-3 * ds0.double().abs().max() + 1e-6
    # and against the dense fp64 formula
    xd, gd, td = xb.double()[:, :d], gy.double()[:, :d], table0.double()[:, :d]
    gm = gd[ei[1]] * ((xd[src] + td[etype.long()]) > 0)
    ref_dx = torch.zeros(N, d, device="cuda", dtype=torch.float64).index_add_(0, src, gm) + 1.25 * gd
    assert (dx1[:, :d].double() - ref_dx).norm() / ref_dx.norm() < 1e-2
