"""CUDA-graph replay of the training step (graphtrans_b200/graphed.py) must reproduce the eager step: same loss and
gradients for the captured batch, correct results for a second batch of the SAME shape signature (static input
buffers are refilled), a new capture for a different signature, BatchNorm buffers advancing on every replay and
fresh dropout masks per replay."""
import copy

import pytest
import torch

from graphtrans_b200 import factory, ops, synth
from graphtrans_b200.ddp import GradBuckets
from graphtrans_b200.graphed import GraphedStep
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _setup(cfg, drop):
    kw = {} if drop else dict(gnn_dropout=0.0, transformer_dropout=0.0)
    if cfg == "code2":
        kw["num_tasks"] = 300
    args = synth.make_args(cfg, **kw)
    torch.manual_seed(0)
    model = factory.build_model(args).cuda().train()
    return args, model, factory.loss_fn(args)


def _batch(args, B, seed):
    b = synth.make_batch(args, B=B, seed=seed)
    if args.dataset == "code2":
        b.y_arr = b.y_arr % args.num_tasks
    return b


@pytest.mark.parametrize("cfg,precision,side", [("molpcba", "fp32", False), ("molpcba", "fp32", True),
                                                ("code2", "bf16", True)])
def test_replay_matches_eager(cfg, precision, side):
    """side=True: weight gradients and the virtual-node branch on their own streams (parallel graph branches), as
    bench.py runs the step"""
    ops.set_precision(precision)
    ops.enable_wgrad_stream(side)
    ops.enable_branch_stream(side)
    try:
        args, model, lossf = _setup(cfg, drop=False)
        init = copy.deepcopy(model.state_dict())
        buckets = GradBuckets(model, n_buckets=2, overlap=False)
        step = GraphedStep(model, lossf, buckets)
        # molpcba: 32 graphs keep the train-mode BatchNorm of the virtual-node MLP (statistics over B rows) well
        # conditioned; with a handful of graphs fp32 summation-order noise alone exceeds the tolerance
        nb = 32 if cfg == "molpcba" else 6
        b1 = _batch(args, nb, seed=1).to("cuda")
        l_cap = float(step(b1))                          # capture + first replay
        g_cap = buckets.flat.clone()
        # eager reference on a fresh copy of the model
        ref = factory.build_model(args).cuda().train()
        ref.load_state_dict(init)
        rb = GradBuckets(ref, n_buckets=1, overlap=False)
        rb.zero_grad()
        loss = lossf(ref(b1), b1)
        loss.backward()
        ops.join_side_streams()
        loss = float(loss.detach())
        tol = 2e-3 if precision == "fp32" else 5e-2
        assert abs(l_cap - loss) < tol * max(1.0, abs(loss))
        assert rel_l2(g_cap, rb.flat) < tol
        # same signature, different content: permute the graphs' labels and node features
        b2 = b1.clone()
        if cfg == "molpcba":
            b2.y = b2.y.flip(0)
            b2.x = b2.x.flip(1)
        else:
            b2.y_arr = b2.y_arr.flip(0)
            b2.node_depth = (b2.node_depth + 3) % 20
        n_graphs = len(step.cache)
        l2 = float(step(b2))
        assert len(step.cache) == n_graphs               # replay, no new capture
        ref.load_state_dict(model.state_dict())          # BN buffers advanced in the graphed model: start from its state
        model_bn = copy.deepcopy(model.state_dict())
        # (compare against eager on the reference initialised from the state BEFORE this replay is not possible any more,
        #  so check self-consistency: the eager step on the graphed model itself gives the same gradients)
        g2 = buckets.flat.clone()
        model.load_state_dict(model_bn)
        buckets.zero_grad()
        le = lossf(model(b2), b2)
        le.backward()
        ops.join_side_streams()
        le = float(le.detach())                           # drop the autograd graph before the next capture
        assert abs(l2 - le) < tol * max(1.0, abs(le))
        assert rel_l2(buckets.flat, g2) < tol
        # different signature -> new capture
        b3 = _batch(args, nb - 1, seed=7).to("cuda")
        step(b3)
        assert len(step.cache) == n_graphs + 1
    finally:
        ops.set_precision("fp32")
        ops.enable_wgrad_stream(False)
        ops.enable_branch_stream(False)


def test_replay_advances_bn_buffers_and_dropout():
    ops.set_precision("bf16")
    try:
        args, model, lossf = _setup("molpcba", drop=True)
        buckets = GradBuckets(model, n_buckets=2, overlap=False)
        step = GraphedStep(model, lossf, buckets)
        b = _batch(args, 8, seed=3).to("cuda")
        bn = model.gnn_node.batch_norms[0]
        step(b)
        n0 = int(bn.num_batches_tracked)
        l1 = float(step(b))
        l2 = float(step(b))
        assert int(bn.num_batches_tracked) == n0 + 2     # running statistics advance inside the graph
        assert l1 != l2                                  # device-side dropout counter advances on every replay
        assert torch.isfinite(buckets.flat).all()
    finally:
        ops.set_precision("fp32")
