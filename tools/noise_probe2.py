"""debug aid: (1) per-parameter error of graph + side streams vs eager, with wgrad / branch streams toggled separately;
(2) bucketed replays vs eager over a stream of batches."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200 import factory, loader, ops, synth  # noqa: E402
from graphtrans_b200.ddp import GradBuckets  # noqa: E402
from graphtrans_b200.graphed import GraphedStep  # noqa: E402

ops.set_precision("fp32")
args = synth.make_args("molpcba", gnn_dropout=0.0, transformer_dropout=0.0)
torch.manual_seed(0)
model0 = factory.build_model(args).cuda().train()
init = copy.deepcopy(model0.state_dict())
lossf = factory.loss_fn(args)


def run(batch, wg, br, graph, bucket=False):
    ops.enable_wgrad_stream(wg)
    ops.enable_branch_stream(br)
    m = factory.build_model(args).cuda().train()
    m.load_state_dict(init)
    gb = GradBuckets(m, n_buckets=2, overlap=False)
    if graph:
        st = GraphedStep(m, lossf, gb, bucket=bucket)
        st(batch if bucket else batch.to("cuda"))
    else:
        gb.zero_grad()
        b = batch.to("cuda")
        loss = lossf(m(b), b)
        loss.backward()
        ops.join_side_streams()
    torch.cuda.synchronize()
    ops.enable_wgrad_stream(False)
    ops.enable_branch_stream(False)
    return {k: p.grad.detach().clone() for k, p in m.named_parameters()}


def report(name, g, g0):
    tot = sum(float((g[k].double() - g0[k].double()).pow(2).sum()) for k in g0) ** 0.5 / sum(float(g0[k].double().pow(2).sum()) for k in g0) ** 0.5
    gn = sum(float(g0[k].double().pow(2).sum()) for k in g0) ** 0.5
    worst = sorted(((float((g[k].double() - g0[k].double()).norm()) / gn, k) for k in g0), reverse=True)[:4]
    print(f"{name:34s} rel {tot:.2e}  top: " + ", ".join(f"{k}={v:.1e}" for v, k in worst), flush=True)


hb = synth.make_batch(args, B=32, seed=1)
g0 = run(hb, False, False, False)
for name, kw in (("graph wgrad only", dict(wg=True, br=False, graph=True)), ("graph branch only", dict(wg=False, br=True, graph=True)),
                 ("graph both", dict(wg=True, br=True, graph=True)), ("eager both", dict(wg=True, br=True, graph=False))):
    report(name, run(hb, **kw), g0)

# (2) stream of bucketed replays against eager, same weights
m = factory.build_model(args).cuda().train()
m.load_state_dict(init)
gb = GradBuckets(m, n_buckets=2, overlap=False)
st = GraphedStep(m, lossf, gb, bucket=True)
ref = factory.build_model(args).cuda().train()
rb = GradBuckets(ref, n_buckets=1, overlap=False)
for i in range(60):
    hb = synth.make_batch(args, B=256, seed=100 + i)
    c0 = st.captures
    loss = float(st(hb))
    ref.load_state_dict(m.state_dict())
    rb.zero_grad()
    b = hb.to("cuda")
    le = lossf(ref(b), b)
    le.backward()
    torch.cuda.synchronize()
    e = float((gb.flat.double() - rb.flat.double()).norm() / rb.flat.double().norm())
    if e > 1e-4 or i % 10 == 0:
        gp = {k: p.grad for k, p in m.named_parameters()}
        gr = {k: p.grad for k, p in ref.named_parameters()}
        report(f"batch {i} captured={st.captures - c0} N={hb.batch.numel()} loss d={abs(loss - float(le.detach())):.1e}", gp, gr)
    del le
