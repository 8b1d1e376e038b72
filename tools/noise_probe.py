"""run-to-run / path-to-path gradient noise of the fp32 step on a small molpcba batch (debug aid):
eager twice, eager with side streams, graph replay, bucket-padded batch."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200 import factory, loader, ops, synth  # noqa: E402
from graphtrans_b200.ddp import GradBuckets  # noqa: E402
from graphtrans_b200.graphed import GraphedStep  # noqa: E402

ops.set_precision("fp32")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
args = synth.make_args("molpcba", gnn_dropout=0.0, transformer_dropout=0.0)
hb = synth.make_batch(args, B=B, seed=1)
torch.manual_seed(0)
model = factory.build_model(args).cuda().train()
init = copy.deepcopy(model.state_dict())
lossf = factory.loss_fn(args)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def eager(batch, side):
    ops.enable_wgrad_stream(side)
    ops.enable_branch_stream(side)
    m = factory.build_model(args).cuda().train()
    m.load_state_dict(init)
    gb = GradBuckets(m, n_buckets=2, overlap=False)
    gb.zero_grad()
    b = batch.to("cuda")
    loss = lossf(m(b), b)
    loss.backward()
    ops.join_side_streams()
    torch.cuda.synchronize()
    ops.enable_wgrad_stream(False)
    ops.enable_branch_stream(False)
    return float(loss.detach()), gb.flat.clone()


def graph(batch, side, bucket=False):
    ops.enable_wgrad_stream(side)
    ops.enable_branch_stream(side)
    m = factory.build_model(args).cuda().train()
    m.load_state_dict(init)
    gb = GradBuckets(m, n_buckets=2, overlap=False)
    st = GraphedStep(m, lossf, gb, bucket=bucket)
    loss = float(st(batch if bucket else batch.to("cuda")))
    torch.cuda.synchronize()
    ops.enable_wgrad_stream(False)
    ops.enable_branch_stream(False)
    return loss, gb.flat.clone()


l0, g0 = eager(hb, False)
for name, fn in (("eager again", lambda: eager(hb, False)), ("eager side streams", lambda: eager(hb, True)),
                 ("graph", lambda: graph(hb, False)), ("graph side streams", lambda: graph(hb, True)),
                 ("graph side streams again", lambda: graph(hb, True)),
                 ("eager padded", lambda: eager(loader.prepare(hb.clone()), False)),
                 ("graph bucketed", lambda: graph(hb, False, True))):
    l, g = fn()
    print(f"B={B} {name:28s} loss diff {abs(l - l0):.2e}  grads rel {rel(g, g0):.2e}", flush=True)
