"""debug aid: eager padded vs eager unpadded, fp32, over a stream of batches; for the worst batch, the first intermediate
(ops.batch_norm / ops.linear / ops.aggregate / ops.segment_sum outputs, real rows only) that differs."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200 import factory, loader, ops, synth  # noqa: E402

ops.set_precision("fp32")
args = synth.make_args("molpcba", gnn_dropout=0.0, transformer_dropout=0.0)
torch.manual_seed(0)
model = factory.build_model(args).cuda().train()
init = copy.deepcopy(model.state_dict())
lossf = factory.loss_fn(args)
rec = None
orig = {n: getattr(ops, n) for n in ("batch_norm", "linear", "aggregate", "segment_sum", "add_graph_vec", "embed_sum")}


def wrap(name):
    def f(*a, **k):
        out = orig[name](*a, **k)
        if rec is not None:
            rec.append((name, out.detach().clone()))
        return out
    return f


for n in orig:
    setattr(ops, n, wrap(n))
from graphtrans_b200.modules import gnn_module, conv  # noqa: E402,F401


def run(b):
    global rec
    model.load_state_dict(init)
    model.zero_grad(set_to_none=True)
    rec = []
    loss = lossf(model(b), b)
    r, rec = rec, None
    loss.backward()
    torch.cuda.synchronize()
    return r, {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def gerr(g, g0):
    gn = sum(float(g0[k].double().pow(2).sum()) for k in g0) ** 0.5
    return sum(float((g[k].double() - g0[k].double()).pow(2).sum()) for k in g0) ** 0.5 / gn


worst = (0, None)
for i in range(30):
    hb = synth.make_batch(args, B=256, seed=100 + i)
    pb = loader.prepare(hb.clone())
    r0, g0 = run(hb.to("cuda"))
    r1, g1 = run(pb.to("cuda"))
    e = gerr(g1, g0)
    print(f"batch {i} N={hb.batch.numel()} slack={pb.batch.numel() - hb.batch.numel()} E={hb.edge_index.shape[1]} eslack={pb.edge_index.shape[1] - hb.edge_index.shape[1]} grads rel {e:.2e}", flush=True)
    if e > worst[0]:
        worst = (e, i, r0, r1, hb.batch.numel())
e, i, r0, r1, N = worst
print("worst batch", i, e)
for k, ((n0, a), (n1, b)) in enumerate(zip(r0, r1)):
    rows = min(a.shape[0], b.shape[0]) if a.shape[0] != b.shape[0] else a.shape[0]
    if a.shape[0] != b.shape[0]:
        rows = N if a.shape[0] >= N else rows
    d = float((a[:rows].double() - b[:rows].double()).abs().max())
    sc = float(a[:rows].double().abs().max())
    print(f"  op {k:3d} {n0:14s} shape {tuple(a.shape)} vs {tuple(b.shape)} max|diff| {d:.3e} (scale {sc:.2e})")
    if k > 40:
        break
