"""Forward / adjoint timing of the message-passing aggregation on the BASELINE batch shapes (CUDA events, 30 launches
behind a GPU-side head start).  GT_AGG_VARIANT=1 selects the per-edge kernels.  python tools/agg_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphtrans_b200 import factory, ops, synth
from graphtrans_b200._lib import CONV_GCN, CONV_GIN
from graphtrans_b200.modules import conv as conv_mod


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(20e-3 * 1.9e9))
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for cfg in ("molpcba", "code2", "syn"):
    ops.set_precision("bf16")
    args = synth.make_args(cfg)
    b = synth.make_batch(args, B=min(args.batch_size, 1024), seed=0).to("cuda")
    model = factory.build_model(args).cuda()
    plan = ops.GraphPlan(b.edge_index, b.batch, b.num_graphs, 1000)
    N, E, d, ld = b.batch.numel(), b.edge_index.shape[1], args.gnn_emb_dim, ops.ldp(args.gnn_emb_dim)
    conv = model.gnn_node.convs[1]
    enc = conv_mod._edge_encoder_args(conv.edge_encoder, b.edge_attr, plan, d, ld)
    kind, sp = (CONV_GCN, conv.root_emb.weight) if args.gnn_type == "gcn" else (CONV_GIN, conv.eps)
    x = (torch.randn(N, ld, device="cuda") * 0.5).bfloat16()
    x[:, d:] = 0
    x.requires_grad_(True)
    gy = torch.randn(N, ld, device="cuda").bfloat16()
    with torch.no_grad():
        t_f = timed(lambda: ops.aggregate(x, plan, kind, d, sp, **enc))
    y = ops.aggregate(x, plan, kind, d, sp, **enc)
    t_b = timed(lambda: torch.autograd.grad(y, x, gy, retain_graph=True))
    ea = 0 if b.edge_attr is None else b.edge_attr.numel() * b.edge_attr.element_size()
    byt = 2 * N * d * 2 + 16 * E + ea
    print(f"{cfg:8s} N={N} E={E} d={d}: fwd {t_f:7.1f} us ({byt / t_f / 1e3:7.1f} GB/s)   bwd(+grad bookkeeping) {t_b:7.1f} us ({byt / t_b / 1e3:7.1f} GB/s)"
          f"   variant={os.environ.get('GT_AGG_VARIANT', '0')}")
    # per-export timing of the adjoint (CUDA events around every C-ABI call)
    from graphtrans_b200 import _lib
    import collections
    _lib.start_profile()
    for _ in range(10):
        torch.autograd.grad(y, x, gy, retain_graph=True)
    per = collections.defaultdict(list)
    for name, ms, _ in _lib.stop_profile():
        per[name].append(ms * 1e3)
    print("          " + "  ".join(f"{k}: {sum(v) / len(v):.1f} us" for k, v in per.items()))
