"""Tile-local vs streamed tcgen05 attention on the packed tokens of a small-graph batch (CUDA graph of 20 back-to-back
launches, warm L2).  python tools/mha_local_bench.py [config]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200 import ops, synth  # noqa: E402
from graphtrans_b200._lib import call, dt_of, ptr  # noqa: E402
from tools.mha_bench import graph_time  # noqa: E402

if __name__ == "__main__":
    for cfg in (sys.argv[1:] or ["molpcba"]):
        args = synth.make_args(cfg)
        hb = synth.make_batch(args, B=args.batch_size, seed=0)
        b = hb.to("cuda")
        plan = ops.GraphPlan(b.edge_index, b.batch, b.num_graphs, int(args.max_input_len), max_nodes=hb.max_nodes)
        nhead, d = args.nhead, args.d_model
        dh, n = d // nhead, plan.n_rows
        qkv = torch.randn(n, 3 * d, device="cuda").bfloat16()
        dout = torch.randn(n, d, device="cuda").bfloat16()
        out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(nhead * n, device="cuda")
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(nhead * n, device="cuda")
        rng = ops.rng_state("cuda")
        print(f"{cfg}: rows={n} tiles used {int(plan.loc_count)} of {plan.loc_max_tiles} slots, dh={dh}")
        for p in (0.0, 0.3):
            def lf():
                call("gt_mha_local_fwd", dt_of(qkv), ptr(qkv), ptr(plan.row_bounds), ptr(plan.loc_tiles), ptr(plan.loc_count), plan.loc_max_tiles, n,
                     nhead, dh, dh ** -0.5, ptr(out), ptr(lse), p, ptr(rng) if p else None, 5)

            def lb():
                call("gt_mha_local_bwd", dt_of(qkv), ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(plan.row_bounds), ptr(plan.loc_tiles),
                     plan.loc_max_tiles, n, nhead, dh, dh ** -0.5, ptr(dqkv), p, ptr(rng) if p else None, 5)

            def sf():
                call("gt_mha_fwd", dt_of(qkv), ptr(qkv), ptr(plan.tok_graph), ptr(plan.tok_off), None, ptr(plan.row_bounds),
                     ptr(plan.tile_bounds), n, plan.B, nhead, dh, dh ** -0.5, ptr(out), ptr(lse), p, ptr(rng) if p else None, 5, 2)

            def sb():
                call("gt_mha_bwd", dt_of(qkv), ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(plan.tok_graph), ptr(plan.tok_off), None,
                     ptr(plan.row_bounds), ptr(plan.tile_bounds), n, plan.B, nhead, dh, dh ** -0.5, ptr(dqkv), ptr(delta), p,
                     ptr(rng) if p else None, 5, 2)

            print(f"  p={p}: local fwd {graph_time(lf):6.1f} bwd {graph_time(lb):6.1f} us | streamed fwd {graph_time(sf):6.1f} bwd {graph_time(sb):6.1f} us", flush=True)
        if os.environ.get("GT_LOC_TRACE"):
            import ctypes
            from graphtrans_b200 import _lib
            p = 0.0
            lf()
            torch.cuda.synchronize()
            buf = (ctypes.c_ulonglong * 16)()
            _lib.load().gtdbg_loc_trace_read(buf)
            t = list(buf)
            names = ["start", "alloc+sync", "ld_full", "s_full", "max done", "p_full", "o_full", "stored", "dealloc"]
            print("  CTA(0,0) phases (ns since start):", {n: t[i] - t[0] for i, n in enumerate(names)})
