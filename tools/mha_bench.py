"""Forward / backward timing of the attention kernels on the packed token layout of the BASELINE batches (CUDA graph of
20 back-to-back launches, warm L2): impl 2 = tcgen05 tiles, 1 = CUDA-core reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200 import ops, synth  # noqa: E402
from graphtrans_b200._lib import call, dt_of, ptr  # noqa: E402


def graph_time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
  for cfg in (sys.argv[1:] or ["molpcba", "nci1", "code2"]):
      args = synth.make_args(cfg)
      b = synth.make_batch(args, B=args.batch_size, seed=0).to("cuda")
      plan = ops.GraphPlan(b.edge_index, b.batch, b.num_graphs, int(args.max_input_len))
      nhead, d = args.nhead, args.d_model
      dh, n = d // nhead, plan.n_rows
      qkv = torch.randn(n, 3 * d, device="cuda").bfloat16()
      dout = torch.randn(n, d, device="cuda").bfloat16()
      out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
      lse = torch.empty(nhead * n, device="cuda")
      dqkv = torch.empty_like(qkv)
      delta = torch.empty(nhead * n, device="cuda")
      rng = ops.rng_state("cuda")
      for p in (0.0, 0.3):
          for impl in (2, 1):

              def fwd():
                  call("gt_mha_fwd", dt_of(qkv), ptr(qkv), ptr(plan.tok_graph), ptr(plan.tok_off), None, ptr(plan.row_bounds),
                       ptr(plan.tile_bounds), n, plan.B, nhead, dh, dh ** -0.5, ptr(out), ptr(lse), p, ptr(rng) if p else None, 5, impl)

              def bwd():
                  call("gt_mha_bwd", dt_of(qkv), ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(plan.tok_graph), ptr(plan.tok_off), None,
                       ptr(plan.row_bounds), ptr(plan.tile_bounds), n, plan.B, nhead, dh, dh ** -0.5, ptr(dqkv), ptr(delta), p,
                       ptr(rng) if p else None, 5, impl)

              print(f"{cfg:8s} rows={n} B={plan.B} dh={dh} p={p} impl={impl}: fwd {graph_time(fwd):7.1f} us  bwd {graph_time(bwd):7.1f} us", flush=True)


if __name__ == "__main__":
    main()
