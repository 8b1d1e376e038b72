"""2-rank GPU worker of tests/test_ddp_gpu.py (SURVEY §4 "distributed" tier), launched with torchrun:
  (1) rank-averaged gradients of the data-parallel step == mean of the per-shard fp64 ORACLE gradients (fp32 mode,
      the <= 1e-3 contract), with the bucketed ncclAllReduce captured INSIDE the CUDA graph of the step
      (GradBuckets(direct=True, overlap=True) + GraphedStep) and, separately, through torch.distributed after the replay;
  (2) eager overlap mode with the weight-gradient / virtual-node side streams enabled == overlap off (the
      reduction of a bucket has to wait for gradients still queued on the side streams);
  (3) the fused AdamW inside the multi-rank graph keeps the ranks' weights identical.
Prints DDP_GPU_OK on rank 0 when everything holds."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from graphtrans_b200 import factory, loader, ops, synth  # noqa: E402
from graphtrans_b200.ddp import GradBuckets, shard_range  # noqa: E402
from graphtrans_b200.graphed import GraphedStep  # noqa: E402
from graphtrans_b200.optim import FusedAdamW  # noqa: E402


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp(min=1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ops.set_precision("fp32")
    for cfg in ("molpcba", "code2"):
        kw = dict(gnn_dropout=0.0, transformer_dropout=0.0, gnn_emb_dim=64, d_model=64, dim_feedforward=128)
        if cfg == "code2":
            kw["num_tasks"] = 100
        args = synth.make_args(cfg, **kw)
        Bg = 16 * world
        full = synth.make_batch(args, B=Bg, seed=5)
        if cfg == "code2":
            full.y_arr = full.y_arr % args.num_tasks
        lo, hi = shard_range(Bg, rank, world)
        mine = loader.shard(full, lo, hi)
        torch.manual_seed(0)
        model = factory.build_model(args).to(dev).train()
        init = copy.deepcopy(model.state_dict())
        lossf = factory.loss_fn(args)

        # ---- (1) in-graph NCCL reduction
        buckets = GradBuckets(model, n_buckets=3, overlap=True, direct=True)
        step = GraphedStep(model, lossf, buckets)
        step(mine.to(dev))
        torch.cuda.synchronize()
        g_graph = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        step(mine.to(dev))                               # a replay gives the same averaged gradients
        torch.cuda.synchronize()
        bad = [(k, rel(p.grad, g_graph[k])) for k, p in model.named_parameters()
               if rel(p.grad, g_graph[k]) > 1e-5 and float(g_graph[k].abs().max()) > 1e-7]
        if bad:
            print(f"[rank {rank}] [{cfg}] replay differs from the first replay in {len(bad)} tensors:", bad[:6], flush=True)
        assert not bad, "replay"

        # ---- torch.distributed reduction after the replay (round-1 path)
        m2 = factory.build_model(args).to(dev).train()
        m2.load_state_dict(init)
        b2 = GradBuckets(m2, n_buckets=3, overlap=False)
        GraphedStep(m2, lossf, b2)(mine.to(dev))
        torch.cuda.synchronize()
        assert rel(b2.flat, buckets.flat) < 1e-5, "in-graph NCCL vs torch.distributed allreduce"

        # ---- (2) eager overlap with side streams vs no overlap
        ops.enable_wgrad_stream(True, dev)
        ops.enable_branch_stream(True, dev)
        flats = []
        for ov in (True, False):
            m3 = factory.build_model(args).to(dev).train()
            m3.load_state_dict(init)
            b3 = GradBuckets(m3, n_buckets=3, overlap=ov)
            for _ in range(3):
                b3.zero_grad()
                bb = mine.to(dev)
                loss = lossf(m3(bb), bb)
                loss.backward()
                ops.join_side_streams()
                b3.finish()
            torch.cuda.synchronize()
            flats.append(b3.flat.clone())
        ops.enable_wgrad_stream(False)
        ops.enable_branch_stream(False)
        assert rel(flats[0], flats[1]) < 1e-5, "eager overlap with side streams vs overlap off"
        assert rel(flats[0], buckets.flat) < 1e-4, "eager vs graphed"

        # ---- oracle: mean over ranks of the per-shard fp64 gradients
        if rank == 0:
            from oracle import graphtrans_oracle as O
            acc = None
            for r in range(world):
                a, b = shard_range(Bg, r, world)
                _, _, og, _ = O.fwd_bwd({k: v.cpu() for k, v in init.items()}, args, loader.shard(full, a, b), dtype=torch.float64)
                acc = og if acc is None else {k: acc[k] + v for k, v in og.items()}
            num = den = 0.0
            for k, v in acc.items():
                v = v / world
                num += (g_graph[k].double().cpu() - v).pow(2).sum().item()
                den += v.pow(2).sum().item()
            err = (num / den) ** 0.5
            print(f"[{cfg}] averaged gradients vs mean of per-shard oracle gradients: rel-L2 {err:.3e}", flush=True)
            assert err < 1e-3, err

        # ---- (3) optimizer inside the multi-rank graph: ranks stay bit-identical
        m4 = factory.build_model(args).to(dev).train()
        m4.load_state_dict(init)
        b4 = GradBuckets(m4, n_buckets=3, overlap=True, direct=True)
        opt = FusedAdamW(b4, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
        s4 = GraphedStep(m4, lossf, b4, optimizer=opt)
        assert s4.opt_in_graph and s4.comm_in_graph
        for _ in range(3):
            s4(mine.to(dev))
        torch.cuda.synchronize()
        w = torch.cat([p.detach().flatten() for p in m4.parameters()])
        w0 = w.clone()
        dist.broadcast(w0, src=0)
        assert torch.equal(w, w0), "weights diverged across ranks"
        assert rel(w, torch.cat([p.detach().flatten() for p in model.parameters()])) > 1e-6     # they did move
        dist.barrier()
    if rank == 0:
        print("DDP_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
