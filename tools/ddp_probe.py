import copy, os, sys
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
from graphtrans_b200 import factory, loader, ops, synth
from graphtrans_b200.ddp import GradBuckets, shard_range
from graphtrans_b200.graphed import GraphedStep
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ops.set_precision("fp32")
kw = dict(gnn_dropout=0.0, transformer_dropout=0.0, gnn_emb_dim=64, d_model=64, dim_feedforward=128)
args = synth.make_args("molpcba", **kw)
full = synth.make_batch(args, B=32, seed=5)
lo, hi = shard_range(32, rank, world)
mine = loader.shard(full, lo, hi)
torch.manual_seed(0)
base = factory.build_model(args).to(dev).train()
init = copy.deepcopy(base.state_dict())
lossf = factory.loss_fn(args)
def rel(a,b): return float((a.double()-b.double()).norm()/b.double().norm().clamp(min=1e-30))
def mk():
    m = factory.build_model(args).to(dev).train(); m.load_state_dict(init); return m
# reference: eager, host allreduce, no overlap
m = mk(); b = GradBuckets(m, n_buckets=3, overlap=False)
b.zero_grad(); bb = mine.to(dev); l = lossf(m(bb), bb); l.backward(); b.finish(); torch.cuda.synchronize()
ref = b.flat.clone(); names = [k for k,_ in m.named_parameters()]; offs = b.offsets; nums=[p.numel() for p in m.parameters()]
del l
def report(tag, flat):
    bad = []
    for n,o,c in zip(names, offs, nums):
        r = rel(flat[o:o+c], ref[o:o+c])
        if r > 1e-4 and float(ref[o:o+c].abs().max()) > 1e-7: bad.append((n, round(r,3)))
    print(f"[rank {rank}] {tag}: total rel {rel(flat, ref):.2e} bad {len(bad)} {bad[:5]}", flush=True)
for ov, direct, tag in ((False, True, "graph, in-graph allreduce at the end"), (True, True, "graph, in-graph overlapped")):
    m = mk(); b = GradBuckets(m, n_buckets=3, overlap=ov, direct=direct); st = GraphedStep(m, lossf, b)
    for it in range(3):
        st(mine.to(dev)); torch.cuda.synchronize(); report(f"{tag} replay {it}", b.flat)
dist.destroy_process_group()
