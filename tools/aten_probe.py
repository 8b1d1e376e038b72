"""Which torch (ATen) kernels are still launched inside one training step?  Runs one eager step under torch.profiler and
lists every aten op that launched a CUDA kernel, with counts and input shapes.
python tools/aten_probe.py [config]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from graphtrans_b200 import factory, loader, ops, synth  # noqa: E402
from graphtrans_b200.ddp import GradBuckets  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "molpcba"
ops.set_precision("bf16")
args = synth.make_args(cfg)
if cfg == "code2-pna":
    args.deg = synth.in_degree_histogram(synth.make_batch(args, B=args.batch_size, seed=1234), 800)
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = factory.build_model(args).to(dev).train()
lossf = factory.loss_fn(args)
buckets = GradBuckets(model, n_buckets=4, overlap=False)
ops.enable_wgrad_stream(True, dev)
ops.enable_branch_stream(True, dev)
b = loader.prepare(synth.make_batch(args, B=args.batch_size, seed=0)).to(dev)


def step():
    ops.begin_step(dev)
    buckets.zero_grad()
    loss = lossf(model(b), b)
    loss.backward()
    ops.join_side_streams()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=4):
    if e.key.startswith("aten::") and getattr(e, "device_time_total", 0) > 0 and e.count > 0:
        rows.append((e.count, e.key, str(e.input_shapes)[:90], e.device_time_total, [s for s in e.stack if "graphtrans_b200" in s or "bench" in s][:2]))
rows.sort(key=lambda r: -r[0])
for r in rows:
    print("%3d  %-28s %-92s %8.1f us  %s" % r)
