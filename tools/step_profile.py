"""In-situ (L2-warm, real operands) duration of every C-ABI call of one eager training step: CUDA events around each
launch (graphtrans_b200._lib.start_profile).  The eager host is slower than the GPU, so every launch starts on an idle
device: numbers are kernel duration + ~1-2 us of launch latency, without queueing effects.
python tools/step_profile.py [config]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200 import _lib, factory, ops, synth  # noqa: E402
from graphtrans_b200.ddp import GradBuckets  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "molpcba"
ops.set_precision("bf16")
args = synth.make_args(cfg)
if cfg == "code2-pna":
    args.deg = synth.in_degree_histogram(synth.make_batch(args, B=args.batch_size, seed=1234), 800)
torch.manual_seed(0)
model = factory.build_model(args).cuda().train()
lossf = factory.loss_fn(args)
buckets = GradBuckets(model, n_buckets=4, overlap=False)
b = synth.make_batch(args, B=args.batch_size, seed=0).to("cuda")


def step():
    buckets.zero_grad()
    loss = lossf(model(b), b)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.start_profile()
R = 5
for _ in range(R):
    step()
rec = _lib.stop_profile()
per = collections.defaultdict(lambda: [0, 0.0])
for name, ms, a in rec:
    key = name
    if name == "gt_gemm":
        key = f"gt_gemm a_mn={a[2]} b_mn={a[5]} M={a[9]} N={a[10]} K={a[11]}"
    elif name in ("gt_bn_norm_fwd", "gt_colstats", "gt_bn_bwd_reduce", "gt_bn_bwd_apply", "gt_colsum", "gt_layernorm_fwd", "gt_layernorm_bwd"):
        key = f"{name} M={[x for x in a if isinstance(x, int) and 1 < x < 10**7][:3]}"
    per[key][0] += 1
    per[key][1] += ms * 1e3
tot = sum(v[1] for v in per.values()) / R
print(f"{cfg}: {len(rec) // R} C-ABI calls per step, sum of per-call event times {tot:.1f} us per step")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / R:9.1f} us/step  n={v[0] // R:3d}  avg {v[1] / v[0]:6.1f}  {k}")
