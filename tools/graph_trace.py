"""In-situ timeline of one CUDA-graph replay of the training step: a one-thread %globaltimer stamp kernel is captured
after every C-ABI call (graphtrans_b200._lib.start_stamps), on the stream the call was issued on, so the replayed
graph reports where the step time goes with warm caches and the real overlap between the main stream, the
weight-gradient stream and the virtual-node branch.  Interval i = stamp[i] - (previous stamp on the same stream, or the
stream's first dependency): kernel duration + dependent-launch gap + the stamp kernel itself (~1.5-2 us, calibrated
from two back-to-back stamps).
python tools/graph_trace.py [config] [--raw]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphtrans_b200 import _lib, factory, ops, synth  # noqa: E402
from graphtrans_b200.ddp import GradBuckets  # noqa: E402
from graphtrans_b200.graphed import GraphedStep  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "molpcba"
raw = "--raw" in sys.argv
ops.set_precision("bf16")
args = synth.make_args(cfg)
if cfg == "code2-pna":
    args.deg = synth.in_degree_histogram(synth.make_batch(args, B=args.batch_size, seed=1234), 800)
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = factory.build_model(args).to(dev).train()
lossf = factory.loss_fn(args)
buckets = GradBuckets(model, n_buckets=4, overlap=False)
if "--no-side" not in sys.argv:
    ops.enable_wgrad_stream(True, dev)
    ops.enable_branch_stream(True, dev)
b = synth.make_batch(args, B=args.batch_size, seed=0).to(dev)

# un-stamped graph first: the reference step time
g0 = GraphedStep(model, lossf, buckets)
for _ in range(3):
    g0(b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g0(b)
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / 10

g = GraphedStep(model, lossf, buckets, warmup_iters=2)
orig_eager = g._eager


def eager_stamped(bb):
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing:
        _lib.start_stamps(device=dev)
        _lib.stamp("<begin>")
        _lib.stamp("<begin2>")
    out = orig_eager(bb)
    if capturing:
        _lib.stamp("<end>")
        eager_stamped.rec, eager_stamped.buf = _lib.stop_stamps()
    return out


g._eager = eager_stamped
for _ in range(4):
    g(b)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    g(b)
e1.record()
torch.cuda.synchronize()
stamped_ms = e0.elapsed_time(e1) / 10
rec, buf = eager_stamped.rec, eager_stamped.buf.cpu().tolist()
t0 = buf[0]
cal = (buf[1] - buf[0]) / 1e3
main = rec[0][1]
streams = {}
for name, st, _ in rec:
    streams.setdefault(st, len(streams))
print(f"{cfg}: plain graph {plain_ms * 1e3:.1f} us/step, stamped graph {stamped_ms * 1e3:.1f} us/step, "
      f"{len(rec) - 3} calls, stamp-to-stamp {cal:.2f} us, streams {len(streams)}; begin->end {(buf[len(rec) - 1] - t0) / 1e3:.1f} us")
last = {}
per = collections.defaultdict(lambda: [0, 0.0])
rows = []
prev_any = t0
for i, (name, st, a) in enumerate(rec):
    t = buf[i]
    s = streams[st]
    if s in last:
        base = last[s]
    else:
        base = prev_any          # first call on a side stream: measured from the latest stamp before it (its fork point)
    dt = (t - base) / 1e3
    key = name
    if name == "gt_gemm":
        key = f"gt_gemm a_mn={a[2]} b_mn={a[5]} M={a[9]} N={a[10]} K={a[11]}"
    elif name in ("gt_bn_norm_fwd", "gt_colstats", "gt_bn_bwd_reduce", "gt_bn_bwd_apply", "gt_colsum",
                  "gt_layernorm_fwd", "gt_layernorm_bwd", "gt_cast_pad", "gt_segment_sum_sorted"):
        key = f"{name} {[x for x in a if isinstance(x, int) and 1 < x < 10**7][:3]}"
    key = f"s{s} {key}"
    per[key][0] += 1
    per[key][1] += dt
    rows.append((i, s, (t - t0) / 1e3, dt, key))
    if s in last or True:
        last[s] = t
    prev_any = max(prev_any, t) if s == 0 else prev_any
tot = collections.defaultdict(float)
for k, v in per.items():
    tot[k.split()[0]] += v[1]
print("sum of intervals per stream:", {k: round(v, 1) for k, v in tot.items()})
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us  n={v[0]:3d}  avg {v[1] / v[0]:6.1f}  {k}")
if raw:
    print("--- timeline (idx, stream, t_end us, interval us, call)")
    for r in rows:
        print(f"{r[0]:4d} s{r[1]} {r[2]:9.1f} {r[3]:7.1f}  {r[4]}")
