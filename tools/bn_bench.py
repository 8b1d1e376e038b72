"""Micro-benchmark of the BatchNorm kernels (gt_colstats / gt_bn_norm_fwd / gt_bn_bwd_reduce / gt_bn_bwd_apply) on the
shapes of config 2, L2-warm (back-to-back launches on the same 8-16 MB operands, as inside the training step) and
L2-cold (256 MiB flush between launches).  Knobs: GT_BN_SLAB (row-slab normalise kernel on/off), GT_BN_BPS, GT_BN_MINROWS (csrc/rowops.cu slab_cfg)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphtrans_b200 import ops  # noqa: E402


def timed(fn, reps=30, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(20e-3 * 1.9e9))
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def main():
    ops.set_precision("bf16")
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cfgs = [dict(GT_BN_SLAB="0"), dict(GT_BN_SLAB="1", GT_BN_BPS="1"), dict(GT_BN_SLAB="1", GT_BN_BPS="2"),
            dict(GT_BN_SLAB="1", GT_BN_BPS="3")]
    for M, d in ((13510, 600), (13510, 300), (512, 600), (512, 300)):
        ld = ops.ldp(d)
        bn = torch.nn.BatchNorm1d(d).to(dev).train()
        x = torch.randn(M, ld, device=dev).bfloat16()
        g = torch.randn(M, ld, device=dev).bfloat16()
        for cfg in cfgs:
            for k in ("GT_BN_SLAB", "GT_BN_BPS", "GT_BN_MINROWS"):
                os.environ.pop(k, None)
            os.environ.update(cfg)
            xr = x.clone().requires_grad_(True)

            def fwd():
                return ops.batch_norm(xr, bn, relu=True, drop_p=0.3)

            y = fwd()

            def bwd():
                torch.autograd.grad(y, xr, g, retain_graph=True)

            prof = {}
            for name, fn in (("fwd(colstats+norm)", fwd), ("bwd(reduce+apply)", bwd)):
                prof[name] = (timed(fn), timed(fn, flush=flush))
            print(f"M={M:6d} d={d:4d} {str(cfg):48s} " + "  ".join(f"{k}: warm {v[0]:6.1f} us cold {v[1]:6.1f} us" for k, v in prof.items()), flush=True)


if __name__ == "__main__":
    main()
