"""Shared test helpers (fixtures loading, error metrics)."""
import argparse
import glob
import os

import torch

from graphtrans_b200.synth import GraphBatch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt"))
                      if not os.path.basename(p).startswith("_"))


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    fx["args"] = argparse.Namespace(**fx["args"])
    fx["batch"] = GraphBatch(**fx["batch"])
    return fx


def rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / den if den > 0 else (a - b).norm().item()


def as_list(p):
    return list(p) if isinstance(p, (list, tuple)) else [p]


def grad_report(grads, ref_grads):
    """global rel-L2 and the worst per-parameter error measured against max(|g_p|, 1e-3*|g|_global)
    (biases feeding train-mode BN have an exactly-zero true gradient)."""
    num = den = 0.0
    for k, g in ref_grads.items():
        num += (grads[k].double().cpu() - g.double()).pow(2).sum().item()
        den += g.double().pow(2).sum().item()
    gnorm = den ** 0.5
    worst, worst_key = 0.0, None
    for k, g in ref_grads.items():
        e = (grads[k].double().cpu() - g.double()).norm().item() / max(g.double().norm().item(), 1e-3 * gnorm)
        if e > worst:
            worst, worst_key = e, k
    return (num ** 0.5) / max(gnorm, 1e-30), worst, worst_key
