"""bf16 throughput mode: RECORDED error budget.

For BASELINE configs 2 / 3 / 4 / 5 at their full widths the product in bf16 (the mode bench.py times: bf16 activations,
tcgen05 contractions, tile-local / streamed tcgen05 attention, bf16 aggregation) is compared with the fp64 CPU oracle,
next to the oracle itself under torch.autocast(bfloat16) - the precision the reference would run at with AMP (the
reference under bf16 autocast is 1.4e-2 / 0.25 off its fp64 run).  The product has to stay within 1.5x
of that reference-autocast error (plus a small absolute floor); the numbers are written to
gpurun_out/r02_bf16_error.json (copied to profiles/ by the builder).  Dropout 0 (RNG streams cannot match)."""
import copy
import json
import os

import pytest
import torch

from graphtrans_b200 import factory, ops, synth
from oracle import graphtrans_oracle as O
from tests.helpers import as_list, grad_report, rel_l2

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# graphs per batch: config 2 at the benched B; the others reduced so the fp64 oracle stays within ~a minute of CPU time
CASES = [("molpcba", 512), ("code2", 24), ("syn", 96), ("code2-pna", 24)]
_results = {}


def _logits_err(pred, ref):
    num = sum((a.double().cpu() - b.double()).pow(2).sum().item() for a, b in zip(as_list(pred), as_list(ref)))
    den = sum(b.double().pow(2).sum().item() for b in as_list(ref))
    return (num / den) ** 0.5


@pytest.mark.parametrize("cfg,B", CASES)
def test_bf16_error_budget(cfg, B):
    kw = dict(gnn_dropout=0.0, transformer_dropout=0.0)
    if cfg in ("code2", "code2-pna"):
        kw["num_tasks"] = 1000            # 5 heads x 1000 classes keep the fp64 oracle affordable; widths are the config's
    args = synth.make_args(cfg, **kw)
    batch = synth.make_batch(args, B=B, seed=11)
    if args.dataset == "code2":
        batch.y_arr = batch.y_arr % args.num_tasks
    if cfg == "code2-pna":
        args.deg = synth.in_degree_histogram(batch, 800)
    torch.manual_seed(0)
    model = factory.build_model(args)
    init = copy.deepcopy(model.state_dict())
    # fp64 oracle = ground truth
    opred, oloss, ograds, _ = O.fwd_bwd(init, args, batch, dtype=torch.float64)
    # the oracle under bf16 autocast = what the reference's PyTorch path gives with AMP
    with torch.autocast("cpu", dtype=torch.bfloat16):
        apred, aloss, agrads, _ = O.fwd_bwd(init, args, batch, dtype=torch.float32)
    a_logits = _logits_err(apred, opred)
    a_glob, _, _ = grad_report({k: v.float() for k, v in agrads.items()}, ograds)
    # product, bf16 mode (exactly the kernels bench.py times)
    ops.set_precision("bf16")
    try:
        model = model.cuda().train()
        b = batch.clone().to("cuda")
        pred = model(b)
        loss = factory.loss_fn(args)(pred, b)
        loss.backward()
        torch.cuda.synchronize()
        grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    finally:
        ops.set_precision("fp32")
    p_logits = _logits_err(pred, opred)
    p_glob, p_worst, p_key = grad_report(grads, ograds)
    _results[cfg] = {"graphs": B, "product_bf16": {"logits_rel_l2": p_logits, "grads_rel_l2": p_glob, "worst_param": p_key,
                                                   "worst_param_err": p_worst, "loss": float(loss.detach())},
                     "reference_autocast_bf16": {"logits_rel_l2": a_logits, "grads_rel_l2": a_glob, "loss": float(aloss)},
                     "oracle_fp64_loss": float(oloss)}
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "r02_bf16_error.json"), "w") as f:
        json.dump(_results, f, indent=1)
    assert p_logits <= 1.5 * a_logits + 2e-3, (p_logits, a_logits)
    assert p_glob <= 1.5 * a_glob + 1e-2, (p_glob, a_glob, p_key)
