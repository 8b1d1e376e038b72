#!/usr/bin/env python
"""graphs/sec forward+backward of the GraphTrans hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config syn|molpcba|code2|code2-pna|nci1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores
    python bench.py --impl torch-gpu ...      # the same restatement on the B200 through stock torch kernels (comparator)

Headline workload = BASELINE.json configs[3] (the one its scaling sweep is defined on): synthetic graphs n~U{64..192},
4 GIN + 4 Tx layers, d = 256, GLOBAL batch 4096, STRONG scaling over N GPUs (4096 / N graphs per rank).  Configs 2, 3 and
5 (molpcba B=512, Code2 B=128, Code2-PNA B=128 per GPU) are timed in the same run and carried under `configs`.

One "step" = zero_grad -> forward -> loss -> backward (+ bucketed NCCL gradient allreduce, captured inside the CUDA graph
of the step, when N > 1) over one synthetic batch.  Rank 0 prints ONE JSON line.
  value : whole-job graphs/s with the batches already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API (GraphedStep over loader.prepare'd HOST batches) on batches the step has
          NEVER seen: the H2D copy of each batch (one pinned blob, prefetched on a copy stream while the previous step
          runs) and a D2H read of the loss inside the timed region
  roofline     : the hot-stage kernel BASELINE.json quotes the config on, timed IN the captured step (device timestamps
                 after every kernel of the replayed graph, graphtrans_b200/trace.py): achieved = algorithmic bytes|flops
                  / in-step time; peak from MEASURED_PEAKS.json.  `alone` = the same kernels replayed back
                 to back outside the step.
  cpu_baseline : the CPU oracle (port of the reference's PyTorch path) timed on the host cores
"""
import argparse
import gc
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from graphtrans_b200 import synth  # noqa: E402

METRIC = "graphs/sec fwd+bwd"
UNIT = "graphs/s"
HEADLINE = "syn"
EXTRA_CONFIGS = ("molpcba", "code2", "code2-pna")
# graphs per CPU step: the full batch where a step costs ~1 s of CPU time, a bounded sample otherwise
CPU_SAMPLE_B = {"molpcba": 512, "code2": 16, "syn": 256, "code2-pna": 16, "nci1": 32}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch-gpu"])
    ap.add_argument("--config", default=HEADLINE, choices=list(synth.CONFIGS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=None, help="graphs per step (global with --scaling strong, per GPU with weak)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="default: strong for syn (BASELINE's sweep), weak otherwise")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true", help="skip the second figure (step + fused AdamW update)")
    ap.add_argument("--no-extra-configs", action="store_true", help="only the headline workload (no `configs` entries)")
    ap.add_argument("--distinct-batches", type=int, default=4)
    ap.add_argument("--no-wgrad-stream", action="store_true", help="keep weight gradients on the main stream")
    ap.add_argument("--no-branch-stream", action="store_true", help="keep the virtual-node branch on the main stream")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph replay: launch every kernel from Python")
    ap.add_argument("--comm", default="graph", choices=["graph", "host"],
                    help="N > 1: gradient allreduce captured inside the step's CUDA graph (default) or issued from the host after the replay")
    ns = ap.parse_args()
    if ns.scaling is None:
        ns.scaling = "strong" if ns.config == "syn" else "weak"
    return ns


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tensor=1590.0, source="fallback")


def config_args(cfg):
    args = synth.make_args(cfg)
    if cfg == "code2-pna":
        probe = synth.make_batch(args, B=args.batch_size, seed=1234)
        args.deg = synth.in_degree_histogram(probe, 800)
    return args


def workload_name(cfg, args, B_global, scaling, world):
    arch = f"{args.gnn_type.upper() if args.model_type == 'gnn-transformer' else 'PNA'}{'-Virtual' if args.gnn_virtual_node else ''}"
    base = (f"{cfg}: GraphTrans {arch} JK={args.gnn_JK} d_g={args.gnn_emb_dim} L_g={args.gnn_num_layer} d={args.d_model} "
            f"L_t={args.num_encoder_layers}")
    if scaling == "strong":
        return base + f" global B={B_global} (strong scaling: B/N graphs per GPU)"
    return base + f" B={B_global // max(world, 1)}/GPU"


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples SM clock + throttle reasons of this process's GPU every 100 ms via NVML"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10, "app_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------- CPU arm / torch-gpu arm
def oracle_run(cfg, B_full, steps, warmup, device="cpu", autocast=False):
    """Times the oracle (restatement of the reference's PyTorch/PyG path) on a bounded sample of the workload, on the
    host cores (device='cpu', fp32, all threads) or on the GPU through stock torch kernels (device='cuda').
    -> (graphs/s, ms/step, description dict)"""
    from graphtrans_b200 import factory
    from oracle import graphtrans_oracle as O
    args = config_args(cfg)
    args.gnn_dropout = 0.0           # the oracle has no dropout (identity); the product runs the configured p
    args.transformer_dropout = 0.0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = min(CPU_SAMPLE_B[cfg], B_full) if device == "cpu" else B_full
    batch = synth.make_batch(args, B=Bs, seed=0)
    torch.manual_seed(0)
    sd = factory.build_model(args).state_dict()
    if device != "cpu":
        sd = {k: v.to(device) for k, v in sd.items()}
        batch = batch.to(device)

    def one():
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                O.fwd_bwd(sd, args, batch, dtype=torch.float32)
        else:
            O.fwd_bwd(sd, args, batch, dtype=torch.float32)
        if device != "cpu":
            torch.cuda.synchronize()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    sample = (f"{Bs} of {B_full} graphs/step ({cfg} shape, seed 0), fp32, dropout 0, {warmup} warm-up + {steps} timed "
              f"fwd+bwd steps of oracle/graphtrans_oracle.py")
    return Bs / dt, dt * 1e3, dict(cores=cores, threads=torch.get_num_threads(), kind="port", sample=sample, graphs=Bs)


def main_reference(ns):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args = config_args(ns.config)
    B_global = ns.batch or args.batch_size * (1 if ns.scaling == "strong" else max(ns.gpus, 1))
    # bound the arm to a few minutes whatever K the driver passes: cap the timed CPU steps
    est_steps = min(ns.steps, 5 if CPU_SAMPLE_B[ns.config] >= 256 else 10)
    gps, ms, desc = oracle_run(ns.config, B_global, est_steps, min(ns.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": gps, "unit": UNIT, "n_gpus": ns.gpus, "steps": ns.steps,
        "warmup": ns.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": ns.scaling, "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload_name(ns.config, args, B_global, ns.scaling, ns.gpus), "timed_steps": est_steps,
                   "note": "CPU path of the reference (oracle port; the reference is pure Python and its third-party deps are "
                           "not installable here, so there is no oracle/_ref); each step = a bounded sample of the workload, "
                           "graphs/s = sample graphs / step time"},
        "cpu_baseline": {"value": gps, "unit": UNIT, "cores": desc["cores"], "kind": desc["kind"], "sample": desc["sample"]},
        "e2e": {"value": gps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main_torch_gpu(ns):
    """the GPU comparator: the same restatement of the reference's path executed on the B200 by stock
    torch kernels (cuBLAS GEMMs, ATen scatter / softmax / norms), fp32 and under bf16 autocast"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    assert torch.cuda.is_available()
    out = {}
    for cfg in ([ns.config] if ns.no_extra_configs else [ns.config] + [c for c in EXTRA_CONFIGS if c != ns.config]):
        args = config_args(cfg)
        B = ns.batch if (ns.batch and cfg == ns.config) else args.batch_size
        if cfg == "syn":
            B = min(B, 1024)          # the padded [T, B, d] attention of the restatement materialises B*h*T*T scores
        res = {}
        for name, ac in (("fp32", False), ("bf16_autocast", True)):
            try:
                gps, ms, desc = oracle_run(cfg, B, min(ns.steps, 10), 2, device="cuda", autocast=ac)
                res[name] = {"value": gps, "ms_per_step": ms, "graphs_per_step": desc["graphs"]}
            except Exception as e:  # noqa: BLE001
                res[name] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
        out[cfg] = res
    head = out[ns.config].get("bf16_autocast", {})
    args = config_args(ns.config)
    line = {"impl": "torch-gpu", "metric": METRIC, "value": head.get("value"), "unit": UNIT, "n_gpus": 1, "steps": ns.steps,
            "warmup": ns.warmup, "ms_per_step": head.get("ms_per_step"), "higher_is_better": True, "scaling": ns.scaling,
            "vs_baseline": None, "dtype": "bf16 autocast (fp32 beside it)", "data": "synthetic",
            "config": {"workload": workload_name(ns.config, args, ns.batch or args.batch_size, ns.scaling, 1),
                       "note": "oracle port (plain torch ops) on the B200: stock cuBLAS / ATen kernels, eager launches, dropout 0; "
                               "syn runs 1024 graphs per step (memory of the materialised attention scores)"},
            "configs": out, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- GPU arm
def algorithmic(args, batch, es):
    """per-step algorithmic work of the two hot stages for one batch"""
    N = batch.batch.numel()
    E = batch.edge_index.shape[1]
    ea = 0 if batch.edge_attr is None else batch.edge_attr.numel() * batch.edge_attr.element_size()
    d_g = args.gnn_emb_dim
    n = torch.bincount(batch.batch)
    t = n.clamp(max=int(args.max_input_len)) + (1 if args.graph_pooling == "cls" else 0)
    sum_t2, sum_t = float((t.double() ** 2).sum()), float(t.sum())
    if args.model_type == "pna-transformer":
        agg_bytes = N * d_g * es * 3 + 13 * N * d_g * es + 16 * E      # x, pi, pj read; 13F x towers written
    else:
        agg_bytes = 2 * N * d_g * es + 16 * E + ea
    d = args.d_model
    return dict(N=N, E=E, agg_bytes_per_launch=agg_bytes, mha_fwd_flops=4.0 * d * sum_t2, mha_bwd_flops=8.0 * d * sum_t2,
                mha_pooled_fwd_flops=4.0 * d * sum_t, mha_pooled_bwd_flops=8.0 * d * sum_t, tokens=int(sum_t))


def _timed(fn, reps=20):
    """average GPU milliseconds of fn() launched `reps` times back to back (head start hides launch latency)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(30e-3 * 1.9e9))
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def alone_replays(args, model, b, hb, precision):
    """the hot-stage kernels replayed back to back OUTSIDE the step, on this batch's real plan, running the SAME variants
    the step runs (trainable edge table: the adjoint writes the per-edge gradient `gm`; the one-hot table-gradient
    contraction is timed separately).  -> {name: ms per (fwd + bwd) pair}"""
    from graphtrans_b200 import ops
    from graphtrans_b200._lib import CONV_GCN, CONV_GIN
    from graphtrans_b200.modules import conv as conv_mod
    act = ops.act_dtype()
    plan = ops.plan_for(b, int(args.max_input_len), cls=args.graph_pooling == "cls")
    N, d_g, ld = b.batch.numel(), args.gnn_emb_dim, ops.ldp(args.gnn_emb_dim)
    dev = b.batch.device
    out = {}
    torch.manual_seed(0)
    if args.model_type == "gnn-transformer":
        conv = model.gnn_node.convs[1]
        x = (torch.randn(N, ld, device=dev) * 0.5).to(act)
        x[:, d_g:] = 0
        x.requires_grad_(True)
        gy = torch.randn(N, ld, device=dev).to(act)
        kind, sp = (CONV_GCN, conv.root_emb.weight) if args.gnn_type == "gcn" else (CONV_GIN, conv.eps)

        enc = conv_mod._edge_encoder_args(conv.edge_encoder, b.edge_attr, plan, d_g, ld)         # trainable table / weights

        def agg():
            y = ops.aggregate(x, plan, kind, d_g, sp, **enc)
            torch.autograd.grad(y, x, gy)
        out["aggregate"] = _timed(agg)
    else:
        layer = model.gnn_node.layers[0]
        x = (torch.randn(N, ld, device=dev) * 0.5).to(act).requires_grad_(True)
        pj = torch.randn(N, ld, device=dev).to(act).requires_grad_(True)
        pi = torch.randn(N, ld, device=dev).to(act).requires_grad_(True)
        gy = torch.randn(N, 4 * 13 * (d_g // 4), device=dev).to(act)

        def agg():
            y = ops.pna_reduce(x, pj, pi, plan, 4, d_g // 4, layer.avg_deg["log"])
            torch.autograd.grad(y, (x, pj, pi), gy)
        out["aggregate"] = _timed(agg)
    d, nh = args.d_model, args.nhead
    qkv = torch.randn(plan.n_rows, 3 * d, device=dev).to(act).requires_grad_(True)
    go = torch.randn(plan.n_rows, d, device=dev).to(act)
    drop = float(args.transformer_dropout)

    def mha():
        o = ops.mha_packed(qkv, plan, nh, drop_p=drop)
        torch.autograd.grad(o, qkv, go)
    out["mha"] = _timed(mha)
    ops.join_side_streams()
    return out


def in_step_roofline(cfg, args, model, lossf, buckets, b, hb, ns, step_ms, rank):
    """roofline objects of the hot-stage kernels from the IN-STEP timeline (every rank runs the stamped replay, rank 0
    builds the objects)"""
    from graphtrans_b200 import ops, trace
    pk = peaks()
    prof = trace.stamped_profile(model, lossf, buckets, b, reps=5)
    if rank != 0:
        return None
    tr = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tr = json.load(open(tpath)).get(cfg, {})
    es = 2 if ns.precision == "bf16" else 4
    alg = algorithmic(args, hb, es)
    alone = {}
    try:
        alone = alone_replays(args, model, b, hb, ns.precision)
    except Exception as e:  # noqa: BLE001
        alone = {"error": repr(e)[:160]}
    step_us = step_ms * 1e3
    out = []
    pna = args.model_type == "pna-transformer"
    names = ("gt_pna_reduce_fwd", "gt_pna_reduce_bwd") if pna else ("gt_aggregate_fwd", "gt_aggregate_bwd")
    n, us = trace.kernel_time(prof, names)
    if n:
        ach = n * alg["agg_bytes_per_launch"] / (us * 1e-6) / 1e9
        o = {"kernel": " + ".join(names) + (" (k_pna_fwd / k_pna_bwd)" if pna else " (k_agg_fwd* / k_agg_bwd*)"),
             "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
             "traffic": tr.get("aggregate"), "avg_launch_us": us / n, "launches_per_step": n, "share_of_step": us / step_us,
             "algorithmic_bytes_per_launch": alg["agg_bytes_per_launch"], "peak_source": pk["source"],
             "timing": "in-step: device timestamps after every kernel of the replayed CUDA graph (trace.stamped_profile), "
                       "minus the stamp kernel's own %.2f us per interval" % prof["cal_us"]}
        if isinstance(alone.get("aggregate"), float):
            a = 2 * alg["agg_bytes_per_launch"] / (alone["aggregate"] * 1e-3) / 1e9
            o["alone"] = {"achieved": a, "frac": a / pk["hbm"], "avg_launch_us": alone["aggregate"] / 2 * 1e3,
                          "note": "fwd + adjoint replayed back to back outside the step (same variants: trainable edge "
                                  "table, per-edge gradient store)"}
        out.append(o)
    full = ("gt_mha_fwd", "gt_mha_bwd", "gt_mha_local_fwd", "gt_mha_local_bwd")
    n, us = trace.kernel_time(prof, full)
    if n:
        n_layers = sum(1 for r in prof["records"] if r[0] in ("gt_mha_fwd", "gt_mha_local_fwd"))
        fl = n_layers * (alg["mha_fwd_flops"] + alg["mha_bwd_flops"])
        ach = fl / (us * 1e-6) / 1e12
        local = any(r[0] == "gt_mha_local_fwd" for r in prof["records"])
        o = {"kernel": ("gt_mha_local_fwd + gt_mha_local_bwd (tile-local tcgen05 attention)" if local else
                        "gt_mha_fwd + gt_mha_bwd (streamed tcgen05 attention: fwd, delta, dQ, dK/dV)"),
             "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
             "traffic": tr.get("mha_local" if local else "mha"), "avg_launch_us": us / n, "launches_per_step": n,
             "share_of_step": us / step_us, "useful_flops_per_step": fl, "peak_source": pk["source"],
             "note": "useful (unpadded, block-diagonal) flops only, %d full layers (the last layer runs the pooled-query kernel); "
                     "recompute flops of the backward are not counted; in-step timing" % n_layers}
        if isinstance(alone.get("mha"), float):
            a = (alg["mha_fwd_flops"] + alg["mha_bwd_flops"]) / (alone["mha"] * 1e-3) / 1e12
            o["alone"] = {"achieved": a, "frac": a / pk["tensor"], "ms_fwd_bwd": alone["mha"]}
        out.append(o)
    n, us = trace.kernel_time(prof, ("gt_mha_cls_fwd", "gt_mha_cls_bwd"))
    if n:
        kvb = alg["tokens"] * 2 * args.d_model * es              # k|v of every token: fwd reads them once; bwd reads them and writes dk|dv
        ach = (kvb + 2 * kvb) / (us * 1e-6) / 1e9
        out.append({"kernel": "gt_mha_cls_fwd + gt_mha_cls_bwd (pooled-query last layer)", "bound": "hbm", "achieved": ach,
                    "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None, "avg_launch_us": us / n,
                    "launches_per_step": n, "share_of_step": us / step_us, "peak_source": pk["source"]})
    # dense contractions: every gt_gemm / gt_gemm_stats of the step (forward, dX, split-K dW) with its own M, N, K
    fl = us = 0.0
    n = 0
    cal = prof["cal_us"]
    for name, _s, dur, a in prof["records"]:
        if name in ("gt_gemm", "gt_gemm_stats"):
            fl += 2.0 * a[9] * a[10] * a[11]
            us += max(dur - cal, 0.2)
            n += 1
    if n:
        ach = fl / (us * 1e-6) / 1e12
        out.append({"kernel": "gt_gemm + gt_gemm_stats (k_gemm_tc: all %d contractions of the step)" % n, "bound": "tensor",
                    "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"], "traffic": tr.get("gemm"),
                    "avg_launch_us": us / n, "launches_per_step": n, "share_of_step_stream_time": us / step_us,
                    "flops_per_step": fl, "peak_source": pk["source"],
                    "note": "summed over three streams (weight gradients overlap the main stream), so the share can exceed the critical path's"})
    primary = "gt_mha_fwd" if cfg == "code2" else ("gt_pna_reduce" if pna else "gt_aggregate_fwd")
    out.sort(key=lambda r: 0 if r["kernel"].startswith(primary) or (cfg == "code2" and r["kernel"].startswith("gt_mha_local")) else 1)
    meta = {"stamped_step_us": prof["span_us"], "calls": prof["n_calls"], "streams": prof["streams"], "stamp_us": prof["cal_us"]}
    return out, meta


def run_workload(ns, cfg, scaling, B_arg, rank, world, dev, K, W, full):
    """one workload end to end -> result dict (on rank 0; None elsewhere)"""
    from graphtrans_b200 import _lib, factory, loader, ops
    from graphtrans_b200.ddp import GradBuckets
    from graphtrans_b200.graphed import GraphedStep
    args = config_args(cfg)
    if scaling == "strong":
        B_global = B_arg or args.batch_size
        B = max(2, B_global // world)
        B_global = B * world
    else:
        B = B_arg or args.batch_size
        B_global = B * world
    lossf = factory.loss_fn(args)
    torch.manual_seed(0)
    model = factory.build_model(args).to(dev).train()
    in_graph = world > 1 and ns.comm == "graph" and not ns.eager
    buckets = GradBuckets(model, n_buckets=4, overlap=ns.eager or in_graph, direct=in_graph)
    graphed = None if ns.eager else GraphedStep(model, lossf, buckets, max_graphs=24, bucket=True)
    nd = ns.distinct_batches
    # every batch goes through the collate-time pipeline (shape bucket + int32 CSR + one pinned blob); the bucket grid is
    # fitted to the shape spread of a sample of batches (what a loader knows after its first pass over the dataset)
    sample = [synth.make_batch(args, B=B, seed=90000 + i) for i in range(12 if cfg == "syn" else 48)]
    grid = loader.Bucketer().fit([(int(s.batch.numel()), int(s.edge_index.shape[1])) for s in sample])
    del sample
    if graphed is not None:
        graphed.bucket = grid
    host = [loader.prepare(synth.make_batch(args, B=B, seed=1000 * rank + i), bucket=grid) for i in range(nd)]
    dev_batches = [b.to(dev) for b in host]
    ops.manual_seed(1234 + rank, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def eager_step(b):
        if not b.batch.is_cuda:
            b = b.to(dev, non_blocking=True)
        buckets.zero_grad()
        loss = lossf(model(b), b)
        loss.backward()
        ops.join_side_streams()
        buckets.finish()
        return loss.detach()

    step = eager_step if graphed is None else graphed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(W, 3, nd)):
        step(dev_batches[i % nd])
    barrier()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    k0 = _lib.kernel_count
    graph_kernels = 0
    evs = []
    barrier()
    wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()                                  # L2 flush between timed iterations (outside the events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(dev_batches[i % nd])
        e1.record()
        if graphed is not None:
            graph_kernels += graphed.last_kernels
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    launches = (_lib.kernel_count - k0) + graph_kernels
    t_dev = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    clocks = sampler.stop()
    tt = torch.tensor([t_dev], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev = float(tt)
    value = B_global * K / t_dev
    step_ms = t_dev / K * 1e3

    # ---- end to end on NEVER-SEEN host batches: collate-time prepare outside, H2D (prefetched blob) + step + loss D2H inside
    e2e = None
    if not ns.no_e2e:
        n_warm = 4 if cfg == "syn" else 16             # batches that warm the bucket grid before the timed, never-seen ones
        raw_warm = [synth.make_batch(args, B=B, seed=50000 + 1000 * rank + i) for i in range(n_warm)]
        fresh = [loader.prepare(r, bucket=grid) for r in raw_warm]
        fresh += [loader.prepare(synth.make_batch(args, B=B, seed=50000 + 1000 * rank + i), bucket=grid) for i in range(n_warm, K + n_warm)]
        cap0 = graphed.captures if graphed is not None else 0
        for hb in fresh[:n_warm]:                      # warm the bucket grid (captures of signatures not met so far)
            float(step(hb))
        timed = fresh[n_warm:]
        # steady state of an epoch: every bucket the timed batches fall into has been captured before - with OTHER data
        # (a warm batch padded up to that bucket); the timed batches themselves are never seen before their step.
        # Every rank replays the same number of times here (the graphs contain the gradient collectives).
        extra = []
        if graphed is not None:
            have = {(int(b.batch.numel()), int(b.edge_index.shape[1])) for b in fresh[:n_warm]} | \
                   {(int(b.batch.numel()), int(b.edge_index.shape[1])) for b in host}
            for shp in sorted({(int(b.batch.numel()), int(b.edge_index.shape[1])) for b in timed} - have):
                src = next((r for r in raw_warm if r.batch.numel() < shp[0] and r.edge_index.shape[1] <= shp[1]), None)
                for t in range(64):                    # no warm batch fits under this bucket: draw more until one does
                    if src is not None:
                        break
                    r = synth.make_batch(args, B=B, seed=70000 + 1000 * rank + t)
                    if r.batch.numel() < shp[0] and r.edge_index.shape[1] <= shp[1]:
                        src = r
                if src is not None:
                    extra.append(loader.pack(loader.attach_csr(loader.pad_to_bucket(src, shp[0], shp[1]))))
        n_extra = torch.tensor([len(extra)], device=dev)
        if world > 1:
            dist.all_reduce(n_extra, op=dist.ReduceOp.MAX)
        for j in range(int(n_extra)):
            float(step(extra[j] if j < len(extra) else fresh[0]))
        cap1 = graphed.captures if graphed is not None else 0
        # a generational collection inside a 40 ms window (20 steps of 2 ms) would be the measurement: collect now, keep
        # the collector off while the host loop is timed (as a training loop that cares about step jitter does)
        gc.collect()
        gc.disable()
        barrier()
        t0 = time.perf_counter()
        if graphed is not None:
            graphed.prefetch(timed[0])
            pending = None
            for i in range(K):
                if i + 1 < K:
                    graphed.prefetch(timed[i + 1])     # H2D of the next batch overlaps this step (copy stream)
                h = graphed.step_async(timed[i])       # the step + an asynchronous D2H copy of its loss
                if pending is not None:
                    pending.item()                     # host reads the loss of step i-1 while step i runs
                pending = h
            pending.item()
        else:
            for i in range(K):
                float(step(timed[i]))                  # D2H read of the loss
        barrier()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        gc.enable()
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        cap2 = graphed.captures if graphed is not None else 0
        capt = torch.tensor([cap2 - cap1], device=dev)
        if world > 1:
            dist.all_reduce(capt, op=dist.ReduceOp.MAX)
        e2e = {"value": B_global * K / float(te), "unit": UNIT,
               "h2d_bytes_per_step": int(statistics.mean(b.nbytes() for b in timed)), "d2h_bytes_per_step": 4,
               "fresh_batches": True, "graph_captures_before_timing": cap1, "graph_captures_while_warming_buckets": cap1 - cap0,
               "graph_captures_inside_timed_region": int(capt),
               "distinct_bucketed_shapes": len({(int(b.batch.numel()), int(b.edge_index.shape[1])) for b in timed}),
               "bucket_grid": ("one bucket %s fitted to a sample of batches (spread <= 6 %%)" % (grid.fixed,) if grid.fixed else
                               "steps of 1/%d .. 1/%d of the size" % (1 << grid.log2_steps, 1 << (grid.log2_steps - 1))),
               "pipeline": "loader.prepare (shape bucket + int32 CSR + one pinned blob, collate time, untimed) -> prefetched H2D of "
                           "the blob into a persistent staging buffer -> CUDA-graph replay -> asynchronous D2H copy of the loss, "
                           "read by the host one step later (every step's loss is read inside the timed region)"}
        del fresh, timed

    roof = None
    if full and not ns.no_roofline and graphed is not None:
        try:
            roof = in_step_roofline(cfg, args, model, lossf, buckets, dev_batches[0], host[0], ns, step_ms, rank)
        except Exception as e:  # noqa: BLE001
            roof = ([{"kernel": "in-step roofline failed", "error": repr(e)[:300]}], {}) if rank == 0 else None

    with_opt = None
    if full and not ns.no_optimizer and graphed is not None:
        from graphtrans_b200.optim import FusedAdamW
        opt = FusedAdamW(buckets, lr=1e-4, weight_decay=1e-5)
        gstep = GraphedStep(model, lossf, buckets, max_graphs=24, optimizer=opt)
        for i in range(max(3, nd)):
            gstep(dev_batches[i % nd])
        barrier()
        evs2 = []
        for i in range(K):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss_t = gstep(dev_batches[i % nd])
            e1.record()
            evs2.append((e0, e1))
        barrier()
        t2 = torch.tensor([sum(a.elapsed_time(b) for a, b in evs2) * 1e-3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        with_opt = {"value": B_global * K / float(t2), "unit": UNIT, "ms_per_step": float(t2) / K * 1e3,
                    "step": "zero_grad+forward+loss+backward" + ("+allreduce" if world > 1 else "") + "+fused AdamW (gt_adamw_multi), "
                            + ("all inside one CUDA graph" if gstep.opt_in_graph else "optimizer after the host-issued allreduce"),
                    "finite_loss": bool(torch.isfinite(loss_t).item())}

    if rank != 0:
        return None
    res = {
        "value": value, "unit": UNIT, "ms_per_step": step_ms, "scaling": scaling, "graphs_per_gpu": B, "global_batch": B_global,
        "config": {"workload": workload_name(cfg, args, B_global, scaling, world),
                   "l2": "256 MiB buffer written between timed steps", "distinct_batches": nd,
                   "dropout": {"gnn": args.gnn_dropout, "transformer": args.transformer_dropout},
                   "step": "zero_grad+forward+loss+backward" + (
                       "+NCCL gradient allreduce (4 buckets, " + ("captured inside the step's CUDA graph, overlapped with the backward)"
                                                                   if in_graph else "issued from the host after the replay)") if world > 1 else ""),
                   "wgrad_stream": not ns.no_wgrad_stream, "branch_stream": not ns.no_branch_stream,
                   "batches": "shape-bucket padded (slack nodes / edges), int32 CSR built at collate time, one blob per batch",
                   "launch": "eager (Python launches every kernel)" if graphed is None else
                             "CUDA-graph replay per bucketed shape signature; inputs copied into static buffers inside the timed region",
                   "wall_ms_per_step_incl_flush": wall / K * 1e3},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "with_optimizer": with_opt,
    }
    if roof is not None:
        res["roofline"] = dict(roof[0][0]) if roof[0] else None
        res["roofline_kernels"] = roof[0]
        res["in_step_trace"] = roof[1]
    return res


def main_b200(ns):
    from graphtrans_b200 import _lib, ops
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != ns.gpus and world == 1 and ns.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    ops.set_precision(ns.precision)
    if not ns.no_wgrad_stream:
        ops.enable_wgrad_stream(True, dev)   # weight / bias / embedding gradients overlap the rest of the backward
    if not ns.no_branch_stream:
        ops.enable_branch_stream(True, dev)  # virtual-node update of a GNN layer runs next to its conv
    K, W = ns.steps, max(ns.warmup, 3)
    head = run_workload(ns, ns.config, ns.scaling, ns.batch, rank, world, dev, K, W, full=True)
    extras = {}
    if not ns.no_extra_configs:
        for cfg in EXTRA_CONFIGS:
            if cfg == ns.config:
                continue
            torch.cuda.empty_cache()
            try:
                r = run_workload(ns, cfg, "weak", None, rank, world, dev, min(K, 20), 3, full=True)
            except Exception as e:  # noqa: BLE001
                r = {"error": repr(e)[:300]}
            if rank == 0:
                if r is not None:
                    r.pop("clocks", None)
                extras[cfg] = r
    cpu = None
    if not ns.no_cpu_baseline and rank == 0 and world == 1:      # reported at N = 1 only (the other ranks would idle)
        args = config_args(ns.config)
        gps, ms, desc = oracle_run(ns.config, head["global_batch"], 2, 1)
        cpu = {"value": gps, "unit": UNIT, "cores": desc["cores"], "kind": desc["kind"], "sample": desc["sample"], "ms_per_step": ms}
        if "molpcba" in extras and isinstance(extras["molpcba"], dict) and "value" in extras["molpcba"]:
            g2, m2, d2 = oracle_run("molpcba", 512, 2, 1)
            extras["molpcba"]["cpu_baseline"] = {"value": g2, "unit": UNIT, "cores": d2["cores"], "kind": d2["kind"],
                                                 "sample": d2["sample"], "ms_per_step": m2}
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": head["scaling"], "vs_baseline": None,
                "dtype": ns.precision, "data": "synthetic", "config": head["config"], "clocks": head["clocks"], "e2e": head["e2e"],
                "gpu_launches": head["gpu_launches"], "with_optimizer": head["with_optimizer"],
                "roofline": head.get("roofline"), "roofline_kernels": head.get("roofline_kernels"),
                "in_step_trace": head.get("in_step_trace"), "cpu_baseline": cpu,
                "graphs_per_gpu": head["graphs_per_gpu"], "global_batch": head["global_batch"],
                "configs": extras}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ns = parse()
    if ns.impl == "reference":
        main_reference(ns)
    elif ns.impl == "torch-gpu":
        main_torch_gpu(ns)
    else:
        main_b200(ns)
