#!/usr/bin/env python
"""graphs/sec forward+backward of the GraphTrans hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config molpcba|code2|syn|code2-pna|nci1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

One "step" = zero_grad -> forward -> loss -> backward (+ bucketed NCCL gradient allreduce when
N > 1) over one synthetic batch of the named config.  Rank 0 prints ONE JSON line.
  value : whole-job graphs/s with the batches already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public model API with HOST (pinned) batches: H2D copy of the
          step's batch and a D2H read of the loss inside the timed region
  roofline     : dominant hot-stage kernel class, achieved = algorithmic bytes|flops (SURVEY §8d,
                 DESIGN.md) / CUDA-event time of those launches, peak from MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (port of the reference's PyTorch path) timed on the host cores
"""
import argparse
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
# CPU sample sizes (graphs per step) keeping the reference arm / cpu_baseline to ~10-30 s of CPU work
CPU_SAMPLE_B = {"molpcba": 128, "code2": 8, "syn": 32, "code2-pna": 8, "nci1": 32}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="molpcba", choices=list(synth.CONFIGS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=None, help="graphs per GPU per step (default: the config's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true", help="skip the second figure (step + fused AdamW update)")
    ap.add_argument("--distinct-batches", type=int, default=4)
    ap.add_argument("--no-wgrad-stream", action="store_true", help="keep weight gradients on the main stream")
    ap.add_argument("--no-branch-stream", action="store_true", help="keep the virtual-node branch on the main stream")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph replay: launch every kernel from Python")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tensor=1590.0, source="fallback")


def config_args(ns):
    args = synth.make_args(ns.config)
    if ns.config == "code2-pna":
        probe = synth.make_batch(args, B=args.batch_size, seed=1234)
        args.deg = synth.in_degree_histogram(probe, 800)
    return args


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


# ----------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(ns, steps, warmup):
    """Times the CPU oracle (restatement of the reference's PyTorch/PyG path, fp32, all host
    threads) on a bounded sample of the workload. -> (graphs/s, ms/step, description dict)"""
    from graphtrans_b200 import factory
    from oracle import graphtrans_oracle as O
    args = config_args(ns)
    args.gnn_dropout = 0.0           # the oracle has no dropout (identity); the product runs the configured p
    args.transformer_dropout = 0.0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = min(CPU_SAMPLE_B[ns.config], ns.batch or args.batch_size)
    batch = synth.make_batch(args, B=Bs, seed=0)
    torch.manual_seed(0)
    sd = factory.build_model(args).state_dict()
    for _ in range(warmup):
        O.fwd_bwd(sd, args, batch, dtype=torch.float32)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.fwd_bwd(sd, args, batch, dtype=torch.float32)
    dt = (time.perf_counter() - t0) / steps
    return Bs / dt, dt * 1e3, dict(cores=cores, threads=torch.get_num_threads(), kind="port",
                                   sample=f"{Bs} of {ns.batch or args.batch_size} graphs/step ({ns.config} shape, seed 0), "
                                          f"fp32, dropout 0, {warmup} warm-up + {steps} timed fwd+bwd steps of "
                                          f"oracle/graphtrans_oracle.py")


def main_reference(ns):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = ns.steps, ns.warmup
    # bound the arm to a few minutes whatever K the driver passes: cap total CPU steps
    est_steps = min(steps, 10)
    gps, ms, desc = cpu_reference_run(ns, est_steps, min(warmup, 2))
    args = synth.make_args(ns.config)
    line = {
        "impl": "reference", "metric": METRIC, "value": gps, "unit": UNIT, "n_gpus": ns.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": ns.scaling, "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload_name(ns, args), "timed_steps": est_steps,
                   "note": "CPU path of the reference (oracle port; the reference is pure Python and its "
                           "third-party deps are not installable here, so there is no oracle/_ref)"},
        "cpu_baseline": {"value": gps, "unit": UNIT, "cores": desc["cores"], "kind": desc["kind"],
                         "sample": desc["sample"]},
        "e2e": {"value": gps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(ns, args):
    B = ns.batch or args.batch_size
    return (f"{ns.config}: GraphTrans {args.gnn_type.upper() if args.model_type == 'gnn-transformer' else 'PNA'}"
            f"{'-Virtual' if args.gnn_virtual_node else ''} JK={args.gnn_JK} d_g={args.gnn_emb_dim} "
            f"L_g={args.gnn_num_layer} d={args.d_model} L_t={args.num_encoder_layers} B={B}/GPU")


# ----------------------------------------------------------------------------------- GPU arm
def algorithmic(args, batch, es):
    """per-step algorithmic work of the two hot stages for one batch"""
    N = batch.batch.numel()
    E = batch.edge_index.shape[1]
    ea = 0 if batch.edge_attr is None else batch.edge_attr.numel() * batch.edge_attr.element_size()
    d_g = args.gnn_emb_dim
    n = torch.bincount(batch.batch)
    t = n.clamp(max=int(args.max_input_len)) + (1 if args.graph_pooling == "cls" else 0)
    sum_t2 = float((t.double() ** 2).sum())
    if args.model_type == "pna-transformer":
        agg_bytes = N * d_g * es * 3 + 13 * N * d_g * es + 16 * E      # x, pi, pj read; 13F x towers written
    else:
        agg_bytes = 2 * N * d_g * es + 16 * E + ea
    return dict(N=N, E=E, agg_bytes_per_launch=agg_bytes, mha_fwd_flops=4.0 * args.d_model * sum_t2,
                mha_bwd_flops=8.0 * args.d_model * sum_t2, tokens=int(t.sum()))


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


def hot_kernel_roofline(args, model, b, hb, ns, step_ms):
    from graphtrans_b200 import ops
    from graphtrans_b200._lib import CONV_GCN, CONV_GIN
    from graphtrans_b200.modules import conv as conv_mod
    pk = peaks()
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(ns.config, {})
    es = 2 if ns.precision == "bf16" else 4
    alg = algorithmic(args, hb, es)
    act = ops.act_dtype()
    plan = ops.GraphPlan(b.edge_index, b.batch, b.num_graphs, int(args.max_input_len), cls=args.graph_pooling == "cls",
                         max_nodes=getattr(hb, "max_nodes", None))     # same attention path as the model takes
    N, d_g, ld = alg["N"], args.gnn_emb_dim, ops.ldp(args.gnn_emb_dim)
    out = []
    torch.manual_seed(0)
    # stage 1: message-passing aggregation fwd + adjoint (one GNN layer)
    if args.model_type == "gnn-transformer":
        conv = model.gnn_node.convs[1]
        x = (torch.randn(N, ld, device=b.batch.device) * 0.5).to(act)
        x[:, d_g:] = 0
        x.requires_grad_(True)
        gy = torch.randn(N, ld, device=b.batch.device).to(act)
        enc = conv_mod._edge_encoder_args(conv.edge_encoder, b.edge_attr, plan, d_g, ld)
        enc = {k: (v.detach() if torch.is_tensor(v) and v.dtype.is_floating_point and k == "table" else v) for k, v in enc.items()}
        kind, sp = (CONV_GCN, conv.root_emb.weight) if args.gnn_type == "gcn" else (CONV_GIN, conv.eps)

        def agg():
            y = ops.aggregate(x, plan, kind, d_g, sp, **enc)
            torch.autograd.grad(y, x, gy)      # no .grad accumulation kernels inside the timed launches
        kname = "k_agg_fwd2 / k_agg_bwd2" if enc.get("edge_kind") == conv_mod.EDGE_LINEAR else "k_agg_fwd3 / k_agg_bwd3"
        name = f"gt_aggregate_fwd + gt_aggregate_bwd ({kname})"
    else:
        layer = model.gnn_node.layers[0]
        x = (torch.randn(N, ld, device=b.batch.device) * 0.5).to(act).requires_grad_(True)
        pj = torch.randn(N, ld, device=b.batch.device).to(act).requires_grad_(True)
        pi = torch.randn(N, ld, device=b.batch.device).to(act).requires_grad_(True)
        gy = torch.randn(N, 4 * 13 * (d_g // 4), device=b.batch.device).to(act)

        def agg():
            y = ops.pna_reduce(x, pj, pi, plan, 4, d_g // 4, layer.avg_deg["log"])
            torch.autograd.grad(y, (x, pj, pi), gy)
        name = "gt_pna_reduce_fwd + gt_pna_reduce_bwd"
    ms = _timed(agg)                       # the edge table is a constant here: exactly the two named kernels run
    ach = 2 * alg["agg_bytes_per_launch"] / (ms * 1e-3) / 1e9
    out.append({"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                "traffic": traffic.get("aggregate"), "avg_launch_us": ms / 2 * 1e3,
                "share_of_step": ms * args.gnn_num_layer / step_ms,
                "algorithmic_bytes_per_launch": alg["agg_bytes_per_launch"], "peak_source": pk["source"],
                "note": "the [N, d_g] matrix of these configs is L2-resident (8-10 MB << 126 MB L2)"})
    if args.model_type == "gnn-transformer" and enc.get("edge_kind") == conv_mod.EDGE_TABLE:
        # the edge-table gradient of the same layer (weight-gradient stream): gathers x[src] and dout[dst] of every edge
        from graphtrans_b200._lib import call as _call, dt_of as _dt, ptr as _p
        table, etype = enc["table"], enc["etype"]
        src_t, dst_t, type_t, _ = plan.edges_by_type(plan._edge_index, etype, table.shape[0])
        dtab = torch.zeros_like(table)
        xc = x.detach()

        def tab_grad():
            _call("gt_aggregate_table_grad", _dt(xc), kind, _p(xc), _p(gy), N, d_g, ld, _p(plan.rowptr_src), plan.E, _p(src_t),
                  _p(dst_t), _p(type_t), _p(table), table.shape[0], _p(dtab))
        ms_t = _timed(tab_grad)
        tg_bytes = 2 * plan.E * d_g * es + 12 * plan.E
        ach_t = tg_bytes / (ms_t * 1e-3) / 1e9
        out.append({"kernel": "gt_aggregate_table_grad (k_agg_table_grad, weight-gradient stream)", "bound": "hbm", "achieved": ach_t,
                    "peak": pk["hbm"], "unit": "GB/s", "frac": ach_t / pk["hbm"], "traffic": None, "avg_launch_us": ms_t * 1e3,
                    "share_of_step": None, "algorithmic_bytes_per_launch": tg_bytes, "peak_source": pk["source"],
                    "note": "gather formulation: 2 E d_g s + 12 E bytes (x[src], dout[dst], sorted edge triples); L2-resident rows. "
                            "The bf16 step no longer launches it (the adjoint writes the per-edge gradient and the table "
                            "gradient is a one-hot contraction, GT_TABLE_GRAD_GEMM=1); it remains the fp32 / small-batch path"})
    # stage 2: masked MHA fwd + bwd over the packed tokens (one encoder layer)
    d, nh = args.d_model, args.nhead
    qkv = torch.randn(plan.n_rows, 3 * d, device=b.batch.device).to(act).requires_grad_(True)
    go = torch.randn(plan.n_rows, d, device=b.batch.device).to(act)
    drop = float(args.transformer_dropout)

    def mha():
        o = ops.mha_packed(qkv, plan, nh, drop_p=drop)
        torch.autograd.grad(o, qkv, go)
    ms = _timed(mha)
    fl = alg["mha_fwd_flops"] + alg["mha_bwd_flops"]
    ach = fl / (ms * 1e-3) / 1e12
    local = plan.loc_tiles is not None and act == torch.bfloat16
    mha_name = ("gt_mha_local_fwd + gt_mha_local_bwd (k_mha_loc_fwd, k_mha_loc_bwd: graph-aligned 128-row tiles)" if local else
                "gt_mha_fwd + gt_mha_bwd (k_mha_tc_fwd, k_mha_tc_bwd<dQ>, k_mha_tc_bwd<dKV>, k_mha_delta)")
    n_launch = 2 if local else 4
    out.append({"kernel": mha_name, "bound": "tensor",
                "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
                "traffic": traffic.get("mha_local" if local else "mha"), "avg_launch_us": ms / n_launch * 1e3,
                "share_of_step": ms * args.num_encoder_layers / step_ms,
                "useful_flops_fwd_bwd": fl, "peak_source": pk["source"],
                "note": "useful (unpadded, block-diagonal) flops only; recompute flops of the backward are not counted"})
    # dense: the largest contraction of the model (fwd + dX + dW)
    M, Nn, Kk = alg["tokens"], args.dim_feedforward, d
    w = torch.randn(Nn, Kk, device=b.batch.device, requires_grad=True)
    bias = torch.zeros(Nn, device=b.batch.device, requires_grad=True)
    xx = torch.randn(plan.n_rows, Kk, device=b.batch.device).to(act).requires_grad_(True)
    gg = torch.randn(plan.n_rows, ops.ldp(Nn), device=b.batch.device).to(act)

    def ffn1():
        y = ops.linear(xx, w, bias, relu=True)
        torch.autograd.grad(y, (xx, w, bias), gg)
    ms = _timed(ffn1)
    fl = 3 * 2.0 * plan.n_rows * Nn * Kk
    ach = fl / (ms * 1e-3) / 1e12
    out.append({"kernel": "gt_gemm fwd + dX + dW of the FFN up-projection [tokens x dim_feedforward x d_model] (k_gemm_tc)",
                "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
                "traffic": traffic.get("gemm"), "avg_launch_us": ms / 5 * 1e3, "share_of_step": None, "peak_source": pk["source"],
                "note": "includes relu_bwd and colsum launches of the linear backward"})
    out.sort(key=lambda r: -(r["share_of_step"] or 0))
    # the headline `roofline` object is the hot stage BASELINE.json quotes the config on ("HBM GB/s (conv) + tensor-pipe %
    # (MHA)"): configs 2 / 4 / 5 / 1 are the scatter-bound ones -> stage-1 aggregation (HBM roofline); config 3 (long padded
    # sequences) stresses the masked-MHA tcgen05 path -> tensor roofline.  Every measured kernel stays in roofline_kernels.
    primary = "gt_mha" if ns.config == "code2" else ("gt_pna_reduce" if args.model_type == "pna-transformer" else "gt_aggregate_fwd")
    out.sort(key=lambda r: 0 if r["kernel"].startswith(primary) else 1)
    return out


def main_b200(ns):
    from graphtrans_b200 import _lib, factory, ops
    from graphtrans_b200.ddp import GradBuckets

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != ns.gpus:
        if world == 1 and ns.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    ops.set_precision(ns.precision)
    args = config_args(ns)
    B = ns.batch or args.batch_size
    if ns.scaling == "strong" and world > 1:
        B = max(2, B // world)
    lossf = factory.loss_fn(args)
    torch.manual_seed(0)
    model = factory.build_model(args).to(dev).train()
    from graphtrans_b200.graphed import GraphedStep
    buckets = GradBuckets(model, n_buckets=4, overlap=ns.eager)
    graphed = None if ns.eager else GraphedStep(model, lossf, buckets, max_graphs=2 * ns.distinct_batches + 2)
    host_batches = [synth.make_batch(args, B=B, seed=1000 * rank + i).pin_memory() for i in range(ns.distinct_batches)]
    dev_batches = [b.to(dev) for b in host_batches]
    ops.manual_seed(1234 + rank, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    if not ns.no_wgrad_stream:
        ops.enable_wgrad_stream(True, dev)   # weight / bias / embedding gradients overlap the rest of the backward
    if not ns.no_branch_stream:
        ops.enable_branch_stream(True, dev)  # virtual-node update of a GNN layer runs next to its conv

    def eager_step(b):
        buckets.zero_grad()
        loss = lossf(model(b), b)
        loss.backward()
        ops.join_side_streams()
        buckets.finish()
        return loss.detach()

    # product step: CUDA-graph replay per batch shape (captured during warm-up), eager with --eager
    step = eager_step if graphed is None else graphed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = ns.steps, ns.warmup
    for i in range(max(W, 3, len(dev_batches))):
        step(dev_batches[i % len(dev_batches)])
    if graphed is not None and not ns.no_e2e:
        for hb in host_batches:          # host-resident batches share the signatures captured above
            step(hb)
    barrier()
    sampler = ClockSampler(local)
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
        step(dev_batches[i % len(dev_batches)])
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
    value = B * world * K / t_dev

    # ---- end to end: host (pinned) batches through the public API, H2D + loss D2H inside the timed region
    e2e = None
    if not ns.no_e2e:
        def e2e_step(hb):
            # public API call with a HOST batch: H2D copy of every input tensor, step, D2H read of the loss
            return float(step(hb) if graphed is not None else step(hb.to(dev, non_blocking=True)))

        for i in range(2):
            e2e_step(host_batches[i % len(host_batches)])
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            e2e_step(host_batches[i % len(host_batches)])
        barrier()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * K / float(te), "unit": UNIT,
               "h2d_bytes_per_step": int(statistics.mean(b.nbytes() for b in host_batches)),
               "d2h_bytes_per_step": 4}

    # ---- roofline (rank 0, right after the timed region, same process): the two hot-stage kernels and the largest
    # dense contraction are replayed on this batch's real plan / shapes, back to back behind a GPU-side head start
    # (so Python launch latency is outside the CUDA events), and timed with CUDA events on the launching stream.
    # achieved = algorithmic bytes|flops per launch / measured launch time.
    roof, roof_all = None, None
    if not ns.no_roofline and rank == 0:
        roof_all = hot_kernel_roofline(args, model, dev_batches[0], host_batches[0], ns, t_dev / K * 1e3)
        roof = dict(roof_all[0]) if roof_all else None

    # ---- second figure: the same step followed by the fused AdamW update (+ allreduce when N > 1);
    # measured last because it moves the weights
    with_opt = None
    if not ns.no_optimizer and graphed is not None:
        from graphtrans_b200.optim import FusedAdamW
        opt = FusedAdamW(buckets, lr=1e-4, weight_decay=1e-5)
        gstep = GraphedStep(model, lossf, buckets, max_graphs=2 * ns.distinct_batches + 2, optimizer=opt)
        for i in range(max(3, len(dev_batches))):
            gstep(dev_batches[i % len(dev_batches)])
        barrier()
        evs2 = []
        for i in range(K):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss_t = gstep(dev_batches[i % len(dev_batches)])
            e1.record()
            evs2.append((e0, e1))
        barrier()
        t2 = torch.tensor([sum(a.elapsed_time(b) for a, b in evs2) * 1e-3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        with_opt = {"value": B * world * K / float(t2), "unit": UNIT, "ms_per_step": float(t2) / K * 1e3,
                    "step": "zero_grad+forward+loss+backward" + ("+allreduce" if world > 1 else "") + "+fused AdamW (gt_adamw_multi)",
                    "finite_loss": bool(torch.isfinite(loss_t).item())}

    cpu = None
    if not ns.no_cpu_baseline and rank == 0 and world == 1:      # reported at N = 1 only (the other ranks would idle)
        gps, ms, desc = cpu_reference_run(ns, 3, 1)
        cpu = {"value": gps, "unit": UNIT, "cores": desc["cores"], "kind": desc["kind"], "sample": desc["sample"],
               "ms_per_step": ms}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": ns.scaling, "vs_baseline": None,
            "dtype": ns.precision, "data": "synthetic",
            "config": {"workload": workload_name(ns, args), "l2": "256 MiB buffer written between timed steps",
                       "distinct_batches": len(dev_batches), "dropout": {"gnn": args.gnn_dropout,
                                                                         "transformer": args.transformer_dropout},
                       "step": "zero_grad+forward+loss+backward" + ("+NCCL gradient allreduce (4 buckets)" if world > 1 else ""),
                       "wgrad_stream": not ns.no_wgrad_stream, "branch_stream": not ns.no_branch_stream,
                       "launch": "eager (Python launches every kernel)" if graphed is None else
                                 "CUDA-graph replay per batch shape signature (captured in warm-up); inputs copied into static buffers inside the timed region",
                       "wall_ms_per_step_incl_flush": wall / K * 1e3},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "with_optimizer": with_opt, "roofline": roof,
            "roofline_kernels": roof_all,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ns = parse()
    if ns.impl == "reference":
        main_reference(ns)
    else:
        main_b200(ns)
