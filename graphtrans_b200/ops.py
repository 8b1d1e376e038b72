"""Autograd-visible operators over the C-ABI kernels (host side of the hot path).

Every forward AND backward here is a hand-written sm_100a kernel reached through
graphtrans_b200._lib.call; torch is used for memory (torch.empty/zeros) and for autograd
bookkeeping only.  Feature matrices are physical [rows, ld] tensors with ld = ldp(d) (logical
width d, zero pad columns), see include/graphtrans_b200.h.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import (CONV_GCN, CONV_GIN, EDGE_LINEAR, EDGE_NONE, EDGE_TABLE, EPI_ACCUM, EPI_OUT_F32,
                   EPI_RELU, GT_BF16, GT_F32, call, dt_of, ptr)

_PRECISION = os.environ.get("GT_PRECISION", "fp32")
GEMM_IMPL = int(os.environ.get("GT_GEMM_IMPL", "0"))   # 0 auto, 1 CUDA-core, 2 tcgen05 only
FUSE_COLSTATS = int(os.environ.get("GT_FUSE_COLSTATS", "1"))         # BatchNorm statistics in the producing GEMM's epilogue
SKIP_ZERO_BIAS_GRAD = int(os.environ.get("GT_SKIP_ZERO_BIAS_GRAD", "1"))   # bias of a Linear feeding train-mode BN: gradient == 0
TABLE_GRAD_GEMM = int(os.environ.get("GT_TABLE_GRAD_GEMM", "1"))           # edge-table gradient as a one-hot contraction (bf16)
# ... formed INSIDE the adjoint kernel (k_agg_bwd4: the per-edge gradients go straight from registers into a tcgen05
# contraction with the one-hot edge types; no [E, ld] tensor); same eligibility rule as csrc/aggregate.cu launch_bwd
# Opt-in: at config 4 it measures 757 us against 796 us for adjoint + one-hot GEMM (both issue bound), see DESIGN.md
TABLE_GRAD_FUSED = int(os.environ.get("GT_AGG_TABLE_FUSED", "0"))
# GIN adjoint with the ReLU mask formed by packed bf16 compares against per-layer bf16 thresholds (exact; k_agg_bwd3p)
AGG_PACKED = int(os.environ.get("GT_AGG_PACKED", "1"))
MHA_IMPL = int(os.environ.get("GT_MHA_IMPL", "0"))
# fp32 parity mode with every contraction ON the tcgen05 kernel: operands split into three bf16 terms, six products
# accumulated in fp32 (SURVEY §7 "fp32 parity on tensor cores"); 0 = exact CUDA-core contractions (default parity path)
GEMM_TC_PARITY = int(os.environ.get("GT_GEMM_TC_PARITY", "0"))


def set_precision(p: str):
    """'fp32': fp32 activations + exact fp32 contractions (parity mode, <=1e-3 contract);
    'bf16': bf16 activations, tcgen05 bf16 contractions with fp32 accumulate (throughput mode)."""
    global _PRECISION
    assert p in ("fp32", "bf16")
    _PRECISION = p


def precision() -> str:
    return _PRECISION


def act_dtype() -> torch.dtype:
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


# ----------------------------------------------------------------------------- dropout RNG
_rng = {}          # device index -> int64[2] tensor {seed, step} (read by the kernels on the device)
_salt = [0]        # call-site counter inside one step


def rng_state(device) -> torch.Tensor:
    idx = torch.device(device).index or 0
    st = _rng.get(idx)
    if st is None:
        st = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
        _rng[idx] = st
    return st


def manual_seed(seed: int, device="cuda"):
    rng_state(device).copy_(torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64))


# ----------------------------------------------------------------------------- bf16 operand copies of the weights
class _W16Registry:
    """All registered fp32 master weights get a zero-padded bf16 operand copy ([rows, ldp(cols)]) that ONE kernel
    (gt_cast_multi) refreshes at the top of every forward, instead of one cast kernel per nn.Linear call.
    `register_heads` additionally STACKS the copies of a list of equally shaped Linear heads (Code2: 5 x [5002, d]) as
    one [H * ldp(N), ldp(K)] operand (+ one fp32 [H * ldp(N)] bias vector), so that all heads are one contraction."""

    def __init__(self):
        self.params, self.copies, self.desc, self.ptrs, self.blocks = [], {}, None, None, 0
        self.stacks, self.stack_copies, self._stacked_ids = [], {}, set()
        self.extra = {}
        self.blockdiags, self.bd_copies = [], {}

    def register_heads(self, linears):
        linears = list(linears)
        if len(linears) < 2 or any(l.bias is None or l.weight.shape != linears[0].weight.shape for l in linears):
            return
        self.stacks.append([(l.weight, l.bias) for l in linears])
        self._stacked_ids.update(id(l.weight) for l in linears)
        self.params = [w for w in self.params if id(w) not in self._stacked_ids]
        self.desc = None

    def register_blockdiag(self, key, weights, biases=None, col_lo=0, K=None):
        """block-diagonal operand of equally shaped tower Linears (PNA pre / post MLPs, reference modules/pna_layer.py:
        102-118): diag_t = weights[t][:, col_lo:col_lo+K] -> ONE bf16 [T*Fo, ldp(T*K)] operand (+ fp32 bias [T*Fo]),
        refreshed with the other operand copies; `bd_lookup(key)` -> (operand, ld, bias or None)"""
        weights = list(weights)
        K = weights[0].shape[1] - col_lo if K is None else K
        self.blockdiags.append((key, weights, list(biases) if biases is not None else None, int(col_lo), int(K)))
        self.desc = None

    def bd_lookup(self, key):
        return self.bd_copies.get(key)

    def register(self, module: torch.nn.Module):
        seen = {id(p) for p in self.params} | self._stacked_ids
        for m in module.modules():
            cand = []
            if isinstance(m, torch.nn.Linear):
                cand.append(m.weight)
            if isinstance(m, torch.nn.MultiheadAttention) and m.in_proj_weight is not None:
                cand.append(m.in_proj_weight)
            for w in cand:
                if id(w) not in seen:
                    seen.add(id(w))
                    self.params.append(w)
        self.desc = None

    def _all_ptrs(self):
        return (tuple(w.data_ptr() for w in self.params) + tuple(t.data_ptr() for st in self.stacks for wb in st for t in wb)
                + tuple(w.data_ptr() for bd in self.blockdiags for w in bd[1]))

    def _build(self, device):
        recs, blk = [], 0
        self.copies, self.stack_copies = {}, {}
        for w in self.params:
            if not w.is_cuda or w.dtype != torch.float32 or not w.is_contiguous():
                continue
            rows, cols = w.shape
            ld = ldp(cols)
            c = torch.empty(rows, ld, dtype=torch.bfloat16, device=w.device)
            self.copies[id(w)] = (c, ld)
            recs.append([w.data_ptr(), c.data_ptr(), rows, cols, ld, blk, 0, 0])
            blk += (rows * ld + 2047) // 2048
        for st in self.stacks:
            w0 = st[0][0]
            if not all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for wb in st for t in wb):
                continue
            rows, cols = w0.shape
            ld, rp = ldp(cols), ldp(rows)
            wst = torch.zeros(len(st) * rp, ld, dtype=torch.bfloat16, device=w0.device)   # pad rows stay zero
            bst = torch.zeros(len(st) * rp, dtype=torch.float32, device=w0.device)
            for h, (w, b) in enumerate(st):
                recs.append([w.data_ptr(), wst.data_ptr() + h * rp * ld * 2, rows, cols, ld, blk, 0, 0])
                blk += (rows * ld + 2047) // 2048
                recs.append([b.data_ptr(), bst.data_ptr() + h * rp * 4, 1, rows, -rp, blk, 0, 0])   # ld < 0: fp32 copy
                blk += (rp + 2047) // 2048
            self.stack_copies[id(w0)] = (wst, bst, rp, ld)
        self.bd_copies = {}
        for key, ws, bs, col_lo, K in self.blockdiags:
            if not all(w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() for w in ws):
                continue
            T, Fo, Kw = len(ws), ws[0].shape[0], ws[0].shape[1]
            ld = ldp(T * K)
            op = torch.zeros(T * Fo, ld, dtype=torch.bfloat16, device=ws[0].device)     # off-diagonal blocks stay zero
            bias = torch.zeros(T * Fo, dtype=torch.float32, device=ws[0].device) if bs is not None else None
            for t, w in enumerate(ws):
                recs.append([w.data_ptr() + col_lo * 4, op.data_ptr() + (t * Fo * ld + t * K) * 2, Fo, K, ld, blk, Kw, K])
                blk += (Fo * K + 2047) // 2048
                if bs is not None:
                    recs.append([bs[t].data_ptr(), bias.data_ptr() + t * Fo * 4, 1, Fo, -Fo, blk, 0, 0])
                    blk += (Fo + 2047) // 2048
            self.bd_copies[key] = (op, ld, bias)
        self.blocks = blk
        self.ptrs = self._all_ptrs()
        self.desc = torch.tensor(recs, dtype=torch.int64, device=device) if recs else None

    def refresh(self, device):
        if not self.params and not self.stacks and not self.blockdiags:
            return
        if self.desc is None or self.ptrs != self._all_ptrs():
            self._build(device)
        if self.desc is not None:
            call("gt_cast_multi", ptr(self.desc), self.desc.shape[0], self.blocks)

    def lookup(self, w):
        ent = self.copies.get(id(w))
        return ent if ent is not None else self.extra.get(id(w))

    def put(self, w, copy, ld):
        """operand copy of a derived (non-parameter) fp32 matrix, e.g. an eval-time BatchNorm-folded weight"""
        self.extra[id(w)] = (copy, ld)

    def stack_lookup(self, w0):
        return self.stack_copies.get(id(w0))


W16Registry = _W16Registry
w16 = _W16Registry()   # registry of the model whose step is running (begin_step(registry=...)); default: a global one


_train_steps = [0]      # bumped by every training forward: eval-time derived weights (fold_bn) are recomputed after it


def begin_step(device, cast_now=True, registry=None):
    """advance the device-side dropout step counter (one launch; graph-capturable) and restart the
    per-step call-site salts.  Called by the model at the top of every training forward.  cast_now=False leaves the
    refresh of the bf16 weight copies to the caller (the model hands it to GraphPlan as `side_work`, so it runs on the
    branch stream next to the CSR build)."""
    global w16
    if registry is not None:   # every model owns its operand copies: a captured step keeps valid pointers when
        w16 = registry         # another model is built or stepped in the same process
    _salt[0] = 0
    _train_steps[0] += 1
    call("gt_rng_advance", ptr(rng_state(device)))
    if _PRECISION == "bf16" and cast_now:
        w16.refresh(device)
    a = _arena(device)
    if a["off"] or not a["armed"]:
        a["buf"].zero_()
    a["off"], a["armed"] = 0, True


def next_salt() -> int:
    _salt[0] += 1
    return _salt[0]


# ----------------------------------------------------------------------------- weight-gradient side stream
# Weight / bias gradients are off the critical path of the backward pass (nothing upstream consumes them), so they
# can run on a second stream next to the memory-bound kernels of the next layer (BN / LayerNorm backward, aggregation
# adjoint ...), which co-reside on an SM with the one-CTA-per-SM persistent GEMM.  Opt-in: whoever enables it must
# call join_side_streams() after backward() and before reading gradients (GraphedStep and bench.py do).
_wgrad = {"stream": None, "used": False, "keep": []}
WGRAD_SIDE_MAX_ELEMS = int(os.environ.get("GT_WGRAD_SIDE_MAX_ELEMS", str(16 << 20)))
WGRAD_SIDE_MAX_ROWS = int(os.environ.get("GT_WGRAD_SIDE_MAX_ROWS", str(32 << 10)))


def enable_wgrad_stream(on=True, device="cuda"):
    _wgrad["stream"] = torch.cuda.Stream(device=device) if on else None
    _wgrad["used"] = False
    _wgrad["keep"] = []


def join_side_streams():
    """make the current stream wait for the weight-gradient stream and the parallel-branch stream (call after
    backward(), before anything reads the gradients)"""
    for side in (_branch, _wgrad):
        # only a stream that forked since the last join: waiting on an untouched stream would make a capturing
        # stream depend on uncaptured work (cudaErrorStreamCaptureIsolation)
        if side["stream"] is not None and side["used"]:
            torch.cuda.current_stream().wait_stream(side["stream"])
            side["used"] = False
    _wgrad["keep"].clear()


join_wgrad_stream = join_side_streams


class _WgradCtx:
    """`with _WgradCtx(on, tensors...)`: if `on` and the side stream is enabled, run the enclosed launches on it,
    ordered after everything issued so far on the current stream; the tensors are kept alive for that stream.
    `on` must be False whenever a result of the enclosed launches is handed back to autograd (which assumes the
    node's own stream)."""

    def __init__(self, on, *tensors):
        self.tensors = [t for t in tensors if t is not None]
        # Large operands (config 4: [5e5, 512] gradients) make the weight-gradient kernels HBM-bound device-filling
        # launches of their own: run next to the equally HBM-bound main stream they only thrash (measured: 39.2 ms with,
        # 38.3 ms without the side stream on config 4; the small configs gain 3-4 % from it) - those stay on the main stream.
        # "Large" = many ROWS (the contraction length of a weight gradient) as well as many elements: a wide but short
        # operand (config 5's [14 k, 3328] tower inputs) is a brief launch inside a latency-bound step and keeps the overlap
        big = (max((t.numel() for t in self.tensors), default=0) > WGRAD_SIDE_MAX_ELEMS and
               max((t.shape[0] for t in self.tensors if t.dim() > 1), default=0) > WGRAD_SIDE_MAX_ROWS)
        self.on = bool(on) and not big
        self.ctx = None

    def __enter__(self):
        st = _wgrad["stream"] if self.on else None
        if st is not None:
            _wgrad["used"] = True
            st.wait_stream(torch.cuda.current_stream())
            # the operands stay referenced until join_side_streams(): the autograd engine accumulates gradients IN PLACE
            # into a tensor it holds the only reference to (e.g. the residual branch of a LayerNorm backward hands the
            # same tensor to two consumers) - a write the side stream's reads are not ordered against.  A second reference
            # makes the engine add out of place.  (record_stream only guards against reuse after free.)
            _wgrad["keep"].extend(self.tensors)
            for t in self.tensors:
                t.record_stream(st)
            self.ctx = torch.cuda.stream(st)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


# ----------------------------------------------------------------------------- parallel branch stream
# The virtual-node update of a GNN layer (global_add_pool -> 2-layer MLP on [B, d_g], reference gnn_module.py:217-229)
# only needs the layer INPUT and is consumed at the very end of the layer, so it is independent of the conv of that
# layer: a dozen kernels of a few CTAs each that can run next to the conv on a second stream (a parallel branch of
# the captured CUDA graph).  autograd runs the backward of every node on the stream of its forward and inserts the
# cross-stream waits itself, so the backward overlaps the same way.  Opt-in like the weight-gradient stream: whoever
# enables it calls join_side_streams() after backward() (GraphedStep and bench.py do).
_branch = {"stream": None, "used": False}


def enable_branch_stream(on=True, device="cuda"):
    # the virtual-node update joins the main stream at the end of every GNN layer, i.e. it is on the critical path
    # (unlike the weight gradients): high priority, like the stream GraphedStep captures the step on
    prio = -1 if os.environ.get("GT_BRANCH_PRIORITY", "1") == "1" else 0
    _branch["stream"] = torch.cuda.Stream(device=device, priority=prio) if on else None
    _branch["used"] = False


class Branch:
    """`with Branch(inputs...) as br: ...` runs the enclosed ops on the branch stream (when enabled), ordered after
    everything issued so far on the current stream; `br.join(outputs...)` makes the current stream wait for them."""

    def __init__(self, *tensors):
        self.st = _branch["stream"]
        self.tensors = [t for t in tensors if t is not None]
        self.ctx = None

    def __enter__(self):
        if self.st is not None:
            _branch["used"] = True
            self.st.wait_stream(torch.cuda.current_stream())
            for t in self.tensors:
                t.record_stream(self.st)
            self.ctx = torch.cuda.stream(self.st)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
            self.ctx = None
        return False

    def join(self, *outputs):
        if self.st is not None:
            cur = torch.cuda.current_stream()
            cur.wait_stream(self.st)
            for t in outputs:
                if t is not None:
                    t.record_stream(cur)


# ----------------------------------------------------------------------------- fused gradient delivery
# A parameter may carry `_gt_main_grad` (an fp32 tensor of its shape, normally a view into the flat arena of
# graphtrans_b200.ddp.GradBuckets, zeroed once per step).  The backward kernels then ACCUMULATE the parameter
# gradient straight into it and return None to autograd: no temporary, no zero fill, no autograd add kernel.
# `_gt_grad_ready()` (if present) tells the bucket manager that one expected contribution has landed.
def _main_grad(param):
    return getattr(param, "_gt_main_grad", None) if param is not None else None


def _grad_target(param, shape=None):
    """-> (tensor to accumulate into, value to return to autograd)"""
    mg = _main_grad(param)
    if mg is not None:
        return mg, None
    # returned to autograd, which may keep it as .grad: must be ordinary memory (never the step arena)
    g = torch.zeros(tuple(param.shape if shape is None else shape), dtype=torch.float32, device=param.device)
    return g, g


def _grad_done(param):
    cb = getattr(param, "_gt_grad_ready", None)
    if cb is not None:
        cb()


# small zero-initialised scratch that never leaves an op (BatchNorm statistics / backward reductions): carved from
# one arena that begin_step() clears with a single memset instead of one fill kernel per buffer
_ZERO_ARENA_BYTES = 8 << 20
_zero_arena = {}


def _arena(device):
    idx = torch.device(device).index or 0
    a = _zero_arena.get(idx)
    if a is None:
        a = {"buf": torch.zeros(_ZERO_ARENA_BYTES, dtype=torch.uint8, device=device), "off": 0, "armed": False}
        _zero_arena[idx] = a
    return a


def zeros_small(numel, dtype, device):
    esz = 8 if dtype == torch.float64 else 4
    nbytes = (numel * esz + 255) // 256 * 256
    a = _arena(device)
    if not a["armed"] or a["off"] + nbytes > _ZERO_ARENA_BYTES:
        return torch.zeros(numel, dtype=dtype, device=device)
    off = a["off"]
    a["off"] = off + nbytes
    return a["buf"][off:off + numel * esz].view(dtype)


def zeros_f32(shape, device):
    """fresh zeroed fp32 tensor that may be handed to autograd (ordinary allocation)"""
    return torch.zeros(shape, dtype=torch.float32, device=device)


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, salt):
        x = x.contiguous()
        y = torch.empty_like(x)
        call("gt_dropout", dt_of(x), ptr(x), x.numel(), ptr(y), float(p), ptr(rng_state(x.device)), salt)
        ctx.meta = (p, salt)
        return y

    @staticmethod
    def backward(ctx, g):
        p, salt = ctx.meta
        g = g.contiguous()
        dx = torch.empty_like(g)
        call("gt_dropout", dt_of(g), ptr(g), g.numel(), ptr(dx), float(p), ptr(rng_state(g.device)), salt)
        return dx, None, None


def dropout(x, p):
    """F.dropout(x, p, training=True) on a physical matrix (numel % 4 == 0)."""
    if not p:
        return x
    return _DropoutFn.apply(x, p, next_salt())


def ldp(d: int) -> int:
    """physical leading dimension: logical width rounded up to 8 elements (16 B in bf16)."""
    return (d + 7) // 8 * 8


def pad_cols(x: torch.Tensor, ld: int, dtype=None) -> torch.Tensor:
    """[rows, d] (any float dtype, any stride) -> contiguous [rows, ld] of `dtype`, zero padded."""
    dtype = dtype or x.dtype
    rows, d = x.shape
    if d == ld and x.dtype == dtype and x.is_contiguous():
        return x
    return _CastPadFn.apply(x, ld, dtype)


class _CastPadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ld, dtype):
        x = x.contiguous() if x.stride(-1) != 1 else x
        rows, d = x.shape
        out = torch.empty(rows, ld, dtype=dtype, device=x.device)
        call("gt_cast_pad", dt_of(x), ptr(x), rows, d, x.stride(0), dt_of(out), ptr(out), rows, ld, ld)
        ctx.meta = (d, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        d, dtype = ctx.meta
        g = g.contiguous()
        rows, ld = g.shape
        out = torch.empty(rows, d, dtype=dtype, device=g.device)
        call("gt_cast_pad", dt_of(g), ptr(g), rows, d, ld, dt_of(out), ptr(out), rows, d, d)
        return out, None, None


def cast_to(x: torch.Tensor, dtype) -> torch.Tensor:
    """dtype change of a physical [rows, ld] matrix (differentiable)."""
    if x.dtype == dtype:
        return x
    return _CastPadFn.apply(x, x.shape[1], dtype)


# ----------------------------------------------------------------------------- graph plan
MHA_LOCAL = int(os.environ.get("GT_MHA_LOCAL", "1"))    # 0: always the streamed attention kernels


def token_bucket(max_nodes, max_input_len=1000, cls=True):
    """tokens-per-graph bound rounded up to {32, 64, 96, 128} (coarse, so that CUDA-graph signatures and launch
    bounds do not change with every batch); None = unknown or some graph exceeds one 128-row tile"""
    if max_nodes is None:
        return None
    t = min(int(max_nodes), int(max_input_len)) + 1     # + <CLS>; kept when there is none so that the bucket is a
    for b in (32, 64, 96, 128):                         # function of ceil((max_nodes + 1) / 32) alone (graphed._signature)
        if t <= b:
            return b
    return None


class GraphPlan:
    """Integer metadata of one batch: both CSRs and the packed-token plan. Built on device by
    gt_csr_build / gt_batch_plan without any host synchronisation (B comes from
    `batch.num_graphs` when present, otherwise one .item() like reference gnn_module.py:195)."""

    def __init__(self, edge_index, batch, num_graphs=None, max_input_len=1000, cls=True, side_work=None, max_nodes=None,
                 prebuilt=None, slack=False):
        """side_work: optional callable launched on the branch stream together with the token plan (the two CSR
        sorts stay on the current stream): the integer prep of a batch is three independent chains.
        prebuilt: the six int32 CSR arrays (rowptr_dst, src_by_dst, eid_by_dst, rowptr_src, dst_by_src, eid_by_src) when
        the loader built them at collate time (loader.attach_csr): gt_csr_build is skipped.
        slack: the batch ends with shape-bucket slack nodes (batch id == num_graphs, loader.pad_to_bucket): they belong
        to no graph; `m_valid` (device int32[1] = number of real nodes) keeps them out of the BatchNorm statistics."""
        # max_nodes: optional host-side upper bound of the nodes per graph (collate-time metadata, e.g.
        # synth.GraphBatch.max_nodes).  When every graph fits one 128-row tile the tile-local attention kernels apply.
        _lib.require_cuda(edge_index, batch)
        dev = batch.device
        N = batch.numel()
        E = edge_index.shape[1]
        if num_graphs is None:
            if slack:
                raise RuntimeError("GraphPlan: a batch with slack nodes needs num_graphs")
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("GraphPlan: batch.num_graphs is required under CUDA-graph capture (batch[-1].item() would sync)")
            B = int(batch[-1].item()) + 1       # one device->host read, like reference modules/gnn_module.py:195
        else:
            B = int(num_graphs)
        self.N, self.E, self.B, self.L, self.cls = N, E, B, int(max_input_len), bool(cls)
        i32 = dict(dtype=torch.int32, device=dev)
        ei = edge_index.contiguous()
        if prebuilt is not None:
            (self.rowptr_dst, self.src_by_dst, self.eid_by_dst, self.rowptr_src, self.dst_by_src, self.eid_by_src) = prebuilt
            for t, n in zip(prebuilt, (N + 1, E, E, N + 1, E, E)):
                if t.dtype != torch.int32 or not t.is_cuda or t.numel() < max(n, 1) or not t.is_contiguous():
                    raise RuntimeError("GraphPlan: prebuilt CSR arrays must be contiguous int32 CUDA tensors of the batch's sizes")
            work = None
        else:
            self.rowptr_dst = torch.empty(N + 1, **i32)
            self.rowptr_src = torch.empty(N + 1, **i32)
            self.src_by_dst = torch.empty(max(E, 1), **i32)
            self.eid_by_dst = torch.empty(max(E, 1), **i32)
            self.dst_by_src = torch.empty(max(E, 1), **i32)
            self.eid_by_src = torch.empty(max(E, 1), **i32)
            work = torch.empty(2 * (N + 1), **i32)
        self.node_off = torch.empty(B + 1, **i32)
        self.kept = torch.empty(B, **i32)
        self.tok_off = torch.empty(B + 1, **i32)
        self.n_rows = N + B  # static upper bound of packed token rows (exact when nothing is truncated)
        self.tok2node = torch.empty(self.n_rows, **i32)
        self.tok_graph = torch.empty(self.n_rows, **i32)
        self.node_graph = torch.empty(N, **i32)
        self.node2tok = torch.empty(N, **i32)
        self.cls_rows = torch.empty(B, **i32)
        self.scalars = torch.empty(4, **i32)
        # attention metadata of the packed layout (row key ranges, tile ranges): once per batch for all layers
        self.row_bounds = torch.empty(2 * self.n_rows, **i32)
        self.tile_bounds = torch.empty(2 * ((self.n_rows + 127) // 128), **i32)
        b = batch.contiguous()
        plan_out = (self.node_off, self.kept, self.tok_off, self.tok2node, self.tok_graph, self.node_graph, self.node2tok,
                    self.cls_rows, self.scalars, self.row_bounds, self.tile_bounds)
        self.loc_tiles = self.loc_count = None
        self.loc_max_tiles = 0
        bucket = token_bucket(max_nodes, self.L, self.cls)
        if bucket is not None and MHA_LOCAL:
            self.loc_max_tiles = max(1, min(B, -(-self.n_rows // (129 - bucket))))
            self.loc_tiles = torch.empty(2 * self.loc_max_tiles, **i32)
            self.loc_count = torch.empty(1, **i32)
            plan_out = plan_out + (self.loc_tiles, self.loc_count)
        br = Branch(b, *plan_out)
        with br:    # token plan (+ the caller's side work) next to the CSR sorts
            if side_work is not None:
                side_work()
            call("gt_batch_plan", ptr(b), N, B, self.L, int(self.cls), ptr(self.node_off), ptr(self.kept), ptr(self.tok_off),
                 ptr(self.tok2node), ptr(self.tok_graph), ptr(self.node_graph), ptr(self.node2tok),
                 ptr(self.cls_rows), ptr(self.scalars))
            call("gt_mha_meta", ptr(self.tok_graph), ptr(self.tok_off), self.n_rows, B, ptr(self.row_bounds),
                 ptr(self.tile_bounds))
            if self.loc_tiles is not None:
                call("gt_mha_local_tiles", ptr(self.tok_off), B, self.loc_max_tiles, ptr(self.loc_tiles), ptr(self.loc_count))
        if prebuilt is None:
            call("gt_csr_build", ptr(ei), E, N, ptr(self.rowptr_dst), ptr(self.src_by_dst), ptr(self.eid_by_dst),
                 ptr(self.rowptr_src), ptr(self.dst_by_src), ptr(self.eid_by_src), ptr(work))
        br.join()
        self.m_valid = self.node_off[B:B + 1] if slack else None
        # the static row bound n_rows = N + B can exceed the token count (slack nodes, truncated graphs): kernels that
        # only write the rows of real tokens then leave a tail that must not hold garbage (it enters later contractions)
        self.has_tail = bool(slack) or max_nodes is None or int(max_nodes) > self.L
        self._etype = {}
        self._slots = {}
        self._by_type = {}
        self._edge_index = ei
        self._S = None

    def edge_type(self, edge_attr: torch.Tensor, dims) -> torch.Tensor:
        """combined categorical edge id (mixed radix over `dims`) for table edge encoders."""
        key = (edge_attr.data_ptr(), tuple(dims))
        if key not in self._etype:
            mult, m = [], 1
            for dsz in reversed(dims):
                mult.append(m)
                m *= dsz
            mult = list(reversed(mult))
            et = torch.empty(max(self.E, 1), dtype=torch.int32, device=edge_attr.device)
            ea = edge_attr.contiguous()
            arr = (ctypes.c_int32 * len(mult))(*mult)
            call("gt_edge_type", ptr(ea), self.E, len(mult), arr, ptr(et))
            self._etype[key] = et
        return self._etype[key]

    def edge_slots(self, conv, edge_kind, edge_attr=None, etype=None):
        """per-slot copies of the per-edge data for both CSRs (computed once per batch and encoder kind, reused by
        all layers): ((norm, etype, attr) in target-sorted order, (norm, etype, attr) in source-sorted order)"""
        key = (conv, edge_kind, 0 if edge_attr is None else edge_attr.data_ptr(), 0 if etype is None else etype.data_ptr())
        hit = self._slots.get(key)
        if hit is not None:
            return hit
        E, dev = max(self.E, 1), self.rowptr_dst.device
        kdim = edge_attr.shape[1] if edge_kind == EDGE_LINEAR else 0
        res = []
        for rp, nbr, eid in ((self.rowptr_dst, self.src_by_dst, self.eid_by_dst),
                             (self.rowptr_src, self.dst_by_src, self.eid_by_src)):
            norm = torch.empty(E, dtype=torch.float32, device=dev) if conv == CONV_GCN else None
            ets = torch.empty(E, dtype=torch.int32, device=dev) if edge_kind == EDGE_TABLE else None
            ats = torch.empty(E, kdim, dtype=torch.float32, device=dev) if edge_kind == EDGE_LINEAR else None
            if norm is not None or ets is not None or ats is not None:
                call("gt_edge_slots", ptr(rp), ptr(nbr), ptr(eid), ptr(self.rowptr_src), self.E, self.N, ptr(etype),
                     ptr(edge_attr) if edge_kind == EDGE_LINEAR else None, kdim, ptr(norm), ptr(ets), ptr(ats))
            res.append((norm, ets, ats))
        self._slots[key] = tuple(res)
        return self._slots[key]

    def edges_by_type(self, edge_index, etype, ntypes):
        """edges counting-sorted by combined edge type (once per batch, reused by the table-gradient kernel of every
        layer): (src_t, dst_t, type_t) int32 [E]"""
        key = (etype.data_ptr(), int(ntypes))
        hit = self._by_type.get(key)
        if hit is None:
            E, dev = max(self.E, 1), etype.device
            i32 = dict(dtype=torch.int32, device=dev)
            type_ptr = torch.empty(ntypes + 1, **i32)
            src_t, dst_t, type_t = torch.empty(E, **i32), torch.empty(E, **i32), torch.empty(E, **i32)
            work = torch.empty(2 * ntypes, **i32)
            call("gt_edges_by_type", ptr(edge_index.contiguous()), ptr(etype), self.E, int(ntypes), ptr(type_ptr),
                 ptr(src_t), ptr(dst_t), ptr(type_t), ptr(work))
            hit = (src_t, dst_t, type_t, type_ptr)
            self._by_type[key] = hit
        return hit

    def zero_table_grad(self, shape, device):
        """zeroed fp32 [ntypes, ld] accumulation target for one layer's edge-table gradient: slices of ONE tensor per
        step and table shape (one fill kernel for all layers instead of one per layer)"""
        pool = self.__dict__.setdefault("_dtab_pool", {})
        ent = pool.get(shape)
        if ent is None or ent[1] >= ent[0].shape[0]:
            ent = [torch.zeros((8,) + tuple(shape), dtype=torch.float32, device=device), 0]
            pool[shape] = ent
        ent[1] += 1
        return ent[0][ent[1] - 1]

    def type_onehot(self, etype_slot, ntypes):
        """bf16 one-hot [E, ldp(ntypes)] of the per-slot edge types (once per batch, reused by every layer's
        edge-table gradient contraction)"""
        key = (etype_slot.data_ptr(), int(ntypes))
        hit = self._by_type.get(("onehot",) + key)
        if hit is None:
            E = max(self.E, 1)
            idx = etype_slot.long()
            r_pad = ldp(int(ntypes))
            oh = torch.empty(E, r_pad, dtype=torch.bfloat16, device=etype_slot.device)
            a_idx = (ctypes.c_void_p * 1)(idx.data_ptr())
            a_str = (ctypes.c_int64 * 1)(1)
            a_clp = (ctypes.c_int64 * 1)(int(ntypes) - 1)
            a_base = (ctypes.c_int32 * 1)(0)
            call("gt_onehot", E, 1, a_idx, a_str, a_clp, a_base, r_pad, ptr(oh))
            hit = (oh, r_pad, idx)
            self._by_type[("onehot",) + key] = hit
        return hit

    @property
    def S(self) -> int:
        """padded length min(max n_i, L) - needs one device->host read (public pad_batch API only)."""
        if self._S is None:
            self._S = int(self.scalars[0].item())
        return self._S


_CSR_FIELDS = ("csr_rowptr_dst", "csr_src_by_dst", "csr_eid_by_dst", "csr_rowptr_src", "csr_dst_by_src", "csr_eid_by_src")


def plan_for(batched_data, max_input_len=1000, cls=True, side_work=None):
    """GraphPlan of a batch object, honouring the optional collate-time metadata the loader attaches: `num_graphs`,
    `max_nodes`, the int32 CSR (`csr_*`, loader.attach_csr) and `slack` (loader.pad_to_bucket)"""
    pre = None
    if getattr(batched_data, _CSR_FIELDS[0], None) is not None:
        pre = tuple(getattr(batched_data, k) for k in _CSR_FIELDS)
    return GraphPlan(batched_data.edge_index, batched_data.batch, getattr(batched_data, "num_graphs", None), max_input_len,
                     cls=cls, side_work=side_work, max_nodes=getattr(batched_data, "max_nodes", None), prebuilt=pre,
                     slack=bool(getattr(batched_data, "slack", False)))


# ----------------------------------------------------------------------------- node encoders
ONEHOT_MIN_ROWS = int(os.environ.get("GT_ONEHOT_MIN_ROWS", "2048"))      # below this many index rows the atomics kernel is cheap (edge-type digit tables)
ONEHOT_MAX_TABLE = 256      # tables with more rows keep global atomics (little contention per row)
ONEHOT_MAX_TOTAL = 1024


class _EmbedSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, *tables):
        idx, strides, clamps, N, d, ld, dtype = meta
        out = torch.empty(N, ld, dtype=dtype, device=tables[0].device)
        n = len(tables)
        a_idx = (ctypes.c_void_p * n)(*[t.data_ptr() for t in idx])
        a_str = (ctypes.c_int64 * n)(*strides)
        a_clp = (ctypes.c_int64 * n)(*clamps)
        a_tab = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tables])
        call("gt_embed_sum_fwd", dt_of(out), ptr(out), N, d, ld, n, a_idx, a_str, a_clp, a_tab)
        ctx.meta = meta
        ctx.tables = tables
        ctx.onehot = None
        # few-row vocabularies: their gradient is the contraction OneHot^T . dout on the tensor cores (bf16 mode);
        # the one-hot operand is written here, once per batch
        if dtype == torch.bfloat16 and N >= ONEHOT_MIN_ROWS and any(t.requires_grad for t in tables):
            base, rows, acc = [], [], 0
            for c, t in enumerate(tables):
                r = min(int(clamps[c]) + 1, t.shape[0])
                if t.requires_grad and r <= ONEHOT_MAX_TABLE and acc + r <= ONEHOT_MAX_TOTAL:
                    base.append(acc)
                    rows.append(r)
                    acc += r
                else:
                    base.append(-1)
                    rows.append(0)
            if acc:
                r_pad = ldp(acc)
                oh = torch.empty(N, r_pad, dtype=torch.bfloat16, device=out.device)
                a_base = (ctypes.c_int32 * n)(*base)
                call("gt_onehot", N, n, a_idx, a_str, a_clp, a_base, r_pad, ptr(oh))
                ctx.onehot = (oh, base, rows, acc, r_pad)
        return out

    @staticmethod
    def backward(ctx, g):
        idx, strides, clamps, N, d, ld, dtype = ctx.meta
        g = g.contiguous()
        n = len(idx)
        targets = [_grad_target(t) for t in ctx.tables]
        side_ok = all(_main_grad(t) is not None for t in ctx.tables)   # leaf-only gradients: off the critical path
        # every tensor the side-stream kernels READ is listed: it must neither be reused by the allocator nor modified in
        # place by the autograd engine before join_side_streams()
        with _WgradCtx(side_ok, g, ctx.onehot[0] if ctx.onehot is not None else None, *idx):
            rest = list(range(n))
            if ctx.onehot is not None:
                oh, base, rows, R, r_pad = ctx.onehot
                ldt = ldp(d)
                temp = zeros_small(R * ldt, torch.float32, g.device)
                # temp[R, d] += OneHot^T [R x N] . g [N x d]  (both operands MN-major, split-K over the nodes)
                _gemm_raw(GT_BF16, oh.data_ptr(), 1, r_pad, g.data_ptr(), 1, ld, temp.data_ptr(), ldt, R, d, N, d, None,
                          None, 0, EPI_ACCUM | EPI_OUT_F32)
                a_base = (ctypes.c_int32 * n)(*base)
                a_rows = (ctypes.c_int32 * n)(*rows)
                a_tab = (ctypes.c_void_p * n)(*[t[0].data_ptr() for t in targets])
                call("gt_embed_unpack", ptr(temp), ldt, d, R, n, a_base, a_rows, a_tab)
                rest = [c for c in range(n) if base[c] < 0]
            if rest:
                m = len(rest)
                a_idx = (ctypes.c_void_p * m)(*[idx[c].data_ptr() for c in rest])
                a_str = (ctypes.c_int64 * m)(*[strides[c] for c in rest])
                a_clp = (ctypes.c_int64 * m)(*[clamps[c] for c in rest])
                a_tab = (ctypes.c_void_p * m)(*[targets[c][0].data_ptr() for c in rest])
                call("gt_embed_sum_bwd", dt_of(g), ptr(g), N, d, ld, m, a_idx, a_str, a_clp, a_tab)
            for t in ctx.tables:
                _grad_done(t)
        return (None, *[t[1] for t in targets])


def embed_sum(index_cols, tables, clamps=None, dtype=None):
    """out[i] = sum_c tables[c][min(index_cols[c][i], clamp_c)] -> physical [N, ldp(d)] (activation dtype unless
    `dtype` is given).  index_cols: list of int64 1-D views (any stride) of length N."""
    d = tables[0].shape[1]
    N = index_cols[0].shape[0]
    if d % 4:
        raise RuntimeError("embedding width must be a multiple of 4")
    strides = [c.stride(0) if c.dim() else 1 for c in index_cols]
    clamps = clamps or [t.shape[0] - 1 for t in tables]
    meta = (list(index_cols), strides, list(clamps), N, d, ldp(d), dtype or act_dtype())
    out = _EmbedSumFn.apply(meta, *tables)
    # every table delivers its gradient straight into the arena: the backward of this op runs on the weight-gradient
    # stream, and so may whatever produces the gradient of `out` if `out` only feeds leaf gradients (edge tables)
    out._gt_leaf_side = all(_main_grad(t) is not None for t in tables)
    return out


# ----------------------------------------------------------------------------- dense layers
def _gemm_raw(dt, A, a_mn, lda, Bm, b_mn, ldb, C, ldc, M, N, K, n_fill, bias, resid, ldr, flags, impl=None,
              drop_p=0.0, rng=None, salt=0, col_stats=None, m_valid=None):
    """raw-pointer gt_gemm (A, Bm, C are device addresses so strided sub-blocks need no copies); col_stats: fp64
    [2 * ldc] that receives the BatchNorm column statistics of C in the same pass (gt_gemm_stats; m_valid: device
    int32[1] = number of leading real rows when the matrix carries shape-bucket slack rows)"""
    if dt == GT_F32 and GEMM_TC_PARITY and (impl is None or impl == 2):
        return _gemm_split3(A, a_mn, lda, Bm, b_mn, ldb, C, ldc, M, N, K, n_fill, bias, resid, ldr, flags, drop_p, rng, salt,
                            col_stats, m_valid)
    if col_stats is not None:
        call("gt_gemm_stats", dt, A, int(a_mn), lda, Bm, int(b_mn), ldb, C, ldc, M, N, K, n_fill, ptr(bias), ptr(resid), ldr,
             flags, float(drop_p), rng, salt, GEMM_IMPL if impl is None else impl, ptr(col_stats), ptr(m_valid))
        return
    call("gt_gemm", dt, A, int(a_mn), lda, Bm, int(b_mn), ldb, C, ldc, M, N, K, n_fill, ptr(bias), ptr(resid), ldr,
         flags, float(drop_p), rng, salt, GEMM_IMPL if impl is None else impl)


def _gemm_split3(A, a_mn, lda, Bm, b_mn, ldb, C, ldc, M, N, K, n_fill, bias, resid, ldr, flags, drop_p, rng, salt, col_stats,
                 m_valid):
    """fp32 contraction on the tcgen05 kernel: A = a0 + a1 + a2, B = b0 + b1 + b2 (bf16 terms, gt_split3), C = sum of the
    six products a_i . b_j with i + j <= 2 (what is dropped is below 2^-22 of |A||B|), accumulated in fp32.  The five
    small terms are accumulated first into an fp32 scratch, the leading term a0 . b0 is the last call and takes the scratch
    as its residual operand, so bias / ReLU / dropout / column fill / BatchNorm statistics are the normal epilogue."""
    dev = torch.device("cuda", torch.cuda.current_device())
    ra, ca = (K, M) if a_mn else (M, K)           # stored rows x contiguous columns of each operand
    rb, cb = (K, N) if b_mn else (N, K)
    la, lb = ldp(ca), ldp(cb)
    sa = torch.empty(3, ra, la, dtype=torch.bfloat16, device=dev)
    sb = torch.empty(3, rb, lb, dtype=torch.bfloat16, device=dev)
    call("gt_split3", A, ra, ca, lda, ptr(sa), la)
    call("gt_split3", Bm, rb, cb, ldb, ptr(sb), lb)
    pa = [sa[i].data_ptr() for i in range(3)]
    pb = [sb[i].data_ptr() for i in range(3)]
    small = ((0, 1), (1, 0), (0, 2), (2, 0), (1, 1))
    if flags & EPI_ACCUM:                          # weight gradients: every term accumulates straight into C
        for i, j in ((0, 0),) + small:
            call("gt_gemm", GT_BF16, pa[i], int(a_mn), la, pb[j], int(b_mn), lb, C, ldc, M, N, K, n_fill, None, None, 0,
                 EPI_ACCUM | EPI_OUT_F32, 0.0, None, 0, 2)
        return
    if resid is not None:
        if ldr != ldc:
            raise RuntimeError("tensor-core parity mode: the residual operand must share the output's row pitch")
        part = resid.to(torch.float32).clone().contiguous()
        if part.stride(0) != ldc:
            raise RuntimeError("tensor-core parity mode: unexpected residual layout")
    else:
        part = torch.zeros(M, ldc, dtype=torch.float32, device=dev)
    for i, j in small:
        call("gt_gemm", GT_BF16, pa[i], int(a_mn), la, pb[j], int(b_mn), lb, ptr(part), ldc, M, N, K, N, None, None, 0,
             EPI_ACCUM | EPI_OUT_F32, 0.0, None, 0, 2)
    fl = (flags & EPI_RELU) | EPI_OUT_F32 | _lib.EPI_RESID_F32
    if col_stats is not None:
        call("gt_gemm_stats", GT_BF16, pa[0], int(a_mn), la, pb[0], int(b_mn), lb, C, ldc, M, N, K, n_fill, ptr(bias), ptr(part), ldc,
             fl, float(drop_p), rng, salt, 2, ptr(col_stats), ptr(m_valid))
    else:
        call("gt_gemm", GT_BF16, pa[0], int(a_mn), la, pb[0], int(b_mn), lb, C, ldc, M, N, K, n_fill, ptr(bias), ptr(part), ldc,
             fl, float(drop_p), rng, salt, 2)


class _LinearFn(torch.autograd.Function):
    """y = x W[:, off:off+K]^T + b (optional fused ReLU / residual). x physical [M, ld_in] (logical
    K columns); y physical [M, ldp(N)] in x.dtype (or fp32 when out_f32).  `off` selects a column
    block of the weight (JK=cat: gnn2transformer applied to the parts without concatenating)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, out_f32, resid, off, K, drop_p, salt, want_stats=False, m_valid=None,
                row_off=0, n_out=None, passthrough=False):
        x_in = x
        x = x.contiguous()
        M, ld_in = x.shape
        N, Kw = weight.shape
        if n_out is not None:       # row block [row_off, row_off + n_out) of the weight / bias (q or k|v part of in_proj)
            N = int(n_out)
        K = Kw - off if K is None else K
        if K > ld_in:
            raise RuntimeError(f"linear: input width {ld_in} < in_features {K}")
        wf = weight.contiguous()
        bias_param = bias
        if bias is not None and (row_off or N != bias.shape[0]):
            bias = bias[row_off:row_off + N]
        if x.dtype == torch.float32:
            w, ldw, wptr = wf, Kw, wf.data_ptr() + (row_off * Kw + off) * 4
        else:
            ent = w16.lookup(weight) if off % 8 == 0 else None
            if ent is not None:   # operand copy refreshed once per step by gt_cast_multi (begin_step)
                w, ldw = ent
                wptr = w.data_ptr() + (row_off * ldw + off) * 2
            else:  # bf16 operand copy of the fp32 master weight block, K padded so rows stay 16-B aligned
                w = torch.empty(N, ld_in, dtype=x.dtype, device=x.device)
                call("gt_cast_pad", GT_F32, wf.data_ptr() + (row_off * Kw + off) * 4, N, K, Kw, dt_of(w), ptr(w), N, ld_in, ld_in)
                ldw, wptr = ld_in, w.data_ptr()
        ld_out = ldp(N)
        out_dtype = torch.float32 if out_f32 else x.dtype
        y = torch.empty(M, ld_out, dtype=out_dtype, device=x.device)
        flags = (EPI_RELU if relu else 0) | (EPI_OUT_F32 if out_f32 and x.dtype != torch.float32 else 0)
        if resid is not None:
            resid = resid.contiguous()
            if resid.dtype != out_dtype or resid.shape != y.shape:
                raise RuntimeError("linear: resid must match the output")
            if out_dtype == torch.float32 and x.dtype != torch.float32:
                flags |= _lib.EPI_RESID_F32
        stats = zeros_small(2 * ld_out, torch.float64, x.device) if want_stats else None
        _gemm_raw(dt_of(x), x.data_ptr(), 0, ld_in, wptr, 0, ldw, y.data_ptr(), ld_out, M, N, K, ld_out, bias, resid,
                  ld_out, flags, drop_p=drop_p, rng=ptr(rng_state(x.device)) if drop_p else None, salt=salt,
                  col_stats=stats, m_valid=m_valid)
        ctx.drop_p = drop_p
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.params = (weight, bias_param)
        ctx.woff = wptr - w.data_ptr()      # byte offset of the operand block inside the saved weight tensor
        ctx.meta = (M, N, K, Kw, off, ld_in, ldw, ld_out, relu, bias is not None, resid is not None)
        ctx.row_off = int(row_off)
        # set by ops.batch_norm when a TRAIN-mode BatchNorm actually consumes this output: its backward returns a gradient
        # whose column sums vanish identically, i.e. the bias gradient of this Linear is exactly zero - no gt_colsum launch.
        # The consumer opts in (a frozen BN, a standalone conv or any other reader keeps the real column sum).
        ctx.bias_grad_zero = False
        ctx.passthrough = bool(passthrough)
        # no zero tensors for outputs without a gradient (the non-differentiable statistics output would otherwise cost
        # one fill kernel per Linear + BatchNorm pair in every backward)
        ctx.set_materialize_grads(False)
        if want_stats:
            ctx.mark_non_differentiable(stats)
            return y, stats
        if passthrough:
            # second output = the input itself, for the caller's residual connection: its gradient comes back HERE and is
            # added in the dX GEMM's epilogue (fp32 accumulator + g_pass, one rounding) instead of by an autograd
            # accumulation kernel over [M, ld_in]
            return y, x_in.view_as(x_in)
        return y

    @staticmethod
    def backward(ctx, gy, _gstats=None):
        x, w, y = ctx.saved_tensors
        M, N, K, Kw, off, ld_in, ldw, ld_out, relu, has_bias, has_resid = ctx.meta
        g_pass = _gstats if ctx.passthrough else None
        if gy is None:      # only the pass-through output was used
            gy = torch.zeros(M, ld_out, dtype=y.dtype if y is not None else x.dtype, device=x.device)
        gy = gy.contiguous()
        g_res = gy if has_resid else None
        if gy.dtype != x.dtype:  # fp32 head logits: bring the gradient to the operand dtype
            g2 = torch.empty(M, ld_out, dtype=x.dtype, device=x.device)
            call("gt_cast_pad", dt_of(gy), ptr(gy), M, ld_out, ld_out, dt_of(g2), ptr(g2), M, ld_out, ld_out)
            gy = g2
        weight, bias = ctx.params
        bias_done = False
        if relu:
            if y.dtype != gy.dtype:
                raise RuntimeError("linear: relu with a widened output is not supported")
            gz = torch.empty_like(gy)
            vw = 16 // gy.element_size()
            # a dropped element has y == 0, so the ReLU test also applies the keep mask; only the 1/(1-p) scale is left
            if (has_bias and ctx.needs_input_grad[2] and not ctx.bias_grad_zero and ld_out % vw == 0 and M >= 4096
                    and _main_grad(bias) is not None):
                # bias gradient = column sums of the masked gradient: taken in the same pass (no separate gt_colsum read)
                tgt, _ = _grad_target(bias)
                call("gt_relu_bwd_colsum", dt_of(gy), ptr(gy), ptr(y), M, N, ld_out, ptr(gz), 1.0 / (1.0 - ctx.drop_p),
                     tgt.data_ptr() + ctx.row_off * 4)
                bias_done = True
            else:
                call("gt_relu_bwd", dt_of(gy), ptr(gy), ptr(y), gy.numel(), ptr(gz), 1.0 / (1.0 - ctx.drop_p))
            gy = gz
        wptr = w.data_ptr() + ctx.woff
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty(M, ld_in, dtype=x.dtype, device=x.device)
            if g_pass is not None:
                g_pass = g_pass.contiguous()
                if g_pass.dtype != x.dtype or g_pass.shape != gx.shape:
                    raise RuntimeError("linear: pass-through gradient must match the input")
            # dX[m,k] = sum_n dY[m,n] W[n,k]: B operand = W read "MN-major" (k contiguous)
            if g_pass is None and x.dtype == torch.bfloat16 and N >= 2048 and M * K <= 256 * 512:
                # a handful of output tiles over a very long contraction (the 5002-class Code2 heads): split-K into a
                # zeroed fp32 scratch (one CTA would otherwise walk ~80 k-blocks alone), then one cast
                g32 = zeros_small(M * ld_in, torch.float32, x.device).view(M, ld_in)
                _gemm_raw(dt_of(x), gy.data_ptr(), 0, ld_out, wptr, 1, ldw, g32.data_ptr(), ld_in, M, K, N, ld_in, None,
                          None, 0, EPI_ACCUM | EPI_OUT_F32)
                call("gt_cast_pad", GT_F32, ptr(g32), M, ld_in, ld_in, dt_of(gx), ptr(gx), M, ld_in, ld_in)
            else:
                _gemm_raw(dt_of(x), gy.data_ptr(), 0, ld_out, wptr, 1, ldw, gx.data_ptr(), ld_in, M, K, N, ld_in, None,
                          g_pass, ld_in, 0)
        # parameter gradients accumulated in place need no ordering with the rest of the backward: side stream
        side_ok = _main_grad(weight) is not None and (not has_bias or _main_grad(bias) is not None)
        with _WgradCtx(side_ok, gy, x):
            if ctx.needs_input_grad[1]:
                tgt, gw = _grad_target(weight)
                # dW[n,k] = sum_m dY[m,n] X[m,k]: both operands MN-major, split-K over the rows, accumulated in place
                _gemm_raw(dt_of(x), gy.data_ptr(), 1, ld_out, x.data_ptr(), 1, ld_in,
                          tgt.data_ptr() + (ctx.row_off * Kw + off) * 4, Kw, N, K, M, K, None, None, 0, EPI_ACCUM | EPI_OUT_F32)
                _grad_done(weight)
            if has_bias and ctx.needs_input_grad[2]:
                tgt, gb = _grad_target(bias)           # (a fresh zero tensor when there is no arena)
                if not ctx.bias_grad_zero and not bias_done:
                    call("gt_colsum", dt_of(gy), ptr(gy), M, N, ld_out, tgt.data_ptr() + ctx.row_off * 4)
                _grad_done(bias)
        return gx, gw, gb, None, None, g_res, None, None, None, None, None, None, None, None, None


def linear(x, weight, bias=None, relu=False, out_f32=False, resid=None, drop_p=0.0, w_col_off=0, K=None, col_stats=False,
           m_valid=None, w_row_off=0, n_out=None, passthrough=False):
    """drop(act(x W[:, off:off+K]^T + b)) [+ resid]; drop(relu(.)) runs in the GEMM epilogue (the FFN pattern of
    nn.TransformerEncoderLayer); dropout without ReLU / together with resid is not a reference pattern.
    col_stats=True: the output carries its BatchNorm column statistics (taken in the GEMM epilogue), which
    ops.batch_norm picks up instead of a separate gt_colstats pass."""
    fused = bool(drop_p) and relu and resid is None
    if col_stats and FUSE_COLSTATS and not drop_p and not out_f32:
        y, stats = _LinearFn.apply(x, weight, bias, relu, out_f32, resid, w_col_off, K, 0.0, 0, True, m_valid)
        y._gt_colstats = stats
        return y
    if passthrough:
        # -> (y, x'): x' is x for the residual connection of the caller, routed through this op so that the gradient of
        # the residual path joins dX inside the GEMM epilogue
        if (drop_p and not fused) or not x.requires_grad:
            return linear(x, weight, bias, relu, out_f32, resid, drop_p, w_col_off, K, False, m_valid, w_row_off, n_out), x
        return _LinearFn.apply(x, weight, bias, relu, out_f32, resid, w_col_off, K, float(drop_p) if fused else 0.0,
                               next_salt() if fused else 0, False, None, w_row_off, n_out, True)
    y = _LinearFn.apply(x, weight, bias, relu, out_f32, resid, w_col_off, K, float(drop_p) if fused else 0.0,
                        next_salt() if fused else 0, False, None, w_row_off, n_out)
    return dropout(y, drop_p) if (drop_p and not fused) else y


class _StackedHeadsFn(torch.autograd.Function):
    """All H equally shaped Linear heads as ONE contraction over the stacked bf16 operand copy of the registry:
    y [M, H * Np] fp32 (Np = ldp(N); head h owns columns [h * Np, h * Np + N), pad columns are zero).  Replaces the
    per-head loop of reference models/gnn_transformer.py:121-128.  Backward: one split-K dX over the H * Np contraction,
    per-head dW / db on the weight-gradient stream."""

    @staticmethod
    def forward(ctx, x, H, *wb):
        ws, bs = wb[:H], wb[H:]
        wst, bst, rp, ldw = w16.stack_lookup(ws[0])
        x = x.contiguous()
        M, ld_in = x.shape
        N, K = ws[0].shape
        y = torch.empty(M, H * rp, dtype=torch.float32, device=x.device)
        _gemm_raw(dt_of(x), x.data_ptr(), 0, ld_in, wst.data_ptr(), 0, ldw, y.data_ptr(), H * rp, M, H * rp, K, H * rp, bst,
                  None, 0, EPI_OUT_F32)
        ctx.save_for_backward(x, wst)
        ctx.params = (ws, bs)
        ctx.meta = (H, M, N, K, rp, ld_in, ldw)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, wst = ctx.saved_tensors
        ws, bs = ctx.params
        H, M, N, K, rp, ld_in, ldw = ctx.meta
        gy = gy.contiguous()
        if gy.dtype != x.dtype:
            g2 = torch.empty(M, H * rp, dtype=x.dtype, device=x.device)
            call("gt_cast_pad", dt_of(gy), ptr(gy), M, H * rp, H * rp, dt_of(g2), ptr(g2), M, H * rp, H * rp)
            gy = g2
        gx = None
        if ctx.needs_input_grad[0]:
            # dX[m, k] = sum over (head, class) of dY W: a single output tile over a 25 k-long contraction -> split-K
            g32 = zeros_small(M * ld_in, torch.float32, x.device).view(M, ld_in)
            _gemm_raw(dt_of(x), gy.data_ptr(), 0, H * rp, wst.data_ptr(), 1, ldw, g32.data_ptr(), ld_in, M, K, H * rp, ld_in,
                      None, None, 0, EPI_ACCUM | EPI_OUT_F32)
            gx = torch.empty(M, ld_in, dtype=x.dtype, device=x.device)
            call("gt_cast_pad", GT_F32, ptr(g32), M, ld_in, ld_in, dt_of(gx), ptr(gx), M, ld_in, ld_in)
        es = gy.element_size()
        side_ok = all(_main_grad(t) is not None for t in (*ws, *bs))
        gws, gbs = [], []
        with _WgradCtx(side_ok, gy, x):
            for h in range(H):
                tgt, gw = _grad_target(ws[h])
                _gemm_raw(dt_of(x), gy.data_ptr() + h * rp * es, 1, H * rp, x.data_ptr(), 1, ld_in, tgt.data_ptr(), K, N, K, M, K,
                          None, None, 0, EPI_ACCUM | EPI_OUT_F32)
                _grad_done(ws[h])
                gws.append(gw)
                tgt, gb = _grad_target(bs[h])
                lib_call_colsum(gy, h * rp, M, N, H * rp, tgt)
                _grad_done(bs[h])
                gbs.append(gb)
        return (gx, None, *gws, *gbs)


def stacked_heads(x, linears):
    """-> ([M, H * Np] fp32 logits of all heads, Np) or None when the stacked operand copy is not available (fp32 parity
    mode, heads not registered with the current step's registry)"""
    linears = list(linears)
    if x.dtype != torch.bfloat16 or w16.stack_lookup(linears[0].weight) is None:
        return None
    rp = w16.stack_lookup(linears[0].weight)[2]
    y = _StackedHeadsFn.apply(x, len(linears), *[l.weight for l in linears], *[l.bias for l in linears])
    return y, rp


class PredList(list):
    """list of per-head logits views that remembers the stacked [M, H * Np] buffer they live in (fused loss)"""
    stacked = None


# ----------------------------------------------------------------------------- aggregation
class _AggregateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan, conv, d, edge_kind, edge_attr, edge_w, edge_b, etype, table, self_param, tab_side=False):
        x = x.contiguous()
        N, ld = x.shape
        out = torch.empty_like(x)
        kdim = edge_w.shape[1] if edge_kind == EDGE_LINEAR else 0
        edge_w_param = edge_w
        if edge_kind == EDGE_LINEAR:
            edge_attr = edge_attr.contiguous()
            edge_w = edge_w.contiguous()
        if edge_kind == EDGE_TABLE:
            table = table.contiguous()
            if table.shape[1] != ld or table.dtype != torch.float32:
                raise RuntimeError("edge table must be fp32 [ntypes, ld]")
        sp = self_param.contiguous().view(-1)
        slots = plan.edge_slots(conv, edge_kind, edge_attr if edge_kind == EDGE_LINEAR else None, etype)
        call("gt_aggregate_fwd", dt_of(x), conv, ptr(x), ptr(out), N, d, ld, ptr(plan.rowptr_dst),
             ptr(plan.src_by_dst), ptr(plan.eid_by_dst), ptr(plan.rowptr_src), edge_kind, ptr(edge_attr), kdim,
             ptr(edge_w), ptr(edge_b), ptr(etype), ptr(table), ptr(sp), ptr(slots[0][0]), ptr(slots[0][1]),
             ptr(slots[0][2]))
        ctx.slots = slots[1]
        ctx.split = edge_kind == EDGE_TABLE and table.shape[0] <= 1024
        ctx.tab_side = bool(tab_side)
        # bf16: the table gradient is the contraction OneHot(type)^T . gm over the per-edge gradients the adjoint writes
        ctx.tab_fused = bool(ctx.split and TABLE_GRAD_FUSED and x.dtype == torch.bfloat16 and ld in (128, 256)
                             and table.shape[0] <= 64 and (conv != CONV_GCN or slots[1][0] is not None)
                             and not os.environ.get("GT_AGG_VARIANT"))
        ctx.tab_gemm = bool(ctx.split and TABLE_GRAD_GEMM and x.dtype == torch.bfloat16 and plan.E >= 1024 and ld % 8 == 0
                            and ld <= 512 and slots[1][1] is not None and not ctx.tab_fused)
        if ctx.tab_fused:
            pass
        elif ctx.tab_gemm:
            plan.type_onehot(slots[1][1], table.shape[0])
        elif ctx.split:   # type-sorted edges for the table-gradient kernel (once per batch)
            plan.edges_by_type(plan._edge_index, etype, table.shape[0])
        ctx.save_for_backward(x, edge_attr, edge_w, edge_b, etype, table, sp)
        ctx.params = (edge_w_param, edge_b, self_param)
        ctx.meta = (plan, conv, d, edge_kind, kdim, self_param.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        x, edge_attr, edge_w, edge_b, etype, table, sp = ctx.saved_tensors
        plan, conv, d, edge_kind, kdim, sp_shape = ctx.meta
        g = g.contiguous()
        N, ld = x.shape
        dx = torch.empty_like(x)
        pw, pb, pself = ctx.params
        tw = tb = (None, None)
        if edge_kind == EDGE_LINEAR:
            tw, tb = _grad_target(pw), _grad_target(pb)
        need_tab = edge_kind == EDGE_TABLE and ctx.needs_input_grad[9]     # a constant table has no gradient kernel
        dtab = plan.zero_table_grad(tuple(table.shape), x.device) if need_tab else None
        # the edge-table gradient is a leaf gradient: computed by its own kernel over the type-sorted edges, on the
        # weight-gradient stream, instead of shared-memory atomics inside the adjoint (which then stays as cheap as
        # the forward)
        fused = ctx.tab_fused and need_tab      # table gradient accumulated by the adjoint kernel itself (d_table given)
        split = ctx.split and need_tab and not fused
        gm = torch.empty(max(plan.E, 1), ld, dtype=x.dtype, device=x.device) if (split and ctx.tab_gemm) else None
        tself = _grad_target(pself)
        # GIN + edge table in bf16: work buffer for the packed-mask adjoint (bf16 ReLU thresholds of this layer's table)
        th = None
        if (AGG_PACKED and edge_kind == EDGE_TABLE and conv == CONV_GIN and x.dtype == torch.bfloat16
                and (split or not need_tab) and ld % 8 == 0 and ld <= 512):
            th = torch.empty(table.shape[0] + 1, ld, dtype=torch.bfloat16, device=x.device)
        call("gt_aggregate_bwd", dt_of(x), conv, ptr(x), ptr(g), ptr(dx), N, d, ld, ptr(plan.rowptr_dst),
             ptr(plan.rowptr_src), ptr(plan.dst_by_src), ptr(plan.eid_by_src), edge_kind, ptr(edge_attr), kdim,
             ptr(edge_w), ptr(edge_b), ptr(etype), ptr(table), table.shape[0] if edge_kind == EDGE_TABLE else 0, ptr(sp),
             ptr(tw[0]), ptr(tb[0]), None if (split or not need_tab) else ptr(dtab), ptr(tself[0]), ptr(ctx.slots[0]), ptr(ctx.slots[1]),
             ptr(ctx.slots[2]), ptr(gm), ptr(th))
        if th is not None:
            _lib.kernel_count += 1          # the threshold builder in front of the packed-mask adjoint
        if split and gm is not None:
            oh, r_pad, _ = plan.type_onehot(ctx.slots[1], table.shape[0])
            with _WgradCtx(ctx.tab_side, gm, dtab, oh):
                # d_table[t, :] += sum over the slots of type t of gm[slot, :]   (split-K over the edges)
                _gemm_raw(GT_BF16, oh.data_ptr(), 1, r_pad, gm.data_ptr(), 1, ld, dtab.data_ptr(), ld, table.shape[0], d, plan.E, d,
                          None, None, 0, EPI_ACCUM | EPI_OUT_F32)
        elif split:
            src_t, dst_t, type_t, _ = plan.edges_by_type(plan._edge_index, etype, table.shape[0])
            # on the side stream only when the consumer of dtab (the embed_sum backward of the table) runs there too
            with _WgradCtx(ctx.tab_side, x, g, dtab, table, src_t, dst_t, type_t, plan.rowptr_src):
                call("gt_aggregate_table_grad", dt_of(x), conv, ptr(x), ptr(g), N, d, ld, ptr(plan.rowptr_src), plan.E,
                     ptr(src_t), ptr(dst_t), ptr(type_t), ptr(table), table.shape[0], ptr(dtab))
        for prm in (pw, pb, pself):
            _grad_done(prm)
        return dx, None, None, None, None, None, tw[1], tb[1], None, dtab, tself[1], None


def aggregate(x, plan, conv, d, self_param, edge_kind=EDGE_NONE, edge_attr=None, edge_w=None, edge_b=None,
              etype=None, table=None):
    tab_side = edge_kind == EDGE_TABLE and bool(getattr(table, "_gt_leaf_side", False))
    return _AggregateFn.apply(x, plan, conv, d, edge_kind, edge_attr, edge_w, edge_b, etype, table, self_param, tab_side)


# ----------------------------------------------------------------------------- segment ops
class _SegmentSumFn(torch.autograd.Function):
    """global_add_pool: [N, ld] -> fp32 [B, ld]"""

    @staticmethod
    def forward(ctx, x, plan, init=None):
        x = x.contiguous()
        N, ld = x.shape
        out = torch.empty(plan.B, ld, dtype=torch.float32, device=x.device)
        if init is not None:
            init = init.contiguous()
        call("gt_segment_sum_sorted", dt_of(x), ptr(x), ptr(plan.node_off), plan.B, ld, ptr(init), ptr(out))
        ctx.meta = (plan, x.dtype, N, ld)
        ctx.has_init = init is not None
        return out

    @staticmethod
    def backward(ctx, g):
        plan, dtype, N, ld = ctx.meta
        g = g.contiguous()
        dx = torch.empty(N, ld, dtype=dtype, device=g.device)
        call("gt_add_graph_vec", dt_of(dx), None, ptr(g), ptr(plan.node_graph), N, ld, ptr(dx))
        return dx, None, (g if ctx.has_init else None)


def segment_sum(x, plan, init=None):
    """out[g] = init[g] + sum_{i in g} x[i]  (global_add_pool; fp32 [B, ld])"""
    return _SegmentSumFn.apply(x, plan, init)


class _AddGraphVecFn(torch.autograd.Function):
    """y[i] = x[i] + v[graph(i)], v fp32 [B, ld]"""

    @staticmethod
    def forward(ctx, x, v, plan):
        x = x.contiguous()
        v = v.contiguous()
        N, ld = x.shape
        y = torch.empty_like(x)
        call("gt_add_graph_vec", dt_of(x), ptr(x), ptr(v), ptr(plan.node_graph), N, ld, ptr(y))
        ctx.meta = (plan, N, ld)
        return y

    @staticmethod
    def backward(ctx, g):
        plan, N, ld = ctx.meta
        g = g.contiguous()
        dv = None
        if ctx.needs_input_grad[1]:
            dv = torch.empty(plan.B, ld, dtype=torch.float32, device=g.device)
            call("gt_segment_sum_sorted", dt_of(g), ptr(g), ptr(plan.node_off), plan.B, ld, None, ptr(dv))
        return g, dv, None


def add_graph_vec(x, v, plan):
    return _AddGraphVecFn.apply(x, v, plan)


class _BroadcastRowFn(torch.autograd.Function):
    """fp32 [rows, ld] filled with the (zero-padded) single row of a [1, d] parameter (the initial virtual-node state,
    reference modules/gnn_module.py:188-189); backward = column sum straight into the parameter gradient"""

    @staticmethod
    def forward(ctx, w, rows, ld):
        d = w.shape[1]
        wc = w.contiguous()
        out = torch.empty(rows, ld, dtype=torch.float32, device=w.device)
        call("gt_cast_pad", GT_F32, ptr(wc), rows, d, 0, GT_F32, ptr(out), rows, ld, ld)   # source row stride 0
        ctx.param, ctx.meta = w, (rows, d, ld)
        return out

    @staticmethod
    def backward(ctx, g):
        rows, d, ld = ctx.meta
        g = g.contiguous()
        tgt, gw = _grad_target(ctx.param)
        call("gt_colsum", dt_of(g), ptr(g), rows, d, ld, ptr(tgt))
        _grad_done(ctx.param)
        return gw, None, None


def broadcast_row(w, rows, ld):
    return _BroadcastRowFn.apply(w, rows, ld)


# ----------------------------------------------------------------------------- BatchNorm1d
class _BatchNormFn(torch.autograd.Function):
    """y = act(BN(x)) [+ resid] [+ gvec[graph]] on physical [M, ld]; train mode uses batch stats
    and updates the running buffers in place exactly like nn.BatchNorm1d."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, training, momentum, eps, relu, resid,
                gvec, plan, drop_p, salt, pre_stats=None, m_valid=None):
        x = x.contiguous()
        M, ld = x.shape
        d = gamma.shape[0]
        dev = x.device
        ssmr = torch.empty(4 * ld, dtype=torch.float32, device=dev)
        stats = None
        if training:
            if pre_stats is not None and pre_stats.numel() == 2 * ld:   # taken in the producing GEMM's epilogue
                stats = pre_stats
            else:
                stats = zeros_small(2 * ld, torch.float64, dev)
                call("gt_colstats", dt_of(x), ptr(x), M, ld, ptr(stats), ptr(m_valid))
        y = torch.empty_like(x)
        if resid is not None:
            resid = resid.contiguous()
        if gvec is not None:
            gvec = gvec.contiguous()
        rng = ptr(rng_state(dev)) if drop_p else None
        call("gt_bn_norm_fwd", dt_of(x), ptr(x), M, d, ld, ptr(stats), ptr(gamma), ptr(beta), ptr(running_mean),
             ptr(running_var), ptr(nbt), float(momentum), float(eps), int(training), int(relu), ptr(resid), ptr(gvec),
             ptr(plan.node_graph) if gvec is not None else None, ptr(y), ptr(ssmr), float(drop_p), rng, salt, ptr(m_valid))
        ctx.m_valid = m_valid
        ctx.save_for_backward(x, ssmr, gamma)
        ctx.params = (gamma, beta)
        ctx.meta = (M, d, ld, relu, training, plan, resid is not None, gvec is not None, float(drop_p), salt)
        return y

    @staticmethod
    def backward(ctx, g):
        x, ssmr, gamma = ctx.saved_tensors
        M, d, ld, relu, training, plan, has_resid, has_gvec, drop_p, salt = ctx.meta
        g = g.contiguous()
        dev = x.device
        rng = ptr(rng_state(dev)) if drop_p else None
        red = zeros_small(2 * ld, torch.float64, dev)
        mv = ptr(ctx.m_valid)
        call("gt_bn_bwd_reduce", dt_of(x), ptr(x), ptr(g), M, d, ld, ptr(ssmr), int(relu), ptr(red), drop_p, rng,
             salt, mv)
        dx = torch.empty_like(x)
        pg, pb = ctx.params
        (tg, dgamma), (tb, dbeta) = _grad_target(pg), _grad_target(pb)
        call("gt_bn_bwd_apply", dt_of(x), ptr(x), ptr(g), M, d, ld, ptr(ssmr), ptr(gamma), int(relu),
             int(training), ptr(red), ptr(dx), ptr(tg), ptr(tb), drop_p, rng, salt, mv)
        _grad_done(pg)
        _grad_done(pb)
        dres = g if has_resid else None
        dgv = None
        if has_gvec and ctx.needs_input_grad[11]:
            dgv = torch.empty(plan.B, ld, dtype=torch.float32, device=dev)
            call("gt_segment_sum_sorted", dt_of(g), ptr(g), ptr(plan.node_off), plan.B, ld, None, ptr(dgv))
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, dres, dgv, None, None, None, None, None


def batch_norm(x, bn: torch.nn.BatchNorm1d, relu=False, resid=None, gvec=None, plan=None, drop_p=0.0, m_valid=None):
    """drop(act(BN(x))) [+ resid] [+ gvec[graph]] in one kernel; an input produced by ops.linear(col_stats=True) brings
    its column statistics along"""
    training = bn.training or bn.running_mean is None
    pre = getattr(x, "_gt_colstats", None) if (training and x.is_contiguous()) else None
    if pre is not None and SKIP_ZERO_BIAS_GRAD and hasattr(x.grad_fn, "bias_grad_zero"):
        x.grad_fn.bias_grad_zero = True      # producer = _LinearFn whose output goes straight into this train-mode BN
    return _BatchNormFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                              training, bn.momentum, bn.eps, relu, resid, gvec, plan, drop_p,
                              next_salt() if drop_p else 0, pre, m_valid)


# eval path: a BatchNorm that directly follows a Linear is, with running statistics, an affine map
# of the Linear's output - folded into the weight / bias once per evaluation phase, the layer is ONE contraction with
# a ReLU epilogue instead of contraction + normalise pass.  y = s * (x W^T + b) + t, s = gamma * rsqrt(var + eps),
# t = beta - mean * s  ->  W' = diag(s) W, b' = s * b + t.
EVAL_FOLD_BN = int(os.environ.get("GT_EVAL_FOLD_BN", "1"))


def fold_bn(lin: torch.nn.Linear, bn: torch.nn.BatchNorm1d):
    """-> (W', b') of Linear followed by eval-mode BatchNorm, or None when folding does not apply (training, autograd on,
    no running statistics).  Cached until the next training forward; bf16 mode keeps an operand copy in the registry."""
    if not EVAL_FOLD_BN or bn.training or torch.is_grad_enabled() or bn.running_mean is None or not lin.weight.is_cuda:
        return None
    ent = lin.__dict__.get("_gt_fold")        # lives (and dies) with the module
    # stale after any training forward (the fused optimizer writes through raw pointers) or any in-place update torch
    # tracks (load_state_dict of best_model.pt before the final eval, reference main.py:262-267)
    stamp = (_train_steps[0],) + tuple(t._version for t in (lin.weight, lin.bias, bn.weight, bn.bias, bn.running_mean,
                                                              bn.running_var) if t is not None)
    if ent is None or ent[0] != stamp or ent[3] is not bn or ent[1].device != lin.weight.device:
        with torch.no_grad():
            s = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            w = (lin.weight * s[:, None]).contiguous()
            b0 = lin.bias if lin.bias is not None else torch.zeros_like(s)
            b = ((b0 - bn.running_mean) * s + bn.bias).contiguous()
        ent = (stamp, w, b, bn)
        lin.__dict__["_gt_fold"] = ent
        if _PRECISION == "bf16":
            rows, cols = w.shape
            ld = ldp(cols)
            c = torch.empty(rows, ld, dtype=torch.bfloat16, device=w.device)
            call("gt_cast_pad", GT_F32, ptr(w), rows, cols, cols, GT_BF16, ptr(c), rows, ld, ld)
            w16.put(w, c, ld)
    elif _PRECISION == "bf16" and w16.lookup(ent[1]) is None:     # another registry became current since the fold
        w = ent[1]
        rows, cols = w.shape
        ld = ldp(cols)
        c = torch.empty(rows, ld, dtype=torch.bfloat16, device=w.device)
        call("gt_cast_pad", GT_F32, ptr(w), rows, cols, cols, GT_BF16, ptr(c), rows, ld, ld)
        w16.put(w, c, ld)
    return ent[1], ent[2]


# ----------------------------------------------------------------------------- LayerNorm / tokens
class _LayerNormFn(torch.autograd.Function):
    """y = LN(x[in_rows] (+cls for -1 rows) + resid). Rows = len(in_rows) if given else x rows."""

    @staticmethod
    def forward(ctx, x, resid, gamma, beta, eps, in_rows, cls, n_rows, drop_p, salt):
        x = x.contiguous()
        d = gamma.shape[0]
        if x.shape[1] != d:
            raise RuntimeError("layer_norm expects unpadded rows (d_model % 8 == 0)")
        M = n_rows if in_rows is not None else x.shape[0]
        dev = x.device
        y = torch.empty(M, d, dtype=x.dtype, device=dev)
        presum = torch.empty(M, d, dtype=x.dtype, device=dev) if (resid is not None or in_rows is not None or drop_p) else None
        mr = torch.empty(2 * M, dtype=torch.float32, device=dev)
        if resid is not None:
            resid = resid.contiguous()
        clsv = cls.contiguous().view(-1) if cls is not None else None
        call("gt_layernorm_fwd", dt_of(x), ptr(x), ptr(resid), ptr(in_rows), ptr(clsv), M, d, ptr(gamma), ptr(beta),
             float(eps), ptr(y), ptr(presum), ptr(mr), float(drop_p), ptr(rng_state(dev)) if drop_p else None, salt)
        ctx.drop = (float(drop_p), salt)
        ctx.save_for_backward(presum if presum is not None else x, mr, gamma)
        ctx.params = (gamma, beta, cls)
        ctx.meta = (M, d, in_rows, x.shape[0], resid is not None, cls.shape if cls is not None else None)
        return y

    @staticmethod
    def backward(ctx, g):
        presum, mr, gamma = ctx.saved_tensors
        M, d, rows, x_rows, has_resid, cls_shape = ctx.meta
        g = g.contiguous()
        dev = g.device
        if rows is not None:
            dx = torch.zeros(x_rows, d, dtype=g.dtype, device=dev)
        else:
            dx = torch.empty(M, d, dtype=g.dtype, device=dev)
        pg, pb, pcls = ctx.params
        (tg, dgamma), (tb, dbeta) = _grad_target(pg), _grad_target(pb)
        tcls, dcls = _grad_target(pcls) if cls_shape is not None else (None, None)
        drop_p, salt = ctx.drop
        dxd = torch.empty_like(dx) if drop_p else None     # gradient of the dropped operand (same mask as forward)
        call("gt_layernorm_bwd", dt_of(g), ptr(g), ptr(presum), ptr(mr), ptr(rows), M, d, ptr(gamma), ptr(dx),
             ptr(tg), ptr(tb), ptr(tcls), ptr(dxd), drop_p, ptr(rng_state(dev)) if drop_p else None, salt)
        for prm in (pg, pb, pcls):
            if prm is not None:
                _grad_done(prm)
        return (dxd if drop_p else dx, dx if has_resid else None, dgamma, dbeta, None, None, dcls, None, None, None)


def layer_norm(x, ln: torch.nn.LayerNorm, resid=None, in_rows=None, cls=None, n_rows=None, drop_p=0.0):
    """LN(drop(x) + resid) - the dropout runs inside the LayerNorm kernels (no gather together with dropout)"""
    if drop_p and in_rows is not None:
        x, drop_p = dropout(x, drop_p), 0.0
    return _LayerNormFn.apply(x, resid, ln.weight, ln.bias, ln.eps, in_rows, cls, n_rows, float(drop_p),
                              next_salt() if drop_p else 0)


class _GatherRowsFn(torch.autograd.Function):
    """dst[r] = src[rows[r]] (rows -1 -> cls vector, -2 -> zeros)"""

    @staticmethod
    def forward(ctx, src, rows, cls, n_rows):
        src = src.contiguous()
        ld = src.shape[1]
        dst = torch.empty(n_rows, ld, dtype=src.dtype, device=src.device)
        clsv = cls.contiguous().view(-1) if cls is not None else None
        call("gt_gather_rows", dt_of(src), ptr(src), ptr(rows), ptr(clsv), n_rows, ld, ptr(dst))
        ctx.meta = (rows, src.shape[0], ld, n_rows, cls.shape if cls is not None else None)
        ctx.cls = cls
        return dst

    @staticmethod
    def backward(ctx, g):
        rows, src_rows, ld, n_rows, cls_shape = ctx.meta
        g = g.contiguous()
        dsrc = torch.zeros(src_rows, ld, dtype=g.dtype, device=g.device)
        tcls, dcls = _grad_target(ctx.cls) if cls_shape is not None else (None, None)
        call("gt_scatter_rows", dt_of(g), ptr(g), ptr(rows), n_rows, ld, ptr(dsrc), ptr(tcls))
        if cls_shape is not None:
            _grad_done(ctx.cls)
        return dsrc, None, dcls, None


def gather_rows(src, rows, cls=None, n_rows=None):
    return _GatherRowsFn.apply(src, rows, cls, n_rows if n_rows is not None else rows.numel())


class _PadBatchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, plan, S):
        h = h.contiguous()
        N, ld = h.shape
        padded = torch.empty(S, plan.B, ld, dtype=h.dtype, device=h.device)
        mask = torch.empty(plan.B, S, dtype=torch.uint8, device=h.device)
        call("gt_pad_batch_fwd", dt_of(h), ptr(h), ptr(plan.node_off), plan.B, S, ld, ptr(padded), ptr(mask))
        ctx.meta = (plan, S, N, ld)
        ctx.mark_non_differentiable(mask)
        return padded, mask

    @staticmethod
    def backward(ctx, g, _gm):
        plan, S, N, ld = ctx.meta
        g = g.contiguous()
        dh = torch.empty(N, ld, dtype=g.dtype, device=g.device)
        call("gt_pad_batch_bwd", dt_of(g), ptr(g), ptr(plan.node_off), ptr(plan.node_graph), plan.B, S, N, ld,
             ptr(dh))
        return dh, None, None


def pad_batch_dense(h, plan, S):
    return _PadBatchFn.apply(h, plan, S)


# ----------------------------------------------------------------------------- attention
class _MHAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, plan, nhead, key_start, drop_p, salt, impl=None):
        impl = MHA_IMPL if impl is None else impl
        qkv = qkv.contiguous()
        n_rows, d3 = qkv.shape
        d = d3 // 3
        dh = d // nhead
        scale = float(dh) ** -0.5
        lse = torch.empty(nhead * n_rows, dtype=torch.float32, device=qkv.device)
        meta = (getattr(plan, "row_bounds", None), getattr(plan, "tile_bounds", None)) if key_start is None else (None, None)
        # every graph inside one graph-aligned 128-row tile: loop-free tile-local kernels (attn_local.cu)
        local = (impl == 0 and key_start is None and getattr(plan, "loc_tiles", None) is not None
                 and qkv.dtype == torch.bfloat16 and dh in (32, 64))
        # (the tile-local kernels clear the rows that belong to no graph themselves: no zero-fill of the output)
        out = torch.empty(n_rows, d, dtype=qkv.dtype, device=qkv.device)
        if local:
            call("gt_mha_local_fwd", dt_of(qkv), ptr(qkv), ptr(plan.row_bounds), ptr(plan.loc_tiles), ptr(plan.loc_count),
                 plan.loc_max_tiles,
                 n_rows, nhead, dh, scale, ptr(out), ptr(lse), float(drop_p),
                 ptr(rng_state(qkv.device)) if drop_p else None, salt)
        else:
            call("gt_mha_fwd", dt_of(qkv), ptr(qkv), ptr(plan.tok_graph), ptr(plan.tok_off), ptr(key_start), ptr(meta[0]),
                 ptr(meta[1]), n_rows,
                 plan.B, nhead, dh, scale, ptr(out), ptr(lse), float(drop_p),
                 ptr(rng_state(qkv.device)) if drop_p else None, salt, impl)
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (plan, nhead, dh, scale, key_start, float(drop_p), salt, impl, local)
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, out, lse = ctx.saved_tensors
        plan, nhead, dh, scale, key_start, drop_p, salt, impl, local = ctx.meta
        g = g.contiguous()
        n_rows = qkv.shape[0]
        dqkv = torch.empty_like(qkv)
        if local:
            call("gt_mha_local_bwd", dt_of(qkv), ptr(qkv), ptr(out), ptr(g), ptr(lse), ptr(plan.row_bounds),
                 ptr(plan.loc_tiles), plan.loc_max_tiles, n_rows, nhead, dh, scale, ptr(dqkv), drop_p,
                 ptr(rng_state(qkv.device)) if drop_p else None, salt)
            return dqkv, None, None, None, None, None, None
        delta = torch.empty(nhead * n_rows, dtype=torch.float32, device=qkv.device)
        meta = (getattr(plan, "row_bounds", None), getattr(plan, "tile_bounds", None)) if key_start is None else (None, None)
        call("gt_mha_bwd", dt_of(qkv), ptr(qkv), ptr(out), ptr(g), ptr(lse), ptr(plan.tok_graph), ptr(plan.tok_off),
             ptr(key_start), ptr(meta[0]), ptr(meta[1]), n_rows, plan.B, nhead, dh, scale, ptr(dqkv), ptr(delta), drop_p,
             ptr(rng_state(qkv.device)) if drop_p else None, salt, impl)
        return dqkv, None, None, None, None, None, None


def mha_packed(qkv, plan, nhead, key_start=None, drop_p=0.0):
    """`plan` needs .tok_graph, .tok_off and .B (GraphPlan or any object with those fields)."""
    return _MHAFn.apply(qkv, plan, nhead, key_start, drop_p, next_salt() if drop_p else 0, None)


class _MHAClsFn(torch.autograd.Function):
    """attention of ONE query per graph (the pooled row) over all token rows of the graph: q [B, d], kv [n_rows, 2d]"""

    @staticmethod
    def forward(ctx, q, kv, plan, nhead, drop_p, salt):
        q, kv = q.contiguous(), kv.contiguous()
        B, d = q.shape
        n_rows = kv.shape[0]
        dh = d // nhead
        scale = float(dh) ** -0.5
        out = torch.empty(B, d, dtype=q.dtype, device=q.device)
        lse = torch.empty(B * nhead, dtype=torch.float32, device=q.device)
        call("gt_mha_cls_fwd", dt_of(q), ptr(q), ptr(kv), ptr(plan.tok_off), ptr(plan.cls_rows), n_rows, B, nhead, dh, scale,
             ptr(out), ptr(lse), float(drop_p), ptr(rng_state(q.device)) if drop_p else None, salt)
        ctx.save_for_backward(q, kv, out, lse)
        ctx.meta = (plan, nhead, dh, scale, float(drop_p), salt)
        return out

    @staticmethod
    def backward(ctx, g):
        q, kv, out, lse = ctx.saved_tensors
        plan, nhead, dh, scale, drop_p, salt = ctx.meta
        g = g.contiguous()
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        call("gt_mha_cls_bwd", dt_of(q), ptr(q), ptr(kv), ptr(out), ptr(g), ptr(lse), ptr(plan.tok_off), ptr(plan.cls_rows),
             kv.shape[0], q.shape[0], nhead, dh, scale, ptr(dq), ptr(dkv), drop_p,
             ptr(rng_state(q.device)) if drop_p else None, salt)
        return dq, dkv, None, None, None, None


def mha_pooled_query(q, kv, plan, nhead, drop_p=0.0):
    """last encoder layer: only the pooled row of every graph is a query (reference models/gnn_transformer.py:114-115
    reads `transformer_out[-1]` only); keys / values are all token rows of the graph"""
    return _MHAClsFn.apply(q, kv, plan, nhead, drop_p, next_salt() if drop_p else 0)


# ----------------------------------------------------------------------------- graph read-outs (baseline models)
class _SegmentPoolFn(torch.autograd.Function):
    """global_mean_pool / global_max_pool: [N, ld] -> fp32 [B, ld]"""

    @staticmethod
    def forward(ctx, x, plan, mode):
        x = x.contiguous()
        N, ld = x.shape
        out = torch.empty(plan.B, ld, dtype=torch.float32, device=x.device)
        arg = torch.empty(plan.B, ld, dtype=torch.int32, device=x.device) if mode == 2 else None
        call("gt_segment_pool_fwd", dt_of(x), mode, ptr(x), ptr(plan.node_off), plan.B, ld, ptr(out), ptr(arg))
        ctx.meta = (plan, mode, N, ld, x.dtype)
        ctx.arg = arg
        return out

    @staticmethod
    def backward(ctx, g):
        plan, mode, N, ld, dtype = ctx.meta
        g = g.contiguous().float()
        dx = torch.empty(N, ld, dtype=dtype, device=g.device)
        call("gt_segment_pool_bwd", dt_of(dx), mode, ptr(g), ptr(plan.node_off), ptr(plan.node_graph), ptr(ctx.arg), N, ld,
             ptr(dx))
        return dx, None, None


def segment_pool(x, plan, kind):
    """kind in ('sum', 'mean', 'max'): PyG global_add_pool / global_mean_pool / global_max_pool over physical [N, ld]"""
    if kind == "sum":
        return segment_sum(x, plan)
    return _SegmentPoolFn.apply(x, plan, {"mean": 1, "max": 2}[kind])


def argmax_rows(x, n_cols=None):
    """eval read-out: int64 [rows] index of the first maximum of the first n_cols columns of fp32 x"""
    if x.dtype != torch.float32 or x.stride(-1) != 1:
        x = x.float().contiguous()
    rows, width = x.shape
    out = torch.empty(rows, dtype=torch.int64, device=x.device)
    call("gt_argmax_rows", ptr(x), rows, int(n_cols or width), x.stride(0), ptr(out))
    return out


# ----------------------------------------------------------------------------- PNA
class _TowerLinearFn(torch.autograd.Function):
    """Block-diagonal ("tower") linear: y[:, t*Fo:(t+1)*Fo] = x[:, xo_t : xo_t+K] W_t[:, wo:wo+K]^T (+ b_t),
    xo_t = t*x_tower_stride.  One gt_gemm per tower on strided views (no copies).  Used for the PNA
    pre-MLP halves (W_i / W_j of pre_nns[t].0) and post-MLP (post_nns[t].0)."""

    @staticmethod
    def forward(ctx, x, x_tower_stride, K, w_col_off, use_bias, n_towers, *wb):
        ws, bs = wb[:n_towers], wb[n_towers:]
        x = x.contiguous()
        M, ldx = x.shape
        Fo = ws[0].shape[0]
        ld_out = ldp(n_towers * Fo)
        y = torch.empty(M, ld_out, dtype=x.dtype, device=x.device)
        if ld_out > n_towers * Fo:
            y[:, n_towers * Fo:].zero_()
        es = x.element_size()
        wop = []
        for t in range(n_towers):
            w = ws[t].contiguous()
            if x.dtype != torch.float32:   # low-precision operand copy of the fp32 master weight
                wl = torch.empty(w.shape, dtype=x.dtype, device=x.device)
                call("gt_cast_pad", GT_F32, ptr(w), w.shape[0], w.shape[1], w.shape[1], dt_of(wl), ptr(wl),
                     w.shape[0], w.shape[1], w.shape[1])
                w = wl
            wop.append(w)
            _gemm_raw(dt_of(x), x.data_ptr() + t * x_tower_stride * es, 0, ldx,
                      w.data_ptr() + w_col_off * es, 0, w.shape[1], y.data_ptr() + t * Fo * es, ld_out, M, Fo, K, Fo,
                      bs[t] if use_bias else None, None, 0, 0)
        ctx.save_for_backward(x, *wop)
        ctx.params = (ws, bs)
        ctx.meta = (x_tower_stride, K, w_col_off, use_bias, n_towers, Fo, ld_out, [w.shape for w in ws])
        return y

    @staticmethod
    def backward(ctx, gy):
        x, *wop = ctx.saved_tensors
        xs, K, wo, use_bias, T, Fo, ld_out, wshapes = ctx.meta
        gy = gy.contiguous()
        M, ldx = x.shape
        es = x.element_size()
        gx = torch.zeros_like(x)
        gws, gbs = [], []
        pws, pbs = ctx.params
        for t in range(T):
            w = wop[t]
            gyp = gy.data_ptr() + t * Fo * es
            # dX_t[m,k] = sum_n dY_t[m,n] W_t[n, wo+k]
            _gemm_raw(dt_of(x), gyp, 0, ld_out, w.data_ptr() + wo * es, 1, w.shape[1],
                      gx.data_ptr() + t * xs * es, ldx, M, K, Fo, K, None, None, 0, 0)
            tgt, gw = _grad_target(pws[t])
            _gemm_raw(dt_of(x), gyp, 1, ld_out, x.data_ptr() + t * xs * es, 1, ldx,
                      tgt.data_ptr() + wo * 4, wshapes[t][1], Fo, K, M, K, None, None, 0, EPI_ACCUM | EPI_OUT_F32)
            _grad_done(pws[t])
            gws.append(gw)
            if use_bias:
                tgt, gb = _grad_target(pbs[t])
                lib_call_colsum(gy, t * Fo, M, Fo, ld_out, tgt)
                _grad_done(pbs[t])
                gbs.append(gb)
            else:
                gbs.append(None)
        return (gx, None, None, None, None, None, *gws, *gbs)


class _BlockDiagLinearFn(torch.autograd.Function):
    """y = x . BD^T (+ b): the T towers of a PNA stage as ONE tensor-core contraction over the block-diagonal bf16 operand
    the registry keeps (W16Registry.register_blockdiag; off-diagonal zero blocks only cost MMA slots - a tower slice
    of width F = 68 is not 16-byte aligned, the whole matrix is).  Backward: dX = dY . BD (one contraction), the
    operand gradient dBD = dY^T X is taken whole into a zeroed fp32 scratch (split-K) and its diagonal blocks are
    added to the per-tower weight gradients by gt_add_blocks; bias gradients are column sums of dY."""

    @staticmethod
    def forward(ctx, x, key, col_lo, K, T, use_bias, *wb):
        ws, bs = wb[:T], wb[T:]
        op, ldk, bias = w16.bd_lookup(key)
        x = x.contiguous()
        M, ldx = x.shape
        Fo = ws[0].shape[0]
        N = T * Fo
        ld_out = ldp(N)
        y = torch.empty(M, ld_out, dtype=x.dtype, device=x.device)
        _gemm_raw(dt_of(x), x.data_ptr(), 0, ldx, op.data_ptr(), 0, ldk, y.data_ptr(), ld_out, M, N, T * K, ld_out,
                  bias if use_bias else None, None, 0, 0)
        ctx.save_for_backward(x, op)
        ctx.params = (ws, bs)
        ctx.meta = (col_lo, K, T, Fo, use_bias, ldk, ld_out)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, op = ctx.saved_tensors
        ws, bs = ctx.params
        col_lo, K, T, Fo, use_bias, ldk, ld_out = ctx.meta
        gy = gy.contiguous()
        M, ldx = x.shape
        N = T * Fo
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty(M, ldx, dtype=x.dtype, device=x.device)
            _gemm_raw(dt_of(x), gy.data_ptr(), 0, ld_out, op.data_ptr(), 1, ldk, gx.data_ptr(), ldx, M, T * K, N, ldx, None,
                      None, 0, 0)
        side_ok = all(_main_grad(t) is not None for t in ws) and (not use_bias or all(_main_grad(t) is not None for t in bs))
        gws, gbs = [], []
        with _WgradCtx(side_ok, gy, x):
            ldt = ldp(T * K)
            temp = zeros_small(N * ldt, torch.float32, x.device)
            _gemm_raw(dt_of(x), gy.data_ptr(), 1, ld_out, x.data_ptr(), 1, ldx, temp.data_ptr(), ldt, N, T * K, M, T * K, None,
                      None, 0, EPI_ACCUM | EPI_OUT_F32)
            tg = [_grad_target(w) for w in ws]
            a_dst = (ctypes.c_void_p * T)(*[t[0].data_ptr() + col_lo * 4 for t in tg])
            a_ld = (ctypes.c_int32 * T)(*[w.shape[1] for w in ws])
            a_r0 = (ctypes.c_int32 * T)(*[t * Fo for t in range(T)])
            a_c0 = (ctypes.c_int32 * T)(*[t * K for t in range(T)])
            a_rows = (ctypes.c_int32 * T)(*[Fo] * T)
            a_cols = (ctypes.c_int32 * T)(*[K] * T)
            call("gt_add_blocks", ptr(temp), ldt, T, a_dst, a_ld, a_r0, a_c0, a_rows, a_cols)
            for t in range(T):
                _grad_done(ws[t])
                gws.append(tg[t][1])
            if use_bias:
                for t in range(T):
                    tgt, gb = _grad_target(bs[t])
                    lib_call_colsum(gy, t * Fo, M, Fo, ld_out, tgt)
                    _grad_done(bs[t])
                    gbs.append(gb)
            else:
                gbs = [None] * T
        return (gx, None, None, None, None, None, *gws, *gbs)


def blockdiag_linear(x, key, weights, biases, col_lo, K):
    """tower Linears through the registry's block-diagonal operand `key`; None when it is not available (fp32 parity
    mode, registry of another model current)"""
    if x.dtype != torch.bfloat16 or w16.bd_lookup(key) is None:
        return None
    use_bias = biases is not None
    bs = list(biases) if use_bias else [None] * len(weights)
    return _BlockDiagLinearFn.apply(x, key, col_lo, K, len(weights), use_bias, *weights, *bs)


def lib_call_colsum(g, col_off, M, n, ld, out):
    call("gt_colsum", dt_of(g), g.data_ptr() + col_off * g.element_size(), M, n, ld, ptr(out))


def tower_linear(x, weights, biases, x_tower_stride, K, w_col_off=0):
    use_bias = biases is not None
    bs = list(biases) if use_bias else [None] * len(weights)
    return _TowerLinearFn.apply(x, x_tower_stride, K, w_col_off, use_bias, len(weights), *weights, *bs)


class _PNAReduceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pj, pi, plan, towers, F, delta):
        x, pj, pi = x.contiguous(), pj.contiguous(), pi.contiguous()
        N, ld = pj.shape
        ld_out = towers * 13 * F
        out = torch.empty(N, ld_out, dtype=pj.dtype, device=pj.device)
        amax = torch.empty(N, ld, dtype=torch.int32, device=pj.device)
        amin = torch.empty(N, ld, dtype=torch.int32, device=pj.device)
        call("gt_pna_reduce_fwd", dt_of(pj), ptr(x), ptr(pj), ptr(pi), N, towers, F, ld, ptr(plan.rowptr_dst),
             ptr(plan.src_by_dst), float(delta), ptr(out), ld_out, ptr(amax), ptr(amin))
        ctx.save_for_backward(pj, out, amax, amin)
        ctx.meta = (plan, towers, F, float(delta), ld_out)
        return out

    @staticmethod
    def backward(ctx, g):
        pj, out, amax, amin = ctx.saved_tensors
        plan, towers, F, delta, ld_out = ctx.meta
        g = g.contiguous()
        N, ld = pj.shape
        dpj = torch.zeros(N, ld, dtype=torch.float32, device=pj.device)
        dpi = torch.empty_like(pj)
        dx = torch.empty_like(pj)
        call("gt_pna_reduce_bwd", dt_of(pj), ptr(pj), ptr(out), ptr(g), N, towers, F, ld, ld_out,
             ptr(plan.rowptr_dst), ptr(plan.src_by_dst), delta, ptr(amax), ptr(amin), ptr(dpj), ptr(dpi), ptr(dx))
        return dx, cast_to(dpj, pj.dtype), dpi, None, None, None, None


def pna_reduce(x, pj, pi, plan, towers, F, delta):
    """[N, towers*13F]: per tower [x_t | scaled (mean,max,min,std) x 3 scalers] (post-MLP operand)"""
    return _PNAReduceFn.apply(x, pj, pi, plan, towers, F, delta)


# ----------------------------------------------------------------------------- fused losses
class _BCEMaskedFn(torch.autograd.Function):
    """mean over labelled (non-NaN) entries of BCE-with-logits (reference dataset/mol.py:24-31), device-resident"""

    @staticmethod
    def forward(ctx, x, y):
        if x.dtype != torch.float32 or x.stride(-1) != 1:
            x = x.float().contiguous()
        y = y.to(torch.float32)
        if y.stride(-1) != 1:
            y = y.contiguous()
        rows, cols = x.shape
        acc = torch.zeros(3, dtype=torch.float32, device=x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        call("gt_bce_masked_fwd", ptr(x), ptr(y), rows, cols, x.stride(0), y.stride(0), ptr(acc), ptr(loss))
        ctx.save_for_backward(x, y, acc)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, y, acc = ctx.saved_tensors
        rows, cols = x.shape
        dx = torch.empty(rows, cols, dtype=torch.float32, device=x.device)
        g = g.contiguous()
        call("gt_bce_masked_bwd", ptr(x), ptr(y), rows, cols, x.stride(0), y.stride(0), ptr(acc), ptr(g), ptr(dx), cols, cols)
        return dx, None


def bce_with_logits_masked_mean(pred, y):
    return _BCEMaskedFn.apply(pred, y)


class _CEFn(torch.autograd.Function):
    """mean cross-entropy over rows (reference dataset/code.py:39-45 per head, dataset/tud.py:25-27); n_cols: the first
    n_cols columns of every row are the classes, the rest is layout padding (its gradient is written as zeros)"""

    @staticmethod
    def forward(ctx, x, target, n_cols):
        if x.dtype != torch.float32 or x.stride(-1) != 1:
            x = x.float().contiguous()
        rows, width = x.shape
        cols = width if n_cols is None else int(n_cols)
        acc = torch.zeros(3, dtype=torch.float32, device=x.device)
        lse = torch.empty(rows, dtype=torch.float32, device=x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        call("gt_ce_fwd", ptr(x), ptr(target), target.stride(0), rows, cols, x.stride(0), ptr(lse), ptr(acc), ptr(loss))
        ctx.save_for_backward(x, target, lse)
        ctx.cols = cols
        return loss

    @staticmethod
    def backward(ctx, g):
        x, target, lse = ctx.saved_tensors
        rows, width = x.shape
        dx = torch.empty(rows, width, dtype=torch.float32, device=x.device)
        g = g.contiguous()
        call("gt_ce_bwd", ptr(x), ptr(target), target.stride(0), rows, ctx.cols, x.stride(0), ptr(lse), ptr(g), ptr(dx), width, width)
        return dx, None, None


def cross_entropy_mean(pred, target, n_cols=None):
    """pred [rows, classes] fp32 (row-strided views are fine), target int64 [rows] (any stride); n_cols < pred.shape[1]
    marks trailing padding columns"""
    return _CEFn.apply(pred, target, n_cols)
