"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain torch ops, fp32 or fp64 following the dtype of the state dict) of the
GraphTrans forward pass of ucbrise/graphtrans for the north-star hot path; the backward is
torch autograd over these ops, exactly as in the reference (trainers/base_trainer.py:33).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product package never does.

Parity pin: the reference has NO tests or golden vectors for this path, so this
restatement is pinned against outputs of the UNMODIFIED reference itself, executed in the build
container through oracle/stubs (oracle/gen_golden.py -> tests/golden/*.pt) and checked by
tests/test_oracle_golden.py.  Third-party arithmetic that is not under /root/reference
(torch-geometric 1.6.3 MessagePassing/degree/global_add_pool/PNAConv, torch-scatter 2.0.6
scatter, torch 1.7.1 nn.TransformerEncoderLayer, ogb 1.2.6 encoders; requirement.yml:37,75,
96-98) is restated from its published behaviour and anchored on the
reference's call sites cited per function below.

The state dict uses the reference's own keys, so a checkpoint of the
reference model (or of the product model) evaluates here directly.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
LN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ------------------------------------------------------------------ primitives
def scatter_sum(src, index, n):
    """torch_scatter.scatter(reduce='sum') (Appendix A.3)."""
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def degree(index, n, dtype):
    """PyG degree (Appendix A.2): zeros(N).scatter_add_(0, index, ones)."""
    return torch.zeros(n, dtype=dtype, device=index.device).index_add_(
        0, index, torch.ones(index.numel(), dtype=dtype, device=index.device))


class _Ctx:
    """Carries train/eval mode and collects the BatchNorm running-stat updates."""

    def __init__(self, sd, training=True):
        self.sd, self.training, self.new_buffers = sd, training, {}


def batchnorm(ctx, x, prefix):
    """nn.BatchNorm1d: train mode = biased batch statistics over all rows, running stats updated
    with the unbiased variance and momentum 0.1 (reference modules/gnn_module.py:58,84,164,167)."""
    sd = ctx.sd
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if ctx.training:
        n = x.shape[0]
        mean = x.mean(0)
        var = ((x - mean) ** 2).mean(0)
        rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
        ctx.new_buffers[prefix + ".running_mean"] = ((1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean).detach()
        ctx.new_buffers[prefix + ".running_var"] = (
            (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))).detach()
        ctx.new_buffers[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    return (x - mean) / torch.sqrt(var + BN_EPS) * w + b


def layernorm(x, w, b):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def linear(sd, prefix, x):
    return x @ sd[prefix + ".weight"].t() + sd[prefix + ".bias"]


# ------------------------------------------------------------------ encoders (a2, a5)
def encode_nodes(sd, prefix, kind, batch):
    if kind == "code2":  # ASTNodeEncoder, reference dataset/utils.py:28-30 (depth clamped to max_depth)
        max_depth = sd[prefix + ".depth_encoder.weight"].shape[0] - 1
        depth = batch.node_depth.view(-1).clamp(max=max_depth)
        return (sd[prefix + ".type_encoder.weight"][batch.x[:, 0]]
                + sd[prefix + ".attribute_encoder.weight"][batch.x[:, 1]]
                + sd[prefix + ".depth_encoder.weight"][depth])
    if kind in ("mol", "syn"):  # ogb AtomEncoder (Appendix A.7), reference dataset/mol.py:83
        out = 0
        for c in range(batch.x.shape[1]):
            out = out + sd[f"{prefix}.atom_embedding_list.{c}.weight"][batch.x[:, c]]
        return out
    if kind == "nci1":  # nn.Linear(num_features, emb), reference dataset/tud.py:65
        return linear(sd, prefix, batch.x.to(sd[prefix + ".weight"].dtype))
    raise ValueError(kind)


def edge_embedding(sd, prefix, kind, edge_attr):
    if kind == "code2":  # nn.Linear(2, emb), reference dataset/code.py:117
        return linear(sd, prefix, edge_attr.to(sd[prefix + ".weight"].dtype))
    if kind in ("mol", "syn"):  # ogb BondEncoder, reference dataset/mol.py:84
        out = 0
        for c in range(edge_attr.shape[1]):
            out = out + sd[f"{prefix}.bond_embedding_list.{c}.weight"][edge_attr[:, c]]
        return out
    if kind == "nci1":  # python int 0, reference dataset/tud.py:67-71
        return 0
    raise ValueError(kind)


# ------------------------------------------------------------------ convs (a3, a4)
def gcn_conv(ctx, prefix, kind, x, edge_index, edge_attr):
    """reference modules/conv.py:50-68."""
    sd = ctx.sd
    n = x.shape[0]
    x = linear(sd, prefix + ".linear", x)
    ee = edge_embedding(sd, prefix + ".edge_encoder", kind, edge_attr)
    row, col = edge_index[0], edge_index[1]
    deg = degree(row, n, x.dtype) + 1  # OUT-degree of the source index + 1 (conv.py:57)
    dis = deg.pow(-0.5)
    norm = dis[row] * dis[col]
    msg = norm.view(-1, 1) * F.relu(x[row] + ee)
    agg = scatter_sum(msg, col, n)
    return agg + F.relu(x + sd[prefix + ".root_emb.weight"]) * 1.0 / deg.view(-1, 1)


def gin_conv(ctx, prefix, kind, x, edge_index, edge_attr):
    """reference modules/conv.py:26-33; mlp = Lin(d,2d)-BN-ReLU-Lin(2d,d) (conv.py:18-20)."""
    sd = ctx.sd
    ee = edge_embedding(sd, prefix + ".edge_encoder", kind, edge_attr)
    row, col = edge_index[0], edge_index[1]
    agg = scatter_sum(F.relu(x[row] + ee), col, x.shape[0])
    z = (1 + sd[prefix + ".eps"]) * x + agg
    z = linear(sd, prefix + ".mlp.0", z)
    z = F.relu(batchnorm(ctx, z, prefix + ".mlp.1"))
    return linear(sd, prefix + ".mlp.3", z)


# ------------------------------------------------------------------ GNN stacks (a6)
def gnn_node(ctx, args, batch, perturb=None, prefix="gnn_node"):
    """GNN_node / GNN_node_Virtualnode forward, reference modules/gnn_module.py:60-107,172-241.
    Dropout is the identity here (parity runs use p=0)."""
    sd, kind = ctx.sd, args.dataset
    L = args.gnn_num_layer
    conv = gcn_conv if args.gnn_type == "gcn" else gin_conv
    h0 = encode_nodes(sd, prefix + ".node_encoder", kind, batch)
    if perturb is not None:
        h0 = h0 + perturb
    h_list = [h0]
    bidx = batch.batch
    B = int(bidx[-1]) + 1
    virtual = args.gnn_virtual_node
    if virtual:
        vn = sd[prefix + ".virtualnode_embedding.weight"][torch.zeros(B, dtype=torch.long, device=bidx.device)]
    for layer in range(L):
        if virtual:
            h_list[layer] = h_list[layer] + vn[bidx]  # mutation seen by JK (gnn_module.py:199)
        h = conv(ctx, f"{prefix}.convs.{layer}", kind, h_list[layer], batch.edge_index, batch.edge_attr)
        h = batchnorm(ctx, h, f"{prefix}.batch_norms.{layer}")
        if layer != L - 1:
            h = F.relu(h)
        if args.gnn_residual:
            h = h + h_list[layer]
        h_list.append(h)
        if virtual and layer < L - 1:
            t = scatter_sum(h_list[layer], bidx, B) + vn  # global_add_pool (gnn_module.py:219)
            p = f"{prefix}.mlp_virtualnode_list.{layer}"
            t = F.relu(batchnorm(ctx, linear(sd, p + ".0", t), p + ".1"))
            t = F.relu(batchnorm(ctx, linear(sd, p + ".3", t), p + ".4"))
            vn = vn + t if args.gnn_residual else t
    if args.gnn_JK == "last":
        return h_list[-1]
    if args.gnn_JK == "sum":  # omits the final layer output (gnn_module.py:100-103)
        out = 0
        for layer in range(L):
            out = out + h_list[layer]
        return out
    if args.gnn_JK == "cat":
        return torch.cat([h_list[0], h_list[-1]], dim=-1)
    raise ValueError(args.gnn_JK)


# ------------------------------------------------------------------ PNA (a14, a15)
def pna_avg_deg_log(deg_hist):
    """reference modules/pna_layer.py:92-97: mean over histogram BINS of log(hist+1) (quirk a15)."""
    return (deg_hist.to(torch.float) + 1).log().mean().item()


def pna_conv(ctx, prefix, args, x, edge_index, towers=4):
    """PNAConv(towers=4, divide_input=True, edge_dim=None), reference modules/pna_layer.py:131-167,
    aggregators modules/pna/aggregators.py:11-34, scalers modules/pna/scalers.py:10-31."""
    sd = ctx.sd
    n, dg = x.shape
    Fd = dg // towers
    delta = pna_avg_deg_log(args.deg)
    xt = x.view(n, towers, Fd)
    src, dst = edge_index[0], edge_index[1]
    h = torch.cat([xt[dst], xt[src]], dim=-1)  # [x_i || x_j]
    m = torch.stack([linear(sd, f"{prefix}.pre_nns.{t}.0", h[:, t]) for t in range(towers)], dim=1)
    cnt = degree(dst, n, x.dtype)
    cdiv = cnt.clamp(min=1).view(-1, 1, 1)
    aggs = []
    for name in args.aggregators:
        if name == "mean":
            aggs.append(scatter_sum(m, dst, n) / cdiv)
        elif name in ("max", "min"):
            init = torch.zeros((n, towers, Fd), dtype=x.dtype, device=x.device)
            idx = dst.view(-1, 1, 1).expand_as(m)
            aggs.append(init.scatter_reduce(0, idx, m, "amax" if name == "max" else "amin", include_self=False))
        elif name == "std":
            mean = scatter_sum(m, dst, n) / cdiv
            msq = scatter_sum(m * m, dst, n) / cdiv
            aggs.append(torch.sqrt(F.relu(msq - mean * mean) + 1e-5))
        elif name == "sum":
            aggs.append(scatter_sum(m, dst, n))
        else:
            raise ValueError(name)
    out = torch.cat(aggs, dim=-1)
    d = cnt.view(-1, 1, 1)
    scaled = []
    for name in args.scalers:
        if name == "identity":
            scaled.append(out)
        elif name == "amplification":
            scaled.append(out * (torch.log(d + 1) / delta))
        elif name == "attenuation":
            s = delta / torch.log(d + 1)
            s = torch.where(d == 0, torch.ones_like(s), s)
            scaled.append(out * s)
        else:
            raise ValueError(name)
    out = torch.cat([xt] + scaled, dim=-1)
    outs = [linear(sd, f"{prefix}.post_nns.{t}.0", out[:, t]) for t in range(towers)]
    return linear(sd, prefix + ".lin", torch.cat(outs, dim=1))


def pna_node(ctx, args, batch, perturb=None, prefix="gnn_node", enc_prefix=None):
    """PNANodeEmbedding.forward, reference modules/pna/pna_module.py:57-78."""
    sd = ctx.sd
    x = encode_nodes(sd, enc_prefix or (prefix + ".node_encoder"), args.dataset, batch)
    if perturb is not None:
        x = x + perturb
    for layer in range(args.gnn_num_layer):
        h = pna_conv(ctx, f"{prefix}.layers.{layer}", args, x, batch.edge_index)
        h = F.relu(batchnorm(ctx, h, f"{prefix}.batch_norms.{layer}.module"))
        x = h + x if args.gnn_residual else x  # NB reference keeps x when not residual (pna_module.py:74-76)
    return x


# ------------------------------------------------------------------ pad_batch (a8)
def pad_plan(batch_idx, max_input_len):
    """Closed form of reference modules/utils.py:5-29 (integer work, bit-exact):
    n_i, S = min(max n_i, L), k_i = min(n_i, S); graph i occupies padded rows [S-k_i, S)."""
    B = int(batch_idx[-1]) + 1
    n = torch.bincount(batch_idx, minlength=B)
    off = torch.cumsum(n, 0) - n
    S = int(min(int(n.max()), int(max_input_len)))
    k = n.clamp(max=S)
    return n, off, k, S


def pad_batch(h, batch_idx, max_input_len):
    n, off, k, S = pad_plan(batch_idx, max_input_len)
    B, d = n.numel(), h.shape[-1]
    pos = torch.arange(S, device=h.device).view(S, 1)     # [S,1]
    valid = pos >= (S - k).view(1, B)                     # [S,B]
    src = (off + n - S).view(1, B) + pos                  # node index feeding padded[p, i]
    padded = torch.zeros(S, B, d, dtype=h.dtype, device=h.device)
    padded[valid] = h[src[valid]]
    mask = ~valid.t()                                     # [B,S] True = PAD
    return padded, mask.contiguous()


# ------------------------------------------------------------------ transformer (a9, a10)
def mha(sd, prefix, x, key_padding_mask, nhead):
    """torch nn.MultiheadAttention as called by nn.TransformerEncoderLayer (Appendix A.5)."""
    T, B, d = x.shape
    dh = d // nhead
    qkv = x @ sd[prefix + ".in_proj_weight"].t() + sd[prefix + ".in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)
    q = q * (dh ** -0.5)
    def heads(t):
        return t.reshape(T, B * nhead, dh).transpose(0, 1)  # [B*h, T, dh]
    q, k, v = heads(q), heads(k), heads(v)
    s = torch.bmm(q, k.transpose(1, 2)).view(B, nhead, T, T)
    s = s.masked_fill(key_padding_mask.view(B, 1, 1, T), float("-inf"))
    p = torch.softmax(s, dim=-1).view(B * nhead, T, T)
    o = torch.bmm(p, v).transpose(0, 1).reshape(T, B, d)
    return linear(sd, prefix + ".out_proj", o)


def transformer_encoder(sd, args, padded, mask, prefix="transformer_encoder"):
    """TransformerNodeEncoder.forward, reference modules/transformer_encoder.py:42-61, with the
    post-norm nn.TransformerEncoderLayer stack + final LayerNorm it builds at :28-32."""
    B = padded.shape[1]
    if args.graph_pooling == "cls":
        cls = sd[prefix + ".cls_embedding"].expand(1, B, -1)
        padded = torch.cat([padded, cls], dim=0)
        mask = torch.cat([mask, torch.zeros(B, 1, dtype=torch.bool, device=mask.device)], dim=1)
    x = padded
    if args.transformer_norm_input:
        x = layernorm(x, sd[prefix + ".norm_input.weight"], sd[prefix + ".norm_input.bias"])
    for l in range(args.num_encoder_layers):
        p = f"{prefix}.transformer.layers.{l}"
        a = mha(sd, p + ".self_attn", x, mask, args.nhead)
        x = layernorm(x + a, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
        f = linear(sd, p + ".linear2", F.relu(linear(sd, p + ".linear1", x)))
        x = layernorm(x + f, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
    x = layernorm(x, sd[prefix + ".transformer.norm.weight"], sd[prefix + ".transformer.norm.bias"])
    return x, mask


# ------------------------------------------------------------------ baseline models
def global_pool(h, bidx, B, kind):
    """PyG global_add_pool / global_mean_pool / global_max_pool = torch_scatter.scatter over `batch` (Appendix A.3:
    mean divides by clamp(count, 1); an empty segment of max gives 0)."""
    if kind == "sum":
        return scatter_sum(h, bidx, B)
    if kind == "mean":
        cnt = degree(bidx, B, h.dtype).clamp(min=1).view(-1, 1)
        return scatter_sum(h, bidx, B) / cnt
    if kind == "max":
        init = torch.zeros((B, h.shape[1]), dtype=h.dtype, device=h.device)
        return init.scatter_reduce(0, bidx.view(-1, 1).expand_as(h), h, "amax", include_self=False)
    raise NotImplementedError(kind)


def heads(sd, args, hg):
    if args.max_seq_len is None:
        return linear(sd, "graph_pred_linear", hg)
    return [linear(sd, f"graph_pred_linear_list.{i}", hg) for i in range(args.max_seq_len)]


def unpad_batch(out, prev, batch_idx, max_input_len):
    """reference modules/utils.py:32-53: the last k_i = min(n_i, S) nodes of graph i take padded rows [S - k_i, S) of
    `out` [S, B, d]; nodes truncated away keep `prev`."""
    n, off, k, S = pad_plan(batch_idx, max_input_len)
    N = prev.shape[0]
    g = batch_idx
    pos = torch.arange(N, device=prev.device) - off[g] - n[g] + S
    valid = pos >= 0
    src = out[pos.clamp(min=0), g]
    return torch.where(valid.view(-1, 1), src, prev)


def forward_baseline(ctx, sd, args, batch, perturb):
    """GNN.forward (reference models/gnn.py:100-115), PNANet.forward (models/pna.py:96-108) and Transformer.forward
    (models/transformer.py:86-115)."""
    bidx = batch.batch
    B = int(bidx[-1]) + 1
    if args.model_type == "gnn":
        hg = global_pool(gnn_node(ctx, args, batch, perturb), bidx, B, args.graph_pooling)
        return heads(sd, args, hg)
    if args.model_type == "pna":
        # PNANet registers the node encoder twice (models/pna.py:49,51): named_parameters() reports it as `node_encoder.*`
        hg = global_pool(pna_node(ctx, args, batch, perturb, prefix="pna_module", enc_prefix="node_encoder"), bidx, B,
                         args.graph_pooling)
        if args.max_seq_len is None:
            h = F.relu(linear(sd, "mlp.0", hg))
            h = F.relu(linear(sd, "mlp.2", h))
            return linear(sd, "mlp.4", h)
        return [linear(sd, f"graph_pred_linear_list.{i}.2", F.relu(linear(sd, f"graph_pred_linear_list.{i}.0", hg)))
                for i in range(args.max_seq_len)]
    if args.model_type == "transformer":
        tmp = encode_nodes(sd, "node_encoder", args.dataset, batch)
        if perturb is not None:
            tmp = tmp + perturb
        padded, mask = pad_batch(tmp, bidx, int(args.max_input_len))
        out, _ = transformer_encoder(sd, args, padded, mask, prefix="transformer")
        if args.graph_pooling == "cls":
            hg = out[-1]
        else:
            hg = global_pool(unpad_batch(out, tmp, bidx, int(args.max_input_len)), bidx, B, args.graph_pooling)
        return heads(sd, args, hg)
    raise ValueError(args.model_type)


# ------------------------------------------------------------------ model (a1, a7, a11, a12)
def forward(sd, args, batch, training=True, perturb=None):
    """GNNTransformer.forward (reference models/gnn_transformer.py:90-128) and
    PNATransformer.forward (reference models/pna_transformer.py:78-100).
    Returns (logits or list of logits, dict of updated BN buffers)."""
    ctx = _Ctx(sd, training)
    if args.model_type in ("gnn", "pna", "transformer"):
        return forward_baseline(ctx, sd, args, batch, perturb), ctx.new_buffers
    if args.model_type == "pna-transformer":
        h = pna_node(ctx, args, batch, perturb)
    else:
        h = gnn_node(ctx, args, batch, perturb)
    h = linear(sd, "gnn2transformer", h)
    padded, mask = pad_batch(h, batch.batch, int(args.max_input_len))
    out, _ = transformer_encoder(sd, args, padded, mask)
    if args.graph_pooling not in ("cls", "last"):
        raise NotImplementedError(args.graph_pooling)
    hg = out[-1]
    if args.max_seq_len is None:
        return linear(sd, "graph_pred_linear", hg), ctx.new_buffers
    return [linear(sd, f"graph_pred_linear_list.{i}", hg) for i in range(args.max_seq_len)], ctx.new_buffers


# ------------------------------------------------------------------ losses (a13)
def loss_fn(args, pred, batch):
    if args.dataset == "code2":  # reference dataset/code.py:39-45 (casts the logits to fp32, :42)
        loss = 0
        for i in range(len(pred)):
            loss = loss + F.cross_entropy(pred[i].to(torch.float32), batch.y_arr[:, i])
        return loss / len(pred)
    if args.dataset in ("mol", "syn"):  # reference dataset/mol.py:24-31 (fp32 cast at :27)
        labeled = batch.y == batch.y
        return F.binary_cross_entropy_with_logits(pred.to(torch.float32)[labeled],
                                                  batch.y.to(torch.float32)[labeled])
    if args.dataset == "nci1":  # reference dataset/tud.py:25-27
        return F.cross_entropy(pred, batch.y)
    raise ValueError(args.dataset)


def fwd_bwd(sd, args, batch, dtype=torch.float64, training=True):
    """One zero_grad -> forward -> loss -> backward. Returns logits, loss, grads, new buffers."""
    sd = {k: (v.detach().to(dtype).requires_grad_(True) if v.is_floating_point() else v.detach().clone())
          for k, v in sd.items()}
    pred, bufs = forward(sd, args, batch, training)
    loss = loss_fn(args, pred, batch)
    leaves = {k: v for k, v in sd.items() if v.is_floating_point() and "running_" not in k}
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(v)) for (k, v), g in zip(leaves.items(), grads)}
    return pred, loss.detach(), grads, bufs
