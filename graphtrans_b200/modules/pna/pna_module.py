"""PNANodeEmbedding with the reference's flags, defaults and state_dict keys (reference
modules/pna/pna_module.py:16-78).  The layers are torch_geometric-1.6.3-style PNAConv parameter
containers (pre_nns / post_nns / lin, towers=4, divide_input=True, no edge features; the in-tree
statement of that arithmetic is reference modules/pna_layer.py:60-167); the math runs through
gt_gemm (per-node tower projections), gt_pna_reduce_* (one pass: mean/max/min/std + the three
degree scalers, no [E, towers, F] tensor) and gt_bn_*."""
import torch
import torch.nn as nn

from ... import ops

_AGGS = ["mean", "max", "min", "std"]
_SCALERS = ["identity", "amplification", "attenuation"]


class PNAConv(nn.Module):
    """Parameter container + forward of PNAConv(in, out, aggregators, scalers, deg, towers=4,
    divide_input=True) (reference modules/pna/pna_module.py:43-51)."""

    def __init__(self, in_channels, out_channels, aggregators, scalers, deg, towers=4, divide_input=True):
        super().__init__()
        if list(aggregators) != _AGGS or list(scalers) != _SCALERS:
            raise NotImplementedError("only the reference defaults aggregators=mean max min std, "
                                      "scalers=identity amplification attenuation are built")
        if not divide_input or in_channels != out_channels or in_channels % towers:
            raise NotImplementedError("PNAConv: needs divide_input and in == out divisible by towers")
        self.in_channels, self.out_channels, self.towers = in_channels, out_channels, towers
        self.F_in = self.F_out = in_channels // towers
        deg = deg.to(torch.float)
        # quirk preserved: means over histogram BINS (reference modules/pna_layer.py:92-97)
        self.avg_deg = {"lin": deg.mean().item(), "log": (deg + 1).log().mean().item(),
                        "exp": deg.exp().mean().item()}
        self.pre_nns = nn.ModuleList([nn.Sequential(nn.Linear(2 * self.F_in, self.F_in)) for _ in range(towers)])
        self.post_nns = nn.ModuleList([nn.Sequential(nn.Linear(13 * self.F_in, self.F_out)) for _ in range(towers)])
        self.lin = nn.Linear(out_channels, out_channels)
        for m in self.pre_nns:          # W_i and W_j halves are applied separately: two gradient contributions
            m[0].weight._gt_uses = 2

    def register_operands(self, registry):
        """block-diagonal bf16 tower operands (W_i | W_j halves of pre_nns, post_nns) in the model's registry: built once,
        refreshed with every other weight copy by the one gt_cast_multi launch of the step"""
        F = self.F_in
        pre_w, pre_b = [m[0].weight for m in self.pre_nns], [m[0].bias for m in self.pre_nns]
        post_w, post_b = [m[0].weight for m in self.post_nns], [m[0].bias for m in self.post_nns]
        self._bd_keys = (("pna_wi", id(self)), ("pna_wj", id(self)), ("pna_post", id(self)))
        registry.register_blockdiag(self._bd_keys[0], pre_w, pre_b, col_lo=0, K=F)
        registry.register_blockdiag(self._bd_keys[1], pre_w, None, col_lo=F, K=F)
        registry.register_blockdiag(self._bd_keys[2], post_w, post_b, col_lo=0, K=13 * F)

    def forward(self, x, edge_index=None, plan=None):
        """x: physical [N, ldp(d)] activation matrix; returns the same layout."""
        F, T = self.F_in, self.towers
        pre_w = [m[0].weight for m in self.pre_nns]
        pre_b = [m[0].bias for m in self.pre_nns]
        post_w = [m[0].weight for m in self.post_nns]
        post_b = [m[0].bias for m in self.post_nns]
        # W_pre [x_i || x_j] = W_i x_i + W_j x_j: project per NODE, then reduce over in-edges
        keys = getattr(self, "_bd_keys", None)
        pi = None
        if keys is not None and ops.precision() == "bf16" and (T * F) % 8 == 0 and (13 * F * T) % 8 == 0:
            # tensor-core path: the four towers as ONE block-diagonal contraction per stage over operand copies the
            # registry keeps (tower slices of width F are not 16-byte aligned for F = 68, whole matrices are)
            pi = ops.blockdiag_linear(x, keys[0], pre_w, pre_b, 0, F)
        if pi is not None:
            pj = ops.blockdiag_linear(x, keys[1], pre_w, None, F, F)
            agg = ops.pna_reduce(x, pj, pi, plan, T, F, self.avg_deg["log"])  # [N, T*13F]
            y = ops.blockdiag_linear(agg, keys[2], post_w, post_b, 0, 13 * F)
        else:
            pi = ops.tower_linear(x, pre_w, pre_b, F, F, w_col_off=0)
            pj = ops.tower_linear(x, pre_w, None, F, F, w_col_off=F)
            agg = ops.pna_reduce(x, pj, pi, plan, T, F, self.avg_deg["log"])  # [N, T*13F]
            y = ops.tower_linear(agg, post_w, post_b, 13 * F, 13 * F)
        return ops.linear(y, self.lin.weight, self.lin.bias)


class BatchNorm(nn.Module):
    """torch_geometric.nn.BatchNorm: wraps BatchNorm1d as `.module` (state_dict keys, Appendix B)."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)


class PNANodeEmbedding(nn.Module):
    @staticmethod
    def add_args(parser):
        group = parser.add_argument_group("PNANet configs")
        group.add_argument("--aggregators", type=str, nargs="+", default=["mean", "max", "min", "std"])
        group.add_argument("--scalers", type=str, nargs="+", default=["identity", "amplification", "attenuation"])
        group.add_argument("--post_layers", type=int, default=1)
        group.add_argument("--add_edge", type=str, default="none")
        group.set_defaults(gnn_residual=True)
        group.set_defaults(gnn_dropout=0.3)
        group.set_defaults(gnn_emb_dim=70)
        group.set_defaults(gnn_num_layer=4)

    def __init__(self, node_encoder, args):
        super().__init__()
        self.num_layer = args.gnn_num_layer
        self.max_seq_len = args.max_seq_len
        self.aggregators = args.aggregators
        self.scalers = args.scalers
        self.residual = args.gnn_residual
        self.drop_ratio = args.gnn_dropout
        self.graph_pooling = args.graph_pooling
        self.emb_dim = args.gnn_emb_dim
        self.node_encoder = node_encoder
        self.layers = nn.ModuleList([
            PNAConv(args.gnn_emb_dim, args.gnn_emb_dim, aggregators=self.aggregators, scalers=self.scalers,
                    deg=args.deg, towers=4, divide_input=True) for _ in range(self.num_layer)])
        self.batch_norms = nn.ModuleList([BatchNorm(args.gnn_emb_dim) for _ in range(self.num_layer)])

    def forward_parts(self, batched_data, perturb=None, plan=None):
        from ..gnn_module import _encode, _plan_of
        plan = plan or _plan_of(batched_data)
        d, ld = self.emb_dim, ops.ldp(self.emb_dim)
        x = _encode(self.node_encoder, batched_data)
        if x.shape[1] != ld or x.dtype != ops.act_dtype():
            x = ops.pad_cols(x, ld, ops.act_dtype())
        if perturb is not None:
            x = x + ops.pad_cols(perturb, ld, x.dtype)
        drop = self.drop_ratio if self.training else 0.0
        for conv, bn in zip(self.layers, self.batch_norms):
            h = conv(x, plan=plan)
            if self.residual:      # x = dropout(relu(BN(h)) + x)   (pna_module.py:73-76)
                x = ops.dropout(ops.batch_norm(h, bn.module, relu=True, resid=x, m_valid=plan.m_valid), drop)
            else:                  # reference keeps x (h is discarded); dropout still applies
                x = ops.dropout(x, drop)
        return [x]

    def forward(self, batched_data, perturb=None):
        return self.forward_parts(batched_data, perturb)[0][:, :self.emb_dim].float()
