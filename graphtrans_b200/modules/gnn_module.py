"""GNN stacks with the reference's API (reference modules/gnn_module.py:18-248): node encoder ->
L x (conv -> BatchNorm1d -> ReLU -> dropout [-> residual]) -> JK, optional virtual node.  The
layer loop is re-expressed over physical [N, ld] matrices; BN-apply, ReLU, residual and the
next layer's virtual-node broadcast are one kernel (gt_bn_norm_fwd); the statistics of a BatchNorm that follows a
Linear come out of that GEMM's epilogue (gt_gemm_stats)."""
import torch

from .. import ops
from .conv import GCNConv, GINConv, edge_encoder_operands


def _encode(node_encoder, batched_data):
    x = batched_data.x
    node_depth = batched_data.node_depth if hasattr(batched_data, "node_depth") else None
    if node_encoder is None:
        return x
    return node_encoder(x) if node_depth is None else node_encoder(x, node_depth.view(-1,))


def _plan_of(batched_data, max_input_len=1000):
    plan = getattr(batched_data, "_gt_plan", None)
    if plan is None or plan.L != int(max_input_len):
        plan = ops.plan_for(batched_data, max_input_len)
    return plan


class _GNNBase(torch.nn.Module):
    @staticmethod
    def need_deg():
        return False

    def _init_common(self, num_layer, emb_dim, node_encoder, edge_encoder_cls, drop_ratio, JK, residual, gnn_type):
        self.num_layer = num_layer
        self.emb_dim = emb_dim
        self.drop_ratio = drop_ratio
        self.JK = JK
        self.residual = residual
        if self.num_layer < 2:
            raise ValueError("Number of GNN layers must be greater than 1.")
        self.node_encoder = node_encoder
        self.convs = torch.nn.ModuleList()
        self.batch_norms = torch.nn.ModuleList()
        for _ in range(num_layer):
            if gnn_type == "gin":
                self.convs.append(GINConv(emb_dim, edge_encoder_cls))
            elif gnn_type == "gcn":
                self.convs.append(GCNConv(emb_dim, edge_encoder_cls))
            else:
                raise ValueError("Undefined GNN type called {}".format(gnn_type))
            self.batch_norms.append(torch.nn.BatchNorm1d(emb_dim))

    def _input(self, batched_data, perturb):
        d, ld = self.emb_dim, ops.ldp(self.emb_dim)
        if isinstance(self.node_encoder, torch.nn.Linear):     # TU datasets, reference dataset/tud.py:65
            h = ops.linear(ops.pad_cols(batched_data.x, ops.ldp(batched_data.x.shape[1]), ops.act_dtype()),
                           self.node_encoder.weight, self.node_encoder.bias)
        else:
            h = _encode(self.node_encoder, batched_data)
            if h.shape[1] != ld or h.dtype != ops.act_dtype():
                h = ops.pad_cols(h, ld, ops.act_dtype())
        if perturb is not None:                                   # FLAG, reference gnn_module.py:78,191
            h = h + ops.pad_cols(perturb, ld, h.dtype)
        return h

    def _jk(self, h_list):
        """returns (list of physical matrices to concatenate logically, logical width of each)"""
        if self.JK == "last":
            return [h_list[-1]]
        if self.JK == "sum":                                     # omits the final layer (gnn_module.py:100-103)
            out = h_list[0]
            for layer in range(1, self.num_layer):
                out = out + h_list[layer]
            return [out]
        if self.JK == "cat":
            return [h_list[0], h_list[-1]]
        raise ValueError(self.JK)

    def forward(self, batched_data, perturb=None):
        parts = self.forward_parts(batched_data, perturb)
        d = self.emb_dim
        return torch.cat([p[:, :d].float() for p in parts], dim=-1)


### GNN to generate node embedding
class GNN_node(_GNNBase):
    def __init__(self, num_layer, emb_dim, node_encoder, edge_encoder_cls, drop_ratio=0.5, JK="last", residual=False,
                 gnn_type="gin"):
        super().__init__()
        self._init_common(num_layer, emb_dim, node_encoder, edge_encoder_cls, drop_ratio, JK, residual, gnn_type)

    def forward_parts(self, batched_data, perturb=None, plan=None):
        plan = plan or _plan_of(batched_data)
        edge_index, edge_attr = batched_data.edge_index, batched_data.edge_attr
        h_list = [self._input(batched_data, perturb)]
        encs = edge_encoder_operands(self.convs, edge_attr, plan, self.emb_dim)
        for layer in range(self.num_layer):
            relu = layer != self.num_layer - 1
            kw = dict(enc=encs[layer])
            if isinstance(self.convs[layer], GINConv) and not self.residual:   # eval: BatchNorm folded into mlp[3]
                kw.update(out_bn=self.batch_norms[layer], out_relu=relu)
            h = self.convs[layer](h_list[layer], edge_index, edge_attr, plan=plan, **kw)
            if not getattr(h, "_gt_bn_folded", False):
                h = ops.batch_norm(h, self.batch_norms[layer], relu=relu,
                                   resid=h_list[layer] if self.residual else None,
                                   drop_p=self.drop_ratio if self.training else 0.0, m_valid=plan.m_valid)
            h_list.append(h)
        return self._jk(h_list)


### Virtual GNN to generate node embedding
class GNN_node_Virtualnode(_GNNBase):
    def __init__(self, num_layer, emb_dim, node_encoder, edge_encoder_cls, drop_ratio=0.5, JK="last", residual=False,
                 gnn_type="gin"):
        super().__init__()
        self._init_common(num_layer, emb_dim, node_encoder, edge_encoder_cls, drop_ratio, JK, residual, gnn_type)
        ### set the initial virtual node embedding to 0.
        self.virtualnode_embedding = torch.nn.Embedding(1, emb_dim)
        torch.nn.init.constant_(self.virtualnode_embedding.weight.data, 0)
        ### List of MLPs to transform virtual node at every layer
        self.mlp_virtualnode_list = torch.nn.ModuleList()
        for _ in range(num_layer - 1):
            self.mlp_virtualnode_list.append(torch.nn.Sequential(
                torch.nn.Linear(emb_dim, 2 * emb_dim), torch.nn.BatchNorm1d(2 * emb_dim), torch.nn.ReLU(),
                torch.nn.Linear(2 * emb_dim, emb_dim), torch.nn.BatchNorm1d(emb_dim), torch.nn.ReLU()))

    def forward_parts(self, batched_data, perturb=None, plan=None):
        plan = plan or _plan_of(batched_data)
        edge_index, edge_attr = batched_data.edge_index, batched_data.edge_attr
        d, ld = self.emb_dim, ops.ldp(self.emb_dim)
        h0 = self._input(batched_data, perturb)
        # per-graph virtual-node state, fp32 [B, ld]
        vn = ops.broadcast_row(self.virtualnode_embedding.weight, plan.B, ld)
        h_list = [ops.add_graph_vec(h0, vn, plan)]              # h + vn[batch]  (gnn_module.py:199)
        encs = edge_encoder_operands(self.convs, edge_attr, plan, self.emb_dim)
        drop = self.drop_ratio if self.training else 0.0
        for layer in range(self.num_layer):
            hv = h_list[layer]
            vn_next = None
            br = None
            if layer < self.num_layer - 1:                       # gnn_module.py:217-229
                mlp = self.mlp_virtualnode_list[layer]
                # independent of this layer's conv: runs as a parallel branch (ops.Branch) joined at the BN below
                br = ops.Branch(hv, vn)
                with br:
                    t = ops.segment_sum(hv, plan, init=vn)        # global_add_pool(h_list[layer]) + vn
                    t = ops.cast_to(t, ops.act_dtype())           # the MLP runs in the activation dtype
                    cs = self.training
                    f0, f3 = ops.fold_bn(mlp[0], mlp[1]), ops.fold_bn(mlp[3], mlp[4])
                    if f0 is not None and f3 is not None:          # eval: both Linear + BatchNorm + ReLU pairs folded
                        t = ops.linear(ops.linear(t, f0[0], f0[1], relu=True), f3[0], f3[1], relu=True)
                    else:
                        t = ops.batch_norm(ops.linear(t, mlp[0].weight, mlp[0].bias, col_stats=cs), mlp[1], relu=True)
                        t = ops.batch_norm(ops.linear(t, mlp[3].weight, mlp[3].bias, col_stats=cs), mlp[4], relu=True, drop_p=drop)
                    t = ops.cast_to(t, torch.float32)             # the virtual-node state itself stays fp32
                    vn_next = vn + t if self.residual else t
            h = self.convs[layer](hv, edge_index, edge_attr, plan=plan, enc=encs[layer])
            if br is not None:
                br.join(vn_next)
            # BN -> ReLU (not last) -> dropout -> (+residual) -> (+ next layer's vn[batch]) in one kernel
            h = ops.batch_norm(h, self.batch_norms[layer], relu=layer != self.num_layer - 1,
                               resid=hv if self.residual else None, gvec=vn_next, plan=plan, drop_p=drop,
                               m_valid=plan.m_valid)
            h_list.append(h)
            vn = vn_next
        return self._jk(h_list)


def GNNNodeEmbedding(virtual_node, *args, **kwargs):
    if virtual_node:
        return GNN_node_Virtualnode(*args, **kwargs)
    else:
        return GNN_node(*args, **kwargs)
