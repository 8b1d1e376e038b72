"""GNNTransformer with the reference's constructor, flags, run name, state_dict keys and
forward signature (reference models/gnn_transformer.py:16-146).  The forward never builds the
padded [S, B, d] tensor: node states go straight from the GNN into PACKED token rows (graph i
owns rows [tok_off[i], tok_off[i+1]), last row = <CLS>), which is what the reference's
left-padded layout + -inf key mask computes for every row anything reads."""
import logging

import torch
import torch.nn as nn

from .. import ops
from ..modules.gnn_module import GNNNodeEmbedding
from ..modules.masked_transformer_encoder import MaskedOnlyTransformerEncoder
from ..modules.transformer_encoder import TransformerNodeEncoder
from .base_model import BaseModel

logger = logging.getLogger(__name__)


def _gnn_node_state(state_dict, module_name="gnn_node"):
    new_state_dict = dict()
    for k, v in state_dict.items():
        if module_name in k:
            new_key = k.split(".")
            module_index = new_key.index(module_name)
            new_state_dict[".".join(new_key[module_index + 1:])] = v
    return new_state_dict


class _GraphTransBase(BaseModel):
    """shared tail: gnn2transformer -> packed transformer -> pooled row -> head(s)"""

    def _init_tail(self, num_tasks, args):
        if args.pretrained_gnn:
            state_dict = self._gnn_node_state(torch.load(args.pretrained_gnn)["model"])
            logger.info("Load GNN state from: %s", state_dict.keys())
            self.gnn_node.load_state_dict(state_dict)
        self.freeze_gnn = args.freeze_gnn
        gnn_emb_dim = 2 * args.gnn_emb_dim if args.gnn_JK == "cat" else args.gnn_emb_dim
        self._gnn_dim = args.gnn_emb_dim
        self.gnn2transformer = nn.Linear(gnn_emb_dim, args.d_model)
        # JK=cat applies the weight once per part (two gradient contributions per step, ops._grad_done)
        self.gnn2transformer.weight._gt_uses = 2 if args.gnn_JK == "cat" else 1
        self.transformer_encoder = TransformerNodeEncoder(args)
        self.num_encoder_layers = args.num_encoder_layers
        if self.num_encoder_layers < 1:
            raise NotImplementedError("num_encoder_layers == 0 (GNN-only ablation) is out of scope, SURVEY §8")
        self.num_tasks = num_tasks
        self.pooling = args.graph_pooling
        if self.pooling not in ("cls", "last"):
            # "mean" divides by the number of PAD positions in the reference (gnn_transformer.py:117),
            # no shipped GraphTrans config uses it
            raise NotImplementedError(f"graph_pooling={self.pooling}")
        self.graph_pred_linear_list = torch.nn.ModuleList()
        self.max_seq_len = args.max_seq_len
        if args.max_seq_len is None:
            self.graph_pred_linear = torch.nn.Linear(args.d_model, self.num_tasks)
        else:
            for _ in range(args.max_seq_len):
                self.graph_pred_linear_list.append(torch.nn.Linear(args.d_model, self.num_tasks))
        self._w16 = ops.W16Registry()
        if len(self.graph_pred_linear_list) > 1:      # the 5 x 5002-class Code2 heads: one stacked operand
            self._w16.register_heads(self.graph_pred_linear_list)
        self._w16.register(self)   # bf16 operand copies of every Linear / in_proj weight, refreshed once per step
        for m in self.modules():
            if hasattr(m, "register_operands"):
                m.register_operands(self._w16)

    def _gnn2transformer(self, parts):
        """Linear over the logical concatenation of the JK parts without materialising the concat:
        part p multiplies the weight columns [p*d_g, (p+1)*d_g) and accumulates."""
        w, b = self.gnn2transformer.weight, self.gnn2transformer.bias
        d_g = self._gnn_dim
        h = None
        for p, part in enumerate(parts):   # part p multiplies weight columns [p*d_g, (p+1)*d_g)
            h = ops.linear(part, w, b if p == 0 else None, resid=h, w_col_off=p * d_g, K=d_g)
        return h

    def forward(self, batched_data, perturb=None):
        ops._lib.require_cuda(batched_data.batch, batched_data.edge_index)
        dev = batched_data.batch.device
        side = None
        if self.training:
            ops.begin_step(dev, cast_now=False, registry=self._w16)
            if ops.precision() == "bf16":       # bf16 weight copies refreshed next to the plan build
                side = lambda: self._w16.refresh(dev)  # noqa: E731
        elif ops.precision() == "bf16":
            ops.w16 = self._w16
            side = lambda: self._w16.refresh(dev)  # noqa: E731
        enc = self.transformer_encoder
        plan = ops.plan_for(batched_data, enc.max_input_len, cls=self.pooling == "cls", side_work=side)
        parts = self.gnn_node.forward_parts(batched_data, perturb, plan=plan)
        h_node = self._gnn2transformer(parts)                          # [N, d_model]
        h_graph = enc.forward_packed(h_node, plan)                       # [B, d_model] (pooled rows only)
        if self.max_seq_len is None:
            w, b = self.graph_pred_linear.weight, self.graph_pred_linear.bias
            return ops.linear(h_graph, w, b, out_f32=True)[:, :self.num_tasks]
        st = ops.stacked_heads(h_graph, self.graph_pred_linear_list)
        if st is not None:     # all heads in one contraction; per-head views for the caller, stacked buffer for the loss
            y, rp = st
            v = y.view(y.shape[0], len(self.graph_pred_linear_list), rp)
            preds = ops.PredList(v[:, h, :self.num_tasks] for h in range(v.shape[1]))
            preds.stacked = (y, rp, self.num_tasks)
            return preds
        return [ops.linear(h_graph, l.weight, l.bias, out_f32=True)[:, :self.num_tasks]
                for l in self.graph_pred_linear_list]

    def epoch_callback(self, epoch):
        if self.freeze_gnn is not None and epoch >= self.freeze_gnn:
            logger.info("Freeze GNN weight after epoch: %d", epoch)
            for param in self.gnn_node.parameters():
                param.requires_grad = False

    def _gnn_node_state(self, state_dict):
        return _gnn_node_state(state_dict)


class GNNTransformer(_GraphTransBase):
    @staticmethod
    def get_emb_dim(args):
        return args.gnn_emb_dim

    @staticmethod
    def add_args(parser):
        TransformerNodeEncoder.add_args(parser)
        MaskedOnlyTransformerEncoder.add_args(parser)
        group = parser.add_argument_group("GNNTransformer - Training Config")
        group.add_argument("--pos_encoder", default=False, action="store_true")
        group.add_argument("--pretrained_gnn", type=str, default=None, help="pretrained gnn_node node embedding path")
        group.add_argument("--freeze_gnn", type=int, default=None, help="Freeze gnn_node weight from epoch `freeze_gnn`")

    @staticmethod
    def name(args):
        name = f"{args.model_type}-pooling={args.graph_pooling}"
        name += "-norm_input" if args.transformer_norm_input else ""
        name += f"+{args.gnn_type}"
        name += "-virtual" if args.gnn_virtual_node else ""
        name += f"-JK={args.gnn_JK}"
        name += f"-enc_layer={args.num_encoder_layers}"
        name += f"-enc_layer_masked={args.num_encoder_layers_masked}"
        name += f"-d={args.d_model}"
        name += f"-act={args.transformer_activation}"
        name += f"-tdrop={args.transformer_dropout}"
        name += f"-gdrop={args.gnn_dropout}"
        name += "-pretrained_gnn" if args.pretrained_gnn else ""
        name += f"-freeze_gnn={args.freeze_gnn}" if args.freeze_gnn is not None else ""
        name += "-prenorm" if args.transformer_prenorm else "-postnorm"
        return name

    def __init__(self, num_tasks, node_encoder, edge_encoder_cls, args):
        super().__init__()
        self.gnn_node = GNNNodeEmbedding(
            args.gnn_virtual_node, args.gnn_num_layer, args.gnn_emb_dim, node_encoder, edge_encoder_cls,
            JK=args.gnn_JK, drop_ratio=args.gnn_dropout, residual=args.gnn_residual, gnn_type=args.gnn_type)
        if args.pos_encoder:
            raise NotImplementedError("pos_encoder (NCI ablation configs only) is out of scope, SURVEY §8")
        self.pos_encoder = None
        self.masked_transformer_encoder = MaskedOnlyTransformerEncoder(args)
        self.num_encoder_layers_masked = args.num_encoder_layers_masked
        self._init_tail(num_tasks, args)
