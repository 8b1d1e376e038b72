"""Transformer-only baseline with the reference's surface (reference models/transformer.py:20-115): node encoder ->
pad_batch -> TransformerNodeEncoder -> <CLS> row or unpad_batch + global pooling -> head(s).  Runs on packed tokens
like GraphTrans (the padded tensor is never built); with a node pooling the rows that truncation dropped keep the
encoder INPUT (reference modules/utils.py:32-53 unpad_batch)."""
import torch

from .. import ops
from ..modules.gnn_module import _encode
from ..modules.transformer_encoder import TransformerNodeEncoder
from . import _readout
from .base_model import BaseModel


class Transformer(BaseModel):
    @staticmethod
    def get_emb_dim(args):
        return args.d_model

    @staticmethod
    def add_args(parser):
        TransformerNodeEncoder.add_args(parser)

    @staticmethod
    def name(args):
        name = f"{args.model_type}-pooling={args.graph_pooling}"
        name += f"+{args.gnn_type}"
        name += "-virtual" if args.gnn_virtual_node else ""
        name += f"-d={args.d_model}"
        name += f"-tdp={args.transformer_dropout}"
        return name

    def __init__(self, num_tasks, node_encoder, edge_encoder_cls, args):
        super().__init__()
        self.transformer = TransformerNodeEncoder(args)
        self.node_encoder = node_encoder
        self.emb_dim = args.d_model
        self.num_tasks = num_tasks
        self.max_seq_len = args.max_seq_len
        self.graph_pooling = args.graph_pooling
        _readout.check_pooling(self.graph_pooling, extra=("cls",))
        _readout.make_heads(self, self.emb_dim, self.num_tasks, self.max_seq_len)
        self._w16 = ops.W16Registry()
        self._w16.register(self)

    def forward(self, batched_data, perturb=None):
        side = _readout.begin(self, batched_data, self._w16)
        enc = self.transformer
        cls = self.graph_pooling == "cls"
        plan = ops.plan_for(batched_data, enc.max_input_len, cls=cls, side_work=side)
        d, ld = self.emb_dim, ops.ldp(self.emb_dim)
        if isinstance(self.node_encoder, torch.nn.Linear):
            tmp = ops.linear(ops.pad_cols(batched_data.x, ops.ldp(batched_data.x.shape[1]), ops.act_dtype()),
                             self.node_encoder.weight, self.node_encoder.bias)
        else:
            tmp = _encode(self.node_encoder, batched_data)
            if tmp.shape[1] != ld or tmp.dtype != ops.act_dtype():
                tmp = ops.pad_cols(tmp, ld, ops.act_dtype())
        if perturb is not None:
            tmp = tmp + ops.pad_cols(perturb, ld, tmp.dtype)
        if cls:
            h_graph = enc.forward_packed(tmp, plan)
        else:
            h_tok = enc.forward_tokens(tmp, plan)                       # [n_rows, d], final norm applied
            max_nodes = getattr(batched_data, "max_nodes", None)
            rows = plan.node2tok
            if max_nodes is None or int(max_nodes) > enc.max_input_len:
                # truncated-away nodes (node2tok = -1) keep the encoder input (unpad_batch); -2 = "no source row" for
                # the gather / its scatter adjoint (-1 would mean the <CLS> vector)
                kept = (rows >= 0).unsqueeze(-1)
                h_node = torch.where(kept, ops.gather_rows(h_tok, torch.where(rows < 0, torch.full_like(rows, -2), rows)), tmp)
            else:
                h_node = ops.gather_rows(h_tok, rows)
            h_graph = _readout.pool_nodes(h_node, plan, self.graph_pooling)
        return _readout.apply_heads(self, h_graph, self.num_tasks, self.max_seq_len)
