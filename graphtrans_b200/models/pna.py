"""PNA baseline with the reference's surface (reference models/pna.py:20-108): PNA stack -> global pooling -> MLP
head(s), on the same multi-aggregator reduce / GEMM / BatchNorm kernels as the PNA GraphTrans."""
import torch.nn as nn

from .. import ops
from ..modules.pna.pna_module import PNANodeEmbedding
from . import _readout
from .base_model import BaseModel


class PNANet(BaseModel):
    @staticmethod
    def get_emb_dim(args):
        return args.gnn_emb_dim

    @staticmethod
    def need_deg():
        return True

    @staticmethod
    def add_args(parser):
        PNANodeEmbedding.add_args(parser)

    @staticmethod
    def name(args):
        return f"{args.model_type}"

    def __init__(self, num_tasks, node_encoder, edge_encoder_cls, args):
        super().__init__()
        self.num_layer = args.gnn_num_layer
        self.num_tasks = num_tasks
        self.max_seq_len = args.max_seq_len
        self.aggregators = args.aggregators
        self.scalers = args.scalers
        self.residual = args.gnn_residual
        self.drop_ratio = args.gnn_dropout
        self.graph_pooling = args.graph_pooling
        self.node_encoder = node_encoder
        self.pna_module = PNANodeEmbedding(node_encoder, args)
        d = args.gnn_emb_dim
        if self.max_seq_len is None:
            self.mlp = nn.Sequential(nn.Linear(d, 35, bias=True), nn.ReLU(), nn.Linear(35, 17, bias=True), nn.ReLU(),
                                     nn.Linear(17, self.num_tasks, bias=True))
        else:
            self.graph_pred_linear_list = nn.ModuleList(
                [nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, self.num_tasks)) for _ in range(self.max_seq_len)])
        _readout.check_pooling(self.graph_pooling)
        self._w16 = ops.W16Registry()
        self._w16.register(self)
        for m in self.modules():
            if hasattr(m, "register_operands"):
                m.register_operands(self._w16)

    @staticmethod
    def _mlp(seq, h):
        lins = [m for m in seq if isinstance(m, nn.Linear)]
        for i, lin in enumerate(lins):
            last = i == len(lins) - 1
            h = ops.linear(h, lin.weight, lin.bias, relu=not last, out_f32=last)
        return h

    def forward(self, batched_data, perturb=None):
        side = _readout.begin(self, batched_data, self._w16)
        plan = ops.plan_for(batched_data, side_work=side)
        x = self.pna_module.forward_parts(batched_data, perturb, plan=plan)[0]
        h_graph = _readout.pool_nodes(x, plan, self.graph_pooling)
        if self.max_seq_len is None:
            return self._mlp(self.mlp, h_graph)[:, :self.num_tasks]
        return [self._mlp(seq, h_graph)[:, :self.num_tasks] for seq in self.graph_pred_linear_list]
