"""GNN baseline with the reference's surface (reference models/gnn.py:16-115): GNN stack -> global pooling ->
head(s), on the same aggregation / GEMM / BatchNorm kernels as GraphTrans."""
from .. import ops
from ..modules.gnn_module import GNNNodeEmbedding
from . import _readout
from .base_model import BaseModel


class GNN(BaseModel):
    @staticmethod
    def get_emb_dim(args):
        return args.gnn_emb_dim

    @staticmethod
    def add_args(parser):
        return

    @staticmethod
    def name(args):
        name = f"{args.model_type}+{args.gnn_type}"
        name += "-virtual" if args.gnn_virtual_node else ""
        return name

    def __init__(self, num_tasks, node_encoder, edge_encoder_cls, args):
        super().__init__()
        self.num_layer = args.gnn_num_layer
        self.drop_ratio = args.gnn_dropout
        self.JK = args.gnn_JK
        self.emb_dim = args.gnn_emb_dim
        self.num_tasks = num_tasks
        self.max_seq_len = args.max_seq_len
        self.graph_pooling = args.graph_pooling
        if self.num_layer < 2:
            raise ValueError("Number of GNN layers must be greater than 1.")
        if self.JK == "cat":
            raise ValueError("JK=cat doubles the node width; the reference's GNN head takes emb_dim (models/gnn.py:88)")
        self.gnn_node = GNNNodeEmbedding(
            args.gnn_virtual_node, self.num_layer, self.emb_dim, node_encoder, edge_encoder_cls, JK=self.JK,
            drop_ratio=self.drop_ratio, residual=args.gnn_residual, gnn_type=args.gnn_type)
        _readout.check_pooling(self.graph_pooling)
        _readout.make_heads(self, self.emb_dim, self.num_tasks, self.max_seq_len)
        self._w16 = ops.W16Registry()
        self._w16.register(self)

    def forward(self, batched_data, perturb=None):
        side = _readout.begin(self, batched_data, self._w16)
        plan = ops.plan_for(batched_data, side_work=side)
        h_node = self.gnn_node.forward_parts(batched_data, perturb, plan=plan)[0]
        h_graph = _readout.pool_nodes(h_node, plan, self.graph_pooling)
        return _readout.apply_heads(self, h_graph, self.num_tasks, self.max_seq_len)
