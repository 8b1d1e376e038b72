"""Registry shell of the reference's Transformer-then-GNN ablation (reference models/transformer_gnn.py:22-192).
Flags and run name are kept so that `MODELS` / `get_model_and_parser` / `name(args)` behave like the reference's;
the model itself is outside the GraphTrans hot path and is not built."""
from ..modules.masked_transformer_encoder import MaskedOnlyTransformerEncoder
from ..modules.transformer_encoder import TransformerNodeEncoder
from .base_model import BaseModel
from .gnn_transformer import GNNTransformer


class TransformerGNN(BaseModel):
    @staticmethod
    def get_emb_dim(args):
        return args.gnn_emb_dim

    @staticmethod
    def add_args(parser):
        TransformerNodeEncoder.add_args(parser)
        MaskedOnlyTransformerEncoder.add_args(parser)
        group = parser.add_argument_group("GNNTransformer - Training Config")
        group.add_argument("--pretrained_gnn", type=str, default=None, help="pretrained gnn_node node embedding path")
        group.add_argument("--freeze_gnn", type=int, default=None, help="Freeze gnn_node weight from epoch `freeze_gnn`")
        group.add_argument("--graph_input_dim", type=int, default=None)

    name = staticmethod(GNNTransformer.name)      # same run-name format (models/transformer_gnn.py:38-54)

    def __init__(self, num_tasks, node_encoder, edge_encoder_cls, args):
        super().__init__()
        raise NotImplementedError("transformer-gnn (Transformer -> GNN ablation) is outside the GraphTrans hot path, SURVEY §8f")
