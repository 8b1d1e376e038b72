"""Model registry with the reference's surface (reference models/__init__.py:9-15).  The two GraphTrans models are
the north-star hot path; the GNN / PNA / Transformer-only baselines (§8f rank 4) run on the same
kernels.  `transformer-gnn` keeps its registry entry, flags and run name only (ablation outside the hot path)."""
from .gnn import GNN
from .gnn_transformer import GNNTransformer
from .pna import PNANet
from .pna_transformer import PNATransformer
from .transformer import Transformer
from .transformer_gnn import TransformerGNN


def get_model_and_parser(args, parser):
    model_cls = MODELS[args.model_type]
    model_cls.add_args(parser)
    return model_cls


MODELS = {"gnn": GNN, "pna": PNANet, "transformer": Transformer, "gnn-transformer": GNNTransformer,
          "pna-transformer": PNATransformer, "transformer-gnn": TransformerGNN}
