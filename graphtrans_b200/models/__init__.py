"""Model registry with the reference's surface (reference models/__init__.py:9-15).  Only the
north-star hot path is built: the GNN / PNA / Transformer-only / TransformerGNN
baselines are out of scope."""
from .gnn_transformer import GNNTransformer
from .pna_transformer import PNATransformer


def get_model_and_parser(args, parser):
    model_cls = MODELS[args.model_type]
    model_cls.add_args(parser)
    return model_cls


MODELS = {"gnn-transformer": GNNTransformer, "pna-transformer": PNATransformer}
