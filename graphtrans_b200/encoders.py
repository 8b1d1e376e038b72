"""Node / edge encoders with the reference's parameter names, executed by gt_embed_sum_*.

* ASTNodeEncoder  - reference dataset/utils.py:8-30 (type + attribute + clamped depth embeddings)
* AtomEncoder / BondEncoder - ogb 1.2.6 `ogb.graphproppred.mol_encoder`:
  sums of per-column nn.Embedding tables, xavier-uniform init, same state_dict keys.
"""
import torch

from . import ops
from .synth import ATOM_DIMS, BOND_DIMS


class ASTNodeEncoder(torch.nn.Module):
    def __init__(self, emb_dim, num_nodetypes, num_nodeattributes, max_depth):
        super().__init__()
        self.max_depth = max_depth
        self.type_encoder = torch.nn.Embedding(num_nodetypes, emb_dim)
        self.attribute_encoder = torch.nn.Embedding(num_nodeattributes, emb_dim)
        self.depth_encoder = torch.nn.Embedding(self.max_depth + 1, emb_dim)

    def forward(self, x, depth):
        # the clamp of dataset/utils.py:29 happens inside the kernel (the batch is not mutated)
        tabs = [self.type_encoder.weight, self.attribute_encoder.weight, self.depth_encoder.weight]
        return ops.embed_sum([x[:, 0], x[:, 1], depth.view(-1)], tabs,
                             clamps=[tabs[0].shape[0] - 1, tabs[1].shape[0] - 1, self.max_depth])


class AtomEncoder(torch.nn.Module):
    def __init__(self, emb_dim):
        super().__init__()
        self.atom_embedding_list = torch.nn.ModuleList()
        for dim in ATOM_DIMS:
            emb = torch.nn.Embedding(dim, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)

    def forward(self, x):
        return ops.embed_sum([x[:, c] for c in range(x.shape[1])],
                             [self.atom_embedding_list[c].weight for c in range(x.shape[1])])


class BondEncoder(torch.nn.Module):
    """Never materialises [E, d]: the convs read `bond_embedding_list` as a combined table."""

    def __init__(self, emb_dim):
        super().__init__()
        self.bond_embedding_list = torch.nn.ModuleList()
        for dim in BOND_DIMS:
            emb = torch.nn.Embedding(dim, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)

    def forward(self, edge_attr):
        return ops.embed_sum([edge_attr[:, c] for c in range(edge_attr.shape[1])],
                             [self.bond_embedding_list[c].weight for c in range(edge_attr.shape[1])])
