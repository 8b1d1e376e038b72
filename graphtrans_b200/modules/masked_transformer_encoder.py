"""API-compatibility shell of the reference's MaskedOnlyTransformerEncoder (reference
modules/masked_transformer_encoder.py:104-130).  It is dead in every shipped config
(num_encoder_layers_masked defaults to 0, contributes no parameters) and its reference
semantics are broken (inverted padding mask), so only the flags and the
zero-layer behaviour are kept."""
import torch.nn as nn


class MaskedOnlyTransformerEncoder(nn.Module):
    @staticmethod
    def add_args(parser):
        group = parser.add_argument_group("Masked Transformer Encoder -- architecture config")
        group.add_argument("--num_encoder_layers_masked", type=int, default=0)
        group.add_argument("--transformer_prenorm", action="store_true", default=False)

    def __init__(self, args):
        super().__init__()
        self.max_input_len = args.max_input_len
        if args.num_encoder_layers_masked > 0:
            raise NotImplementedError("num_encoder_layers_masked > 0 is out of scope (SURVEY.md §2 row 7)")

    def forward(self, x, attn_mask=None, valid_input_mask=None):
        return x
