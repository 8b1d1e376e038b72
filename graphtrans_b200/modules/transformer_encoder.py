"""TransformerNodeEncoder with the reference's flags, attributes and state_dict keys (reference
modules/transformer_encoder.py:9-61).  The nn.TransformerEncoder it owns is used as a PARAMETER
CONTAINER only (identical keys / init as the reference); the math runs on packed tokens through
gt_gemm / gt_mha_* / gt_layernorm_* (post-norm layers, ReLU FFN, final LayerNorm)."""
import os
import types

import torch
import torch.nn as nn

from .. import ops

# last encoder layer: compute only the pooled query row of every graph (exactly what the model reads, reference
# models/gnn_transformer.py:114-115); GT_POOLED_LAST=0 runs the full layer (A/B switch for the parity tests)
POOLED_LAST = int(os.environ.get("GT_POOLED_LAST", "1"))


class TransformerNodeEncoder(nn.Module):
    @staticmethod
    def add_args(parser):
        group = parser.add_argument_group("transformer")
        group.add_argument("--d_model", type=int, default=128, help="transformer d_model.")
        group.add_argument("--nhead", type=int, default=4, help="transformer heads")
        group.add_argument("--dim_feedforward", type=int, default=512, help="transformer feedforward dim")
        group.add_argument("--transformer_dropout", type=float, default=0.3)
        group.add_argument("--transformer_activation", type=str, default="relu")
        group.add_argument("--num_encoder_layers", type=int, default=4)
        group.add_argument("--max_input_len", default=1000, help="The max input length of transformer input")
        group.add_argument("--transformer_norm_input", action="store_true", default=False)

    def __init__(self, args):
        super().__init__()
        self.d_model = args.d_model
        self.num_layer = args.num_encoder_layers
        self.nhead = args.nhead
        self.dropout = args.transformer_dropout
        if args.transformer_activation != "relu":
            raise NotImplementedError("only the reference default transformer_activation='relu' is built")
        if args.d_model % 8 or (args.d_model // args.nhead) % 4 or args.d_model // args.nhead > 64:
            raise ValueError("d_model must be a multiple of 8 with head dim a multiple of 4 and <= 64")
        encoder_layer = nn.TransformerEncoderLayer(args.d_model, args.nhead, args.dim_feedforward,
                                                   args.transformer_dropout, args.transformer_activation)
        encoder_norm = nn.LayerNorm(args.d_model)
        self.transformer = nn.TransformerEncoder(encoder_layer, args.num_encoder_layers, encoder_norm,
                                                 enable_nested_tensor=False)
        self.max_input_len = int(args.max_input_len)   # no type= on the flag (transformer_encoder.py:19)
        self.norm_input = None
        if args.transformer_norm_input:
            self.norm_input = nn.LayerNorm(args.d_model)
        self.cls_embedding = None
        if args.graph_pooling == "cls":
            self.cls_embedding = nn.Parameter(torch.randn([1, 1, args.d_model], requires_grad=True))
        # the pooled last layer applies in_proj in two row blocks (q | k,v): two gradient contributions per step
        at = self.transformer.layers[-1].self_attn
        at.in_proj_weight._gt_uses = at.in_proj_bias._gt_uses = 2

    # ------------------------------------------------------------------ packed path (model)
    def _ffn_block(self, layer, a, x, drop):
        """out-projection .. norm2 of one post-norm encoder layer on the rows of `a` (residual stream `x`)"""
        at = layer.self_attn
        a = ops.linear(a, at.out_proj.weight, at.out_proj.bias)
        x1 = ops.layer_norm(a, layer.norm1, resid=x, drop_p=drop)                           # norm1(x + drop(a))
        f, x1 = ops.linear(x1, layer.linear1.weight, layer.linear1.bias, relu=True, drop_p=drop, passthrough=True)
        f = ops.linear(f, layer.linear2.weight, layer.linear2.bias)
        return ops.layer_norm(f, layer.norm2, resid=x1, drop_p=drop)                         # norm2(x + drop(f))

    def encode_pooled_last(self, x, plan):
        """LAST layer for the pooled rows only: k | v of every token, q / out-proj / norms / FFN of the B pooled rows
        (plan.cls_rows: <CLS>, or the last node with pooling == 'last').  x: [n_rows, d] -> [B, d]."""
        drop = self.dropout if self.training else 0.0
        layer = self.transformer.layers[-1]
        at = layer.self_attn
        d = self.d_model
        kv, x = ops.linear(x, at.in_proj_weight, at.in_proj_bias, w_row_off=d, n_out=2 * d, passthrough=True)   # [n_rows, 2d]
        xq = ops.gather_rows(x, plan.cls_rows, n_rows=plan.B)                                 # [B, d] residual stream
        q = ops.linear(xq, at.in_proj_weight, at.in_proj_bias, w_row_off=0, n_out=d)         # [B, d]
        a = ops.mha_pooled_query(q, kv, plan, self.nhead, drop_p=drop)
        return self._ffn_block(layer, a, xq, drop)

    def encode_layers(self, x, plan, key_start=None, skip_last=False):
        """x: [n_rows, d] packed tokens -> last layer output (before the final norm)."""
        drop = self.dropout if self.training else 0.0
        layers = list(self.transformer.layers)
        for layer in (layers[:-1] if skip_last else layers):
            at = layer.self_attn
            # passthrough: the residual stream is handed on by the Linear that also reads it, so the residual path's
            # gradient is added in that Linear's dX epilogue (no autograd accumulation kernel per sublayer)
            qkv, x = ops.linear(x, at.in_proj_weight, at.in_proj_bias, passthrough=True)
            a = ops.mha_packed(qkv, plan, self.nhead, key_start, drop_p=drop)
            a = ops.linear(a, at.out_proj.weight, at.out_proj.bias)
            x1 = ops.layer_norm(a, layer.norm1, resid=x, drop_p=drop)                       # norm1(x + drop(a))
            f, x1 = ops.linear(x1, layer.linear1.weight, layer.linear1.bias, relu=True, drop_p=drop, passthrough=True)
            f = ops.linear(f, layer.linear2.weight, layer.linear2.bias)
            x = ops.layer_norm(f, layer.norm2, resid=x1, drop_p=drop)                       # norm2(x + drop(f))
        return x

    def _tokens(self, h_node, plan):
        cls = self.cls_embedding
        if plan.cls != (cls is not None):
            raise RuntimeError("GraphPlan was built with a different <CLS> setting than the encoder")
        if self.norm_input is not None:
            return ops.layer_norm(h_node, self.norm_input, in_rows=plan.tok2node, cls=cls, n_rows=plan.n_rows)
        return ops.gather_rows(h_node, plan.tok2node, cls=cls, n_rows=plan.n_rows)

    def forward_tokens(self, h_node, plan):
        """h_node [N, d] -> encoder output of EVERY packed token row [n_rows, d] (final norm applied): the callers that
        pool over nodes afterwards (reference models/transformer.py:105-106)"""
        x = self.encode_layers(self._tokens(h_node, plan), plan)
        return ops.layer_norm(x, self.transformer.norm)

    def forward_packed(self, h_node, plan):
        """h_node: [N, d] node states (gnn2transformer output) -> [B, d] encoder output at the pooled
        position (<CLS> row, or the last node when pooling == 'last')."""
        x = self._tokens(h_node, plan)
        if POOLED_LAST:
            x = self.encode_layers(x, plan, skip_last=True)
            return ops.layer_norm(self.encode_pooled_last(x, plan), self.transformer.norm)
        x = self.encode_layers(x, plan)
        return ops.layer_norm(x, self.transformer.norm, in_rows=plan.cls_rows, n_rows=plan.B)

    # ------------------------------------------------------------------ dense public API
    def forward(self, padded_h_node, src_padding_mask):
        """padded_h_node: [S, B, d]; src_padding_mask: [B, S] bool, True = PAD, left-padded as
        pad_batch produces it.  Returns ([T, B, d], [B, T]) like the reference."""
        S, B, d = padded_h_node.shape
        if self.cls_embedding is not None:
            expand_cls = self.cls_embedding.expand(1, B, -1).to(padded_h_node.dtype)
            padded_h_node = torch.cat([padded_h_node, expand_cls], dim=0)
            zeros = src_padding_mask.new_zeros(B, 1)
            src_padding_mask = torch.cat([src_padding_mask, zeros], dim=1)
        T = padded_h_node.shape[0]
        x = padded_h_node.transpose(0, 1).reshape(B * T, d)           # batch-first rows: graph g owns [gT, gT+T)
        x = ops.pad_cols(x, d, ops.act_dtype())
        if self.norm_input is not None:
            x = ops.layer_norm(x, self.norm_input)
        dev = x.device
        dense = types.SimpleNamespace(
            B=B, tok_off=(torch.arange(B + 1, device=dev, dtype=torch.int32) * T).contiguous(),
            tok_graph=torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(T).contiguous())
        key_start = (dense.tok_off[:-1] + src_padding_mask.sum(1).to(torch.int32)).contiguous()
        x = self.encode_layers(x, dense, key_start)
        x = ops.layer_norm(x, self.transformer.norm)
        return x.float().view(B, T, d).transpose(0, 1), src_padding_mask
