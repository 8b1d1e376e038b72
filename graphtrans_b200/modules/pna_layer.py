"""`modules.pna_layer` surface of the reference (reference modules/pna_layer.py:20-283).

`PNAConv` has the reference's constructor signature and parameter names (pre_nns / post_nns / lin) and runs on the
fused kernels (per-node tower projections on the tensor cores + gt_pna_reduce_*: one pass for mean / max / min / std and
the three degree scalers).  Built for the configuration the reference actually instantiates
(modules/pna/pna_module.py:43-51: edge_dim=None, divide_input=True, pre_layers = post_layers = 1, in == out); anything
else raises.  `PNAConvSimple` is dead code in the reference (imported nowhere) and is a raising shell.
"""
import torch

from .. import ops
from .pna.pna_module import PNAConv as _FusedPNAConv


class PNAConv(_FusedPNAConv):
    def __init__(self, in_channels, out_channels, aggregators, scalers, deg, edge_dim=None, towers=1, pre_layers=1,
                 post_layers=1, divide_input=False, **kwargs):
        if edge_dim is not None or pre_layers != 1 or post_layers != 1:
            raise NotImplementedError("PNAConv: only edge_dim=None, pre_layers = post_layers = 1 are built "
                                      "(the configuration of reference modules/pna/pna_module.py:43-51)")
        super().__init__(in_channels, out_channels, aggregators, scalers, deg, towers=towers, divide_input=divide_input)

    def forward(self, x, edge_index, edge_attr=None, plan=None):
        """x: logical [N, in_channels] (any float dtype) or the physical activation matrix; edge_index int64 [2, E]"""
        d, ld = self.in_channels, ops.ldp(self.in_channels)
        logical = x.shape[1] != ld or x.dtype != ops.act_dtype()
        if logical:
            x = ops.pad_cols(x, ld, ops.act_dtype())
        if plan is None:
            plan = ops.GraphPlan(edge_index, torch.zeros(x.shape[0], dtype=torch.long, device=x.device), 1)
        out = super().forward(x, plan=plan)
        return out[:, :d].float() if logical else out


class PNAConvSimple(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("PNAConvSimple is not used by any model of the reference (SURVEY §2 row 9)")
