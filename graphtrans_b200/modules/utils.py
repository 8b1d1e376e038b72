"""pad_batch / unpad_batch with the reference's signatures (reference modules/utils.py:5-53).

The model itself never pads: GNNTransformer runs the transformer over PACKED tokens
(ops.GraphPlan).  These functions serve callers of the public API; the padded tensor and the
mask are produced by one kernel each way (gt_pad_batch_fwd/bwd), bit-exactly equal to the
reference's Python loop (left padding, truncation keeps the LAST max_input_len nodes)."""
import torch

from .. import ops


class _LazyMasks:
    """list-like of the B per-graph boolean node masks (reference utils.py:8-10), built on demand"""

    def __init__(self, batch, B):
        self.batch, self.B = batch, B

    def __len__(self):
        return self.B

    def __getitem__(self, i):
        if i < 0:
            i += self.B
        if not 0 <= i < self.B:
            raise IndexError(i)
        return self.batch.eq(i)

    def __iter__(self):
        return (self.batch.eq(i) for i in range(self.B))


def pad_batch(h_node, batch, max_input_len, get_mask=False, plan=None):
    d = h_node.shape[-1]
    ld = ops.ldp(d)
    if plan is None or plan.L != int(max_input_len):
        plan = ops.GraphPlan(torch.zeros(2, 0, dtype=torch.long, device=batch.device), batch, None, int(max_input_len))
    S = plan.S
    x = ops.pad_cols(h_node, ld)
    padded, mask = ops.pad_batch_dense(x, plan, S)
    padded_h_node = padded[:, :, :d] if ld != d else padded
    src_padding_mask = mask.bool()
    if get_mask:
        n = (plan.node_off[1:] - plan.node_off[:-1]).to(torch.long)
        num_nodes = list(n.unbind())
        return padded_h_node, src_padding_mask, num_nodes, _LazyMasks(batch, plan.B), S
    return padded_h_node, src_padding_mask


def unpad_batch(padded_h_node, prev_h_node, num_nodes, origin_mask, max_num_nodes):
    """inverse gather of pad_batch: rows of truncated-away nodes keep `prev_h_node`
    (reference modules/utils.py:32-53)."""
    S, B, d = padded_h_node.shape
    n = torch.stack([torch.as_tensor(v, device=padded_h_node.device).reshape(()) for v in num_nodes]).to(torch.long)
    off = torch.cumsum(n, 0) - n
    N = prev_h_node.shape[0]
    graph = torch.repeat_interleave(torch.arange(B, device=n.device), n, output_size=N)
    pos = torch.arange(N, device=n.device) - off[graph] - n[graph] + S      # padded position of node r
    valid = pos >= 0
    src = padded_h_node[pos.clamp(min=0), graph]
    return torch.where(valid.unsqueeze(-1), src.to(prev_h_node.dtype), prev_h_node)
