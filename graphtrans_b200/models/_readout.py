"""Shared tail of the reference's baseline models (models/gnn.py, models/pna.py, models/transformer.py): global pooling
over the nodes of each graph followed by the prediction head(s).  Pooling runs on the existing segment kernels
(gt_segment_sum_sorted / gt_segment_pool_*); `attention` and `set2set` read-outs (PyG GlobalAttention / Set2Set) are
outside the hot path and raise."""
import torch

from .. import ops

POOLINGS = ("sum", "mean", "max")


def check_pooling(kind, extra=()):
    if kind in ("attention", "set2set"):
        raise NotImplementedError(f"graph_pooling={kind} (PyG GlobalAttention / Set2Set) is out of scope, SURVEY §8")
    if kind not in POOLINGS + tuple(extra):
        raise ValueError("Invalid graph pooling type.")


def pool_nodes(h, plan, kind):
    """physical [N, ld] node states -> [B, ld] in the activation dtype"""
    return ops.cast_to(ops.segment_pool(h, plan, kind), ops.act_dtype())


def make_heads(module, in_dim, num_tasks, max_seq_len):
    if max_seq_len is None:
        module.graph_pred_linear = torch.nn.Linear(in_dim, num_tasks)
    else:
        module.graph_pred_linear_list = torch.nn.ModuleList(
            [torch.nn.Linear(in_dim, num_tasks) for _ in range(max_seq_len)])


def apply_heads(module, h_graph, num_tasks, max_seq_len):
    if max_seq_len is None:
        lin = module.graph_pred_linear
        return ops.linear(h_graph, lin.weight, lin.bias, out_f32=True)[:, :num_tasks]
    return [ops.linear(h_graph, lin.weight, lin.bias, out_f32=True)[:, :num_tasks] for lin in module.graph_pred_linear_list]


def begin(model, batched_data, registry):
    """per-forward bookkeeping shared by the baseline models: dropout step / bf16 weight copies / batch plan"""
    ops._lib.require_cuda(batched_data.batch, batched_data.edge_index)
    dev = batched_data.batch.device
    side = None
    if model.training:
        ops.begin_step(dev, cast_now=False, registry=registry)
    else:
        ops.w16 = registry
    if ops.precision() == "bf16":
        side = lambda: registry.refresh(dev)  # noqa: E731
    return side
