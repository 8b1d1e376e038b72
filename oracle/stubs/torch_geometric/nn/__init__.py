import torch
from torch_scatter import scatter

from .conv import MessagePassing  # noqa: F401


def _num_graphs(batch, size):
    return int(batch.max()) + 1 if size is None else size


def global_add_pool(x, batch, size=None):
    return scatter(x, batch, dim=0, dim_size=_num_graphs(batch, size), reduce="sum")


def global_mean_pool(x, batch, size=None):
    return scatter(x, batch, dim=0, dim_size=_num_graphs(batch, size), reduce="mean")


def global_max_pool(x, batch, size=None):
    return scatter(x, batch, dim=0, dim_size=_num_graphs(batch, size), reduce="max")


class BatchNorm(torch.nn.Module):
    """PyG BatchNorm: wraps BatchNorm1d as `.module` (state keys `module.*`)."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def reset_parameters(self):
        self.module.reset_parameters()

    def forward(self, x):
        return self.module(x)


class GlobalAttention(torch.nn.Module):  # placeholder (models/gnn.py attention pooling only)
    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn, self.nn = gate_nn, nn


class Set2Set(torch.nn.Module):  # placeholder
    def __init__(self, in_channels, processing_steps, num_layers=1):
        super().__init__()


def __getattr__(name):
    # torch_geometric.nn.PNAConv == the reference's own vendored statement of it
    # (/root/reference/modules/pna_layer.py:20-171, edge_dim=None path), whose four missing
    # torch.nn names are injected here; resolved lazily to avoid an import cycle.
    if name == "PNAConv":
        import importlib
        pl = importlib.import_module("modules.pna_layer")
        for n in ("ModuleList", "Linear", "ReLU", "Sequential"):
            setattr(pl, n, getattr(torch.nn, n))
        return pl.PNAConv
    raise AttributeError(name)
