import torch


def degree(index, num_nodes=None, dtype=None):
    """PyG 1.6.3 utils.degree: zeros(N).scatter_add_(0, index, ones)."""
    if num_nodes is None:
        num_nodes = int(index.max()) + 1
    out = torch.zeros((num_nodes,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones((index.size(0),)))
