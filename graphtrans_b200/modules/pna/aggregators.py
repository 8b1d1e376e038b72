"""PNA aggregators with the reference's names and call signature (reference modules/pna/aggregators.py:11-44):
`AGGREGATORS[name](src, index, dim_size)` reduces the rows of `src` that share `index` (torch_scatter semantics: empty
segments give 0).  The MODEL never calls these: mean / max / min / std of all messages are taken by the single-pass
kernel gt_pna_reduce_* (csrc/pna.cu).  They exist for callers of the reference's module surface and are plain device
tensor ops (index_add_ / scatter_reduce_), differentiable through autograd."""
import torch


def _rows(index, src):
    return index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)


def aggregate_sum(src, index, dim_size):
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def aggregate_mean(src, index, dim_size):
    cnt = torch.bincount(index, minlength=dim_size).clamp(min=1).to(src.dtype)
    return aggregate_sum(src, index, dim_size) / cnt.view(-1, *([1] * (src.dim() - 1)))


def _extreme(src, index, dim_size, mode):
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.scatter_reduce(0, _rows(index, src), src, mode, include_self=False)


def aggregate_min(src, index, dim_size):
    return _extreme(src, index, dim_size, "amin")


def aggregate_max(src, index, dim_size):
    return _extreme(src, index, dim_size, "amax")


def aggregate_var(src, index, dim_size):
    mean = aggregate_mean(src, index, dim_size)
    mean_squares = aggregate_mean(src * src, index, dim_size)
    return mean_squares - mean * mean


def aggregate_std(src, index, dim_size):
    return torch.sqrt(torch.relu(aggregate_var(src, index, dim_size)) + 1e-5)


AGGREGATORS = {"sum": aggregate_sum, "mean": aggregate_mean, "min": aggregate_min, "max": aggregate_max,
               "var": aggregate_var, "std": aggregate_std}
BUILT = ("mean", "max", "min", "std")      # the set the fused kernel computes (reference default, pna_module.py:20)
