"""TEST INFRASTRUCTURE ONLY. Pure-torch stand-in for torch-scatter==2.0.6 `scatter`
(pinned by the reference's requirement.yml:98) so that the UNMODIFIED reference under
/root/reference imports and runs on CPU. Semantics follow the published torch-scatter 2.0.6
behaviour: sum; mean = sum / clamp(count, 1); min/max give 0 for
empty segments and route gradient to one arg element per (segment, channel).
Never imported by the product package."""
import torch


def _expand_index(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src), dim


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None
    idx, dim = _expand_index(index, src, dim)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.shape)
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add(dim, idx, src)
    if reduce == "mean":
        total = torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add(dim, idx, src)
        cnt_shape = [1] * src.dim()
        cnt_shape[dim] = dim_size
        count = torch.zeros(dim_size, dtype=src.dtype, device=src.device).scatter_add(
            0, index, torch.ones_like(index, dtype=src.dtype))
        return total / count.clamp(min=1).view(cnt_shape)
    if reduce in ("min", "max"):
        # include_self=False leaves untouched (empty) segments at the initial 0, like torch-scatter
        init = torch.zeros(size, dtype=src.dtype, device=src.device)
        return init.scatter_reduce(dim, idx, src, "amin" if reduce == "min" else "amax", include_self=False)
    raise ValueError(reduce)
