"""Stand-in for torch-scatter 2.0.x: only the entry points the reference's hot path calls."""
import torch


def _expand(index, src, dim):
  shape = [1] * src.dim()
  shape[dim] = -1
  return index.view(shape).expand_as(src)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
  dim = dim if dim >= 0 else src.dim() + dim
  if index.dim() == 1 and src.dim() > 1:
    index = _expand(index, src, dim)
  if dim_size is None:
    dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
  size = list(src.shape)
  size[dim] = dim_size
  if reduce in ('sum', 'add'):
    return torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, index, src)
  if reduce == 'mean':
    tot = torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, index, src)
    cnt = torch.zeros(size, dtype=src.dtype, device=src.device).scatter_add_(dim, index, torch.ones_like(src))
    return tot / cnt.clamp(min=1)
  if reduce == 'max':
    res = torch.full(size, -float('inf'), dtype=src.dtype, device=src.device)
    res = res.scatter_reduce_(dim, index, src, reduce='amax', include_self=True)
    return torch.where(torch.isinf(res) & (res < 0), torch.zeros_like(res), res)
  if reduce == 'min':
    res = torch.full(size, float('inf'), dtype=src.dtype, device=src.device)
    res = res.scatter_reduce_(dim, index, src, reduce='amin', include_self=True)
    return torch.where(torch.isinf(res) & (res > 0), torch.zeros_like(res), res)
  raise ValueError(reduce)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
  return scatter(src, index, dim, out, dim_size, 'sum')


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
  return scatter(src, index, dim, out, dim_size, 'max'), None


def segment_csr(src, indptr, out=None, reduce='sum'):
  n = indptr.numel() - 1
  index = torch.repeat_interleave(torch.arange(n, device=src.device), indptr[1:] - indptr[:-1])
  return scatter(src, index, 0, None, n, reduce)


def gather_csr(src, indptr, out=None):
  n = indptr.numel() - 1
  index = torch.repeat_interleave(torch.arange(n, device=src.device), indptr[1:] - indptr[:-1])
  return src[index]
