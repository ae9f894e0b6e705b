"""Stand-in for torch-sparse 0.6.x: spmm as index_select * value -> scatter_add (the published algorithm)."""
import torch
from torch_scatter import scatter_add


def spmm(index, value, m, n, matrix):
  assert n == matrix.size(-2)
  row, col = index[0], index[1]
  matrix = matrix if matrix.dim() > 1 else matrix.unsqueeze(-1)
  out = matrix.index_select(-2, col)
  out = out * value.unsqueeze(-1)
  return scatter_add(out, row, dim=-2, dim_size=m)


def coalesce(index, value, m, n, op='add'):
  raise NotImplementedError('coalesce is outside the hot path')


def spspmm(*a, **k):
  raise NotImplementedError('spspmm is outside the hot path')
