import torch
from torch_scatter import scatter
from .num_nodes import maybe_num_nodes
from .loop import add_remaining_self_loops, add_self_loops, remove_self_loops


def softmax(src, index=None, ptr=None, num_nodes=None):
  """PyG 1.7.0 softmax, index branch."""
  assert ptr is None and index is not None
  N = maybe_num_nodes(index, num_nodes)
  out = src - scatter(src, index, dim=0, dim_size=N, reduce='max')[index]
  out = out.exp()
  out_sum = scatter(out, index, dim=0, dim_size=N, reduce='sum')[index]
  return out / (out_sum + 1e-16)


def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):
  N = int(edge_index.max()) + 1 if max_num_nodes is None else max_num_nodes
  if edge_attr is None:
    edge_attr = torch.ones(edge_index.size(1))
  adj = torch.zeros((1, N, N) + tuple(edge_attr.shape[1:]), dtype=edge_attr.dtype)
  adj[0].index_put_((edge_index[0], edge_index[1]), edge_attr, accumulate=True)
  return adj


def to_undirected(edge_index, num_nodes=None):
  both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
  return torch.unique(both, dim=1)


def is_undirected(edge_index, *a, **k):
  fw = set(map(tuple, edge_index.t().tolist()))
  return all((c, r) in fw for r, c in fw)


def dense_to_sparse(adj):
  index = adj.nonzero(as_tuple=False).t().contiguous()
  return index, adj[index[0], index[1]]
