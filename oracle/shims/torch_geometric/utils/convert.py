import scipy.sparse


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
  import torch
  row, col = edge_index.cpu()
  if edge_attr is None:
    edge_attr = torch.ones(row.size(0))
  N = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
  return scipy.sparse.coo_matrix((edge_attr.view(-1).cpu().numpy(), (row.numpy(), col.numpy())), (N, N))
