import torch
from .num_nodes import maybe_num_nodes


def remove_self_loops(edge_index, edge_attr=None):
  mask = edge_index[0] != edge_index[1]
  edge_index = edge_index[:, mask]
  return (edge_index, None) if edge_attr is None else (edge_index, edge_attr[mask])


def add_self_loops(edge_index, edge_weight=None, fill_value=1., num_nodes=None):
  N = maybe_num_nodes(edge_index, num_nodes)
  loop_index = torch.arange(0, N, dtype=torch.long, device=edge_index.device).unsqueeze(0).repeat(2, 1)
  if edge_weight is not None:
    loop_weight = edge_weight.new_full((N,), fill_value)
    edge_weight = torch.cat([edge_weight, loop_weight], dim=0)
  return torch.cat([edge_index, loop_index], dim=1), edge_weight


def add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1., num_nodes=None):
  N = maybe_num_nodes(edge_index, num_nodes)
  row, col = edge_index[0], edge_index[1]
  mask = row != col
  loop_index = torch.arange(0, N, dtype=row.dtype, device=row.device)
  loop_index = loop_index.unsqueeze(0).repeat(2, 1)
  edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
  if edge_weight is not None:
    inv_mask = ~mask
    loop_weight = torch.full((N,), fill_value, dtype=edge_weight.dtype, device=edge_weight.device)
    remaining_edge_weight = edge_weight[inv_mask]
    if remaining_edge_weight.numel() > 0:
      loop_weight[row[inv_mask]] = remaining_edge_weight
    edge_weight = torch.cat([edge_weight[mask], loop_weight], dim=0)
  return edge_index, edge_weight
