"""Host-side utilities mirroring the reference's src/utils.py for the hot path: graph normalisation (one-time setup,
plain index arithmetic in torch ops on whatever device the edge list lives on), the NFE guard exception, meters and
the dummy dataset shims the reference's tests use."""
import os

import torch

ROOT_DIR = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


class MaxNFEException(Exception):
  """reference src/utils.py:18"""
  pass


def maybe_num_nodes(index, num_nodes=None):
  if num_nodes is not None:
    return int(num_nodes)
  return int(index.max()) + 1 if index.numel() > 0 else 0


def _segment_sum(values, index, size):
  return torch.zeros(size, dtype=values.dtype, device=values.device).index_add_(0, index, values)


def add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1., num_nodes=None):
  """PyG semantics the reference relies on (src/function_transformer_attention.py:18, src/utils.py:63,113):
  existing self loops leave the list, one loop per node is appended at the end; an existing loop keeps its weight,
  the rest get `fill_value`; `None` weights stay `None`."""
  n = maybe_num_nodes(edge_index, num_nodes)
  src, dst = edge_index[0], edge_index[1]
  keep = src != dst
  loops = torch.arange(n, dtype=src.dtype, device=src.device)
  out_index = torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)
  if edge_weight is None:
    return out_index, None
  loop_w = torch.full((n,), fill_value, dtype=edge_weight.dtype, device=edge_weight.device)
  had_loop = ~keep
  if bool(had_loop.any()):
    loop_w[src[had_loop]] = edge_weight[had_loop]
  return out_index, torch.cat([edge_weight[keep], loop_w])


def get_rw_adj(edge_index, edge_weight=None, norm_dim=1, fill_value=0., num_nodes=None, dtype=None):
  """Random-walk normalisation, reference src/utils.py:105-123 (norm_dim=1 divides by the COLUMN degree)."""
  n = maybe_num_nodes(edge_index, num_nodes)
  if edge_weight is None:
    edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
  if not fill_value == 0:
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill_value, n)
  which = edge_index[0] if norm_dim == 0 else edge_index[1]
  inv_deg = _segment_sum(edge_weight, which, n).pow(-1)
  edge_weight = inv_deg[which] * edge_weight if norm_dim == 0 else edge_weight * inv_deg[which]
  return edge_index, edge_weight


def gcn_norm_fill_val(edge_index, edge_weight=None, fill_value=0., num_nodes=None, dtype=None):
  """Symmetric normalisation, reference src/utils.py:55-72 (self loops only when int(fill_value) != 0)."""
  n = maybe_num_nodes(edge_index, num_nodes)
  if edge_weight is None:
    edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
  if not int(fill_value) == 0:
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill_value, n)
  src, dst = edge_index[0], edge_index[1]
  inv_sqrt = _segment_sum(edge_weight, dst, n).pow(-0.5)
  inv_sqrt = inv_sqrt.masked_fill(inv_sqrt == float('inf'), 0)
  return edge_index, inv_sqrt[src] * edge_weight * inv_sqrt[dst]


class Meter(object):
  """Forward / backward NFE counter, reference src/utils.py:212-233."""

  def __init__(self):
    self.reset()

  def reset(self):
    self.val, self.sum, self.cnt = None, 0, 0

  def update(self, val):
    self.val = val
    self.sum += val
    self.cnt += 1

  def get_average(self):
    return 0 if self.cnt == 0 else self.sum / self.cnt

  def get_value(self):
    return self.val


class DummyDataset(object):
  def __init__(self, data, num_classes):
    self.data = data
    self.num_classes = num_classes


class DummyData(object):
  def __init__(self, edge_index=None, edge_Attr=None, num_nodes=None):
    self.edge_index = edge_index
    self.edge_attr = edge_Attr
    self.num_nodes = num_nodes


def print_model_params(model):
  print(model)
  for name, param in model.named_parameters():
    if param.requires_grad:
      print(name)
      print(param.data.shape)
