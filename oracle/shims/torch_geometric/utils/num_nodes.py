def maybe_num_nodes(index, num_nodes=None):
  if num_nodes is not None:
    return num_nodes
  return int(index.max()) + 1 if index.numel() > 0 else 0
