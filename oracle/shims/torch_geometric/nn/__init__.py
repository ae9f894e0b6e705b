from .conv import MessagePassing


class GCNConv:  # imported by name in run_GNN.py / GNN_early.py, never used on the hot path
  pass


class ChebConv:
  pass
