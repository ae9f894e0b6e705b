import torch


class MessagePassing(torch.nn.Module):
  """The reference subclasses MessagePassing (src/base_classes.py:77,98) but never calls propagate()."""

  def __init__(self, aggr='add', flow='source_to_target', node_dim=-2):
    super().__init__()
