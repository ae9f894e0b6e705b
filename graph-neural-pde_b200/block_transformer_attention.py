"""AttODEblock -- reference src/block_transformer_attention.py:7-76: the paper's GRAND-l.

The block owns its own SpGraphTransAttentionLayer.  Every forward it evaluates the attention ONCE at x(0)
(`get_attention_weights`), freezes the [E, H] result into `odefunc.attention_weights`, and the Laplacian function
integrates with the head mean of it."""
import torch

from .base_classes import ODEblock
from .function_transformer_attention import SpGraphTransAttentionLayer
from .utils import get_rw_adj


class AttODEblock(ODEblock):
  def __init__(self, odefunc, regularization_fns, opt, data, device, t=torch.tensor([0, 1]), gamma=0.5):
    super(AttODEblock, self).__init__(odefunc, regularization_fns, opt, data, device, t)
    self.rebuild_function(odefunc, data)
    self.install_graph(*get_rw_adj(data.edge_index, edge_weight=data.edge_attr, norm_dim=1,
                                   fill_value=opt['self_loop_weight'], num_nodes=data.num_nodes, dtype=data.x.dtype))
    self.use_default_integrators()
    width = opt['hidden_dim']
    self.multihead_att_layer = SpGraphTransAttentionLayer(width, width, opt, device,
                                                          edge_weights=self.odefunc.edge_weight).to(device)

  def get_attention_weights(self, x):
    attention, _ = self.multihead_att_layer(x, self.odefunc.edge_index, engine=self.odefunc.engine(x))
    return attention

  def forward(self, x):
    frozen = self.get_attention_weights(x)
    self.odefunc.attention_weights = frozen
    self.reg_odefunc.odefunc.attention_weights = frozen
    o = self.opt
    return self.integrate(x, solver_options={'step_size': o['step_size']},
                          adjoint_options=lambda: {'step_size': o['adjoint_step_size']})
