"""AttODEblock (reference src/block_transformer_attention.py:7-76): the paper's GRAND-l -- attention is computed ONCE
per forward from x(0) by the block's own SpGraphTransAttentionLayer, frozen into `odefunc.attention_weights`, and the
Laplacian function integrates with it."""
import torch

from .base_classes import ODEblock
from .function_transformer_attention import SpGraphTransAttentionLayer
from .odeint import odeint, odeint_adjoint
from .utils import get_rw_adj


class AttODEblock(ODEblock):
  def __init__(self, odefunc, regularization_fns, opt, data, device, t=torch.tensor([0, 1]), gamma=0.5):
    super(AttODEblock, self).__init__(odefunc, regularization_fns, opt, data, device, t)
    self.odefunc = odefunc(self.aug_dim * opt['hidden_dim'], self.aug_dim * opt['hidden_dim'], opt, data, device)
    edge_index, edge_weight = get_rw_adj(data.edge_index, edge_weight=data.edge_attr, norm_dim=1,
                                         fill_value=opt['self_loop_weight'], num_nodes=data.num_nodes,
                                         dtype=data.x.dtype)
    self.odefunc.edge_index = edge_index.to(device)
    self.odefunc.edge_weight = edge_weight.to(device)
    self.reg_odefunc.odefunc.edge_index = self.odefunc.edge_index
    self.reg_odefunc.odefunc.edge_weight = self.odefunc.edge_weight
    self.train_integrator = odeint_adjoint if opt['adjoint'] else odeint
    self.test_integrator = odeint_adjoint if opt['adjoint'] else odeint
    self.set_tol()
    self.multihead_att_layer = SpGraphTransAttentionLayer(opt['hidden_dim'], opt['hidden_dim'], opt, device,
                                                          edge_weights=self.odefunc.edge_weight).to(device)

  def get_attention_weights(self, x):
    attention, values = self.multihead_att_layer(x, self.odefunc.edge_index, engine=self.odefunc.engine(x))
    return attention

  def forward(self, x):
    self.odefunc.attention_weights = self.get_attention_weights(x)
    self.reg_odefunc.odefunc.attention_weights = self.odefunc.attention_weights
    return self._integrate(x, {'step_size': self.opt['step_size']},
                           lambda: {'step_size': self.opt['adjoint_step_size']})
