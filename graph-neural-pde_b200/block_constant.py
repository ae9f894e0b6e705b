"""ConstantODEblock (reference src/block_constant.py:6-74): the graph is normalised once at construction and stays
fixed; GRAND-nl when the function is ODEFuncTransformerAtt, GRAND-l with random-walk weights when it is
LaplacianODEFunc.  `forward(x)` hands the whole integration to the CUDA engine."""
import torch

from .base_classes import ODEblock
from .odeint import odeint, odeint_adjoint
from .utils import get_rw_adj, gcn_norm_fill_val


class ConstantODEblock(ODEblock):
  def __init__(self, odefunc, regularization_fns, opt, data, device, t=torch.tensor([0, 1])):
    super(ConstantODEblock, self).__init__(odefunc, regularization_fns, opt, data, device, t)
    self.aug_dim = 2 if opt['augment'] else 1
    # the reference builds a second function here (C.1): `reg_odefunc.odefunc` keeps the first one
    self.odefunc = odefunc(self.aug_dim * opt['hidden_dim'], self.aug_dim * opt['hidden_dim'], opt, data, device)
    normalise = get_rw_adj if opt['data_norm'] == 'rw' else gcn_norm_fill_val
    kwargs = dict(edge_weight=data.edge_attr, fill_value=opt['self_loop_weight'], num_nodes=data.num_nodes,
                  dtype=data.x.dtype)
    if opt['data_norm'] == 'rw':
      kwargs['norm_dim'] = 1
    edge_index, edge_weight = normalise(data.edge_index, **kwargs)
    self.odefunc.edge_index = edge_index.to(device)
    self.odefunc.edge_weight = edge_weight.to(device)
    self.reg_odefunc.odefunc.edge_index = self.odefunc.edge_index
    self.reg_odefunc.odefunc.edge_weight = self.odefunc.edge_weight
    self.train_integrator = odeint_adjoint if opt['adjoint'] else odeint
    self.test_integrator = odeint_adjoint if opt['adjoint'] else odeint
    self.set_tol()

  def forward(self, x):
    options = dict(step_size=self.opt['step_size'], max_iters=self.opt['max_iters'])
    adjoint_options = lambda: dict(step_size=self.opt['adjoint_step_size'], max_iters=self.opt['max_iters'])
    return self._integrate(x, options, adjoint_options)
