"""ConstantODEblock -- reference src/block_constant.py:6-74.

The graph is normalised once at construction and never changes: GRAND-nl when the function is ODEFuncTransformerAtt,
GRAND-l with random-walk (or symmetric) weights when it is LaplacianODEFunc.  `forward(x)` hands the whole integration
to the CUDA engine through the integrator chosen in `ODEblock.integrate`."""
import torch

from .base_classes import ODEblock
from .utils import gcn_norm_fill_val, get_rw_adj


class ConstantODEblock(ODEblock):
  def __init__(self, odefunc, regularization_fns, opt, data, device, t=torch.tensor([0, 1])):
    super(ConstantODEblock, self).__init__(odefunc, regularization_fns, opt, data, device, t)
    self.rebuild_function(odefunc, data)       # reference quirk C.1: a second function instance replaces the first
    common = dict(edge_weight=data.edge_attr, fill_value=opt['self_loop_weight'], num_nodes=data.num_nodes,
                  dtype=data.x.dtype)
    if opt['data_norm'] == 'rw':
      graph = get_rw_adj(data.edge_index, norm_dim=1, **common)
    else:
      graph = gcn_norm_fill_val(data.edge_index, **common)
    self.install_graph(*graph)
    self.use_default_integrators()

  def forward(self, x):
    o = self.opt
    return self.integrate(x, solver_options=dict(step_size=o['step_size'], max_iters=o['max_iters']),
                          adjoint_options=lambda: dict(step_size=o['adjoint_step_size'], max_iters=o['max_iters']))
