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

  def _differentiable_attention(self, x):
    """Training without the adjoint: the reference backpropagates through the frozen attention into the block's Q / K
    and into x(0) (autograd over :28-39).  -> attention [E, H]; leaves (mean_csr with history, flipped-order copy) on
    the function.  Under the adjoint the reference's torchdiffeq only differentiates the function's own parameters, so
    the frozen weights are constants there -- as here."""
    from .autograd_rhs import FrozenAttentionFn, _transposed_engine
    layer = self.multihead_att_layer
    if not hasattr(layer, 'Q'):
      raise NotImplementedError('grand_b200 backward: the beltrami exp_kernel layer is forward-only')
    eng = self.odefunc.engine(x)
    frozen, mean_csr = FrozenAttentionFn.apply(x, layer, eng, self.odefunc.edge_index, *layer.parameters())
    eng_t = _transposed_engine(self.odefunc, x)
    self.odefunc.frozen_weights = (mean_csr, eng_t.edge_values_to_csr(frozen))
    return frozen

  def forward(self, x):
    self.odefunc.frozen_weights = None
    wants_grad = torch.is_grad_enabled() and (x.requires_grad or any(
      p.requires_grad for p in self.multihead_att_layer.parameters()))
    if wants_grad and not (self.opt['adjoint'] and self.training):
      frozen = self._differentiable_attention(x)
    else:
      frozen = self.get_attention_weights(x.detach())
    self.odefunc.attention_weights = frozen
    self.reg_odefunc.odefunc.attention_weights = frozen
    o = self.opt
    return self.integrate(x, solver_options={'step_size': o['step_size']},
                          adjoint_options=lambda: {'step_size': o['adjoint_step_size']})
