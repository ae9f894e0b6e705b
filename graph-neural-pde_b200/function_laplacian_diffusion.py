"""GRAND-l ODE function on the CUDA engine; API of reference src/function_laplacian_diffusion.py:15-51.

`forward(t, x)` is one C-ABI call (grand_rhs_laplacian): (A x) with the per-edge weight the reference's
sparse_multiply picks by block type (:28-36), fused with the alpha (ax - x) + beta x0 residual (:43-51).
"""
import torch
from torch import nn

from . import _native as N
from .autograd_rhs import FrozenLaplacianRhsFn, LaplacianRhsFn, needs_grad
from .base_classes import ODEFunc
from .engine import RhsEngine


class LaplacianODEFunc(ODEFunc):

  # currently requires in_features = out_features
  def __init__(self, in_features, out_features, opt, data, device):
    super(LaplacianODEFunc, self).__init__(opt, data, device)
    self.in_features = in_features
    self.out_features = out_features
    # unused by forward, present in the reference's state_dict (:23-26) and kept for checkpoint compatibility
    self.w = nn.Parameter(torch.eye(opt['hidden_dim']))
    self.d = nn.Parameter(torch.zeros(opt['hidden_dim']) + 1)
    self.alpha_sc = nn.Parameter(torch.ones(1))
    self.beta_sc = nn.Parameter(torch.ones(1))
    self._w_csr = None
    self._w_key = None
    # AttODEblock in (non-adjoint) training: (mean_h attention in CSR order WITH autograd history, the same values in
    # the flipped graph's CSR order); None otherwise
    self.frozen_weights = None

  def edge_values(self):
    """Which tensor sparse_multiply would use (reference :28-36)."""
    if self.opt['block'] in ['attention', 'mixed', 'hard_attention']:
      return self.attention_weights      # [E, H] -> head mean for 'attention', [E] otherwise
    return self.edge_weight

  def weights_csr(self, engine):
    """The per-edge weight in CSR order; recomputed only when the source tensor (or the graph) changes.  The
    reference recomputes attention_weights.mean(dim=1) at every evaluation (:29-31); hoisting it is value-identical."""
    w = self.edge_values()
    key = (id(engine), w.data_ptr(), tuple(w.shape), w._version)
    if self._w_key != key:
      N.require_cuda(w, 'edge weights')
      if w.dim() == 2 and self.opt['block'] not in ['attention']:
        raise RuntimeError('attention_weights must be [E] for block %r' % self.opt['block'])
      self._w_csr = engine.edge_values_to_csr(w.detach())
      self._w_key = key
    return self._w_csr

  def native_description(self, engine):
    """(None, CSR weights): what the solvers hand to the kernels instead of calling forward per stage."""
    return None, self.weights_csr(engine)

  def residual_epilogue(self, engine, x0_buffer, **outputs):
    x0 = x0_buffer if self.opt['add_source'] else None
    return RhsEngine.epilogue(self.alpha_train.detach(), self.beta_train.detach(), x0, self.opt['no_alpha_sigmoid'],
                              **outputs)

  def stage_x0(self, engine):
    if not self.opt['add_source']:
      return None
    buf = engine.buffer('x0', rows=engine.plan.n_rows)
    buf.copy_(self.x0)
    return buf

  def forward(self, t, x):  # the t param is needed by the ODE solver.
    self._bump_nfe()
    N.require_cuda(x, 'x')
    if needs_grad(self, x):
      if self.frozen_weights is not None:
        w_csr, w_csr_t = self.frozen_weights
        return FrozenLaplacianRhsFn.apply(x, w_csr, self.alpha_train, self.beta_train, self, w_csr_t)
      return LaplacianRhsFn.apply(x, self.alpha_train, self.beta_train, self)
    x = x.contiguous()
    engine = self.engine(x)
    f = torch.empty_like(x)
    ep = self.residual_epilogue(engine, self.stage_x0(engine), f_out=f)
    engine.rhs_laplacian(self.weights_csr(engine), x, ep)
    return f
