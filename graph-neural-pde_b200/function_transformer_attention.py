"""GRAND-nl ODE function and its attention layer on the CUDA engine; API of reference
src/function_transformer_attention.py (ODEFuncTransformerAtt :12-56, SpGraphTransAttentionLayer :59-217).

`ODEFuncTransformerAtt.forward(t, x)` is ONE C-ABI call (grand_rhs_attention): Q/K projection, per-edge scores,
segment softmax / squareplus, head mean, attention-weighted aggregation and the alpha (ax - x) + beta x0 residual.
`SpGraphTransAttentionLayer.forward(x, edge)` returns `(attention[E,H], (v, prods))` in the caller's edge order via
grand_attention_weights.
"""
import torch
from torch import nn

from . import _native as N
from .autograd_rhs import TransformerRhsFn, TransformerRhsGeneralFn, needs_grad
from .base_classes import ODEFunc
from .engine import AttentionSpec, RhsEngine
from .graph_plan import GraphPlan
from .utils import add_remaining_self_loops


_MIX_FEATURES = ('mix_features=True is not built: the reference itself cannot run it for this function (its forward passes '
                 'the layer\'s (v, prods) tuple to multiply_attention, whose v.shape raises -- reference '
                 'src/function_transformer_attention.py:43-44, :29), so there is no behaviour to reproduce')


class SpGraphTransAttentionLayer(nn.Module):
  """Parameters, their names and their initialisation follow the reference (:66-126): Q, K, V, Wout Linear layers
  with weight = 1e-5 and default bias; exp_kernel scalars; the beltrami + exp_kernel variant owns Qx/Kx/Vx/Qp/Kp/Vp.
  V and Wout are parameters of the state_dict only: the mix_features=False path never uses them."""

  def __init__(self, in_features, out_features, opt, device, concat=True, edge_weights=None):
    super(SpGraphTransAttentionLayer, self).__init__()
    self.in_features = in_features
    self.out_features = out_features
    self.alpha = opt['leaky_relu_slope']
    self.concat = concat
    self.device = device
    self.opt = opt
    self.h = int(opt['heads'])
    self.edge_weights = edge_weights
    self.attention_dim = opt['attention_dim'] if 'attention_dim' in opt else out_features
    assert self.attention_dim % self.h == 0, "Number of heads ({}) must be a factor of the dimension size ({})".format(
      self.h, self.attention_dim)
    self.d_k = self.attention_dim // self.h

    def linear(n_in, n_out):
      m = nn.Linear(n_in, n_out)
      self.init_weights(m)
      return m

    if opt['beltrami'] and opt['attention_type'] == "exp_kernel":
      for name in ('output_var_x', 'lengthscale_x', 'output_var_p', 'lengthscale_p'):
        setattr(self, name, nn.Parameter(torch.ones(1)))
      feat = opt['hidden_dim'] - opt['pos_enc_hidden_dim']
      self.Qx, self.Vx, self.Kx = (linear(feat, self.attention_dim) for _ in range(3))
      pos = opt['pos_enc_hidden_dim']
      self.Qp, self.Vp, self.Kp = (linear(pos, self.attention_dim) for _ in range(3))
    else:
      if opt['attention_type'] == "exp_kernel":
        self.output_var = nn.Parameter(torch.ones(1))
        self.lengthscale = nn.Parameter(torch.ones(1))
      self.Q = linear(in_features, self.attention_dim)
      self.V = linear(in_features, self.attention_dim)
      self.K = linear(in_features, self.attention_dim)
    self.activation = nn.Sigmoid()
    self.Wout = linear(self.d_k, in_features)
    self._plan_cache = None

  def init_weights(self, m):
    if type(m) == nn.Linear:
      nn.init.constant_(m.weight, 1e-5)

  def packed_beltrami_weights(self):
    """[Qx | Qp] and [Kx | Kp] as block-diagonal [2 ad, d] matrices over the full state row, so that one projection
    pass yields both halves.  Column selection follows reference :134-136: the positional slice is
    [feat_hidden_dim, feat_hidden_dim + pos_enc_hidden_dim), the feature slice is everything else."""
    opt = self.opt
    f, p = opt['feat_hidden_dim'], opt['pos_enc_hidden_dim']
    d = self.in_features
    ad = self.attention_dim
    feat_cols = list(range(0, f)) + list(range(f + p, d))
    pos_cols = list(range(f, f + p))

    def pack(lin_x, lin_p):
      w = torch.zeros((2 * ad, d), dtype=lin_x.weight.dtype, device=lin_x.weight.device)
      w[:ad, feat_cols] = lin_x.weight.detach()
      w[ad:, pos_cols] = lin_p.weight.detach()
      b = torch.cat([lin_x.bias.detach(), lin_p.bias.detach()])
      return w.contiguous(), b.contiguous()

    wq, bq = pack(self.Qx, self.Qp)
    wk, bk = pack(self.Kx, self.Kp)
    return wq, bq, wk, bk

  def _engine_for(self, x, edge):
    """Standalone use (AttODEblock, tests): cache one plan per edge tensor identity."""
    key = (edge.data_ptr(), tuple(edge.shape), x.shape[0], x.shape[1], str(x.device))
    if self._plan_cache is None or self._plan_cache[0] != key:
      plan = GraphPlan(edge, num_nodes=x.shape[0], device=x.device)
      self._plan_cache = (key, RhsEngine(plan, x.shape[1]), edge)
    return self._plan_cache[1]

  def reweight_csr(self, engine):
    """reweight_attention multiplies the scores by the layer's ctor-time edge weights (reference :208-209)."""
    if self.opt['reweight_attention'] and self.edge_weights is not None:
      w = self.edge_weights.to(device=engine.device, dtype=torch.float32)
      if w.shape[0] != engine.plan.n_edges:
        raise RuntimeError('The size of tensor a (%d) must match the size of tensor b (%d) at non-singleton dimension 0'
                           % (engine.plan.n_edges, w.shape[0]))
      return engine.edge_values_to_csr(w)
    return None

  def spec(self, engine):
    return AttentionSpec(self, engine.d, self.reweight_csr(engine))

  def forward(self, x, edge, engine=None):
    """-> (attention[E, H], (v, prods[E, H])) with rows in the order of `edge` (reference :128-214).  `v` is None:
    the V projection is dead on the mix_features=False path and is not computed."""
    N.require_cuda(x, 'x')
    if x.requires_grad and torch.is_grad_enabled():
      raise NotImplementedError('grand_b200: the standalone attention layer is forward-only (gradients flow through '
                                'ODEFuncTransformerAtt.forward, not through frozen attention weights); call under '
                                'torch.no_grad()')
    engine = engine if engine is not None else self._engine_for(x, edge)
    attention, prods, _ = engine.attention_weights(self.spec(engine), x.contiguous())
    return attention, (None, prods)

  def __repr__(self):
    return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class ODEFuncTransformerAtt(ODEFunc):

  def __init__(self, in_features, out_features, opt, data, device):
    super(ODEFuncTransformerAtt, self).__init__(opt, data, device)
    self.in_features = in_features
    self.out_features = out_features
    if opt['self_loop_weight'] > 0:
      self.edge_index, self.edge_weight = add_remaining_self_loops(data.edge_index, data.edge_attr,
                                                                   fill_value=opt['self_loop_weight'])
    else:
      self.edge_index, self.edge_weight = data.edge_index, data.edge_attr
    self.multihead_att_layer = SpGraphTransAttentionLayer(in_features, out_features, opt, device,
                                                          edge_weights=self.edge_weight).to(device)

  def residual_epilogue(self, engine, x0_buffer, **outputs):
    x0 = x0_buffer if self.opt['add_source'] else None
    return RhsEngine.epilogue(self.alpha_train.detach(), self.beta_train.detach(), x0, self.opt['no_alpha_sigmoid'],
                              **outputs)

  def native_description(self, engine):
    """(attention spec, None): what the solvers hand to the kernels instead of calling forward per stage."""
    if self.opt['mix_features']:
      raise NotImplementedError(_MIX_FEATURES)
    return self.multihead_att_layer.spec(engine), None

  def stage_x0(self, engine):
    """Copy the block-assigned x0 into the engine's persistent buffer (stable address for graph replay)."""
    if not self.opt['add_source']:
      return None
    buf = engine.buffer('x0', rows=engine.plan.n_rows)
    buf.copy_(self.x0)
    return buf

  def forward(self, t, x):  # t is needed when called by the integrator
    if self.opt['mix_features']:
      raise NotImplementedError(_MIX_FEATURES)
    self._bump_nfe()
    N.require_cuda(x, 'x')
    if needs_grad(self, x):
      layer = self.multihead_att_layer
      if not hasattr(layer, 'Q'):
        raise NotImplementedError('grand_b200 backward: the beltrami exp_kernel layer is forward-only')
      if layer.spec(self.engine(x)).fused and x.shape[1] % 4 == 0:
        return TransformerRhsFn.apply(x, layer.Q.weight, layer.Q.bias, layer.K.weight, layer.K.bias, self.alpha_train,
                                      self.beta_train, self)
      return TransformerRhsGeneralFn.apply(x, self.alpha_train, self.beta_train, self, *layer.parameters())
    return self.native_forward(x.contiguous())

  def native_forward(self, x):
    """The evaluation itself, one C call (no autograd graph)."""
    engine = self.engine(x)
    f = torch.empty_like(x)
    ep = self.residual_epilogue(engine, self.stage_x0(engine), f_out=f)
    engine.rhs_attention(self.multihead_att_layer.spec(engine), x, ep)
    return f

  def torch_rhs(self, x, x0):
    """Differentiable torch restatement of the same evaluation (backward of the unfused configurations)."""
    from .attention_autograd import transformer_rhs
    return transformer_rhs(x, self, x0)

  def __repr__(self):
    return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'
