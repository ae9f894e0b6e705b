"""GAT-scored diffusion function on the CUDA engine; API of reference src/function_GAT_attention.py (ODEFuncAtt :11-67,
SpGraphAttentionLayer :70-119), one of the "next" rows of SURVEY.md 8(f).4.

The additive scorer  e_ij = LeakyReLU(a_src . h_i + a_dst . h_j),  h = x W split into heads, is score type
GRAND_SCORE_GAT of the general attention pipeline (csrc/attention.cuh): Q = K = x W, the vector `a` and the slope travel
as `kernel_scalars`.  With mix_features off the reference averages H separate spmm's (:38-42); by linearity that is the
spmm of the head-mean attention, i.e. the same aggregation kernel the transformer function uses.  With mix_features on
(:31-37) the same aggregation runs on x W (width attention_dim) between two dense layers (grand_linear).
"""
import torch
from torch import nn

from . import _native as N
from .autograd_rhs import needs_grad
from .base_classes import ODEFunc
from .engine import AttentionSpec, RhsEngine
from .graph_plan import GraphPlan
from .utils import add_remaining_self_loops


class SpGraphAttentionLayer(nn.Module):
  """Sparse GAT layer (https://arxiv.org/abs/1710.10903); parameters W [in, ad], Wout [ad, in], a [2 d_k, 1, 1] with
  xavier_normal_(gain=1.414) as in the reference (:92-99)."""

  def __init__(self, in_features, out_features, opt, device, concat=True):
    super(SpGraphAttentionLayer, self).__init__()
    self.in_features, self.out_features = in_features, out_features
    self.alpha = opt['leaky_relu_slope']
    self.concat, self.device, self.opt = concat, device, opt
    self.h = int(opt['heads'])
    self.attention_dim = opt['attention_dim'] if 'attention_dim' in opt else out_features
    assert self.attention_dim % self.h == 0, "Number of heads must be a factor of the dimension size"
    self.d_k = self.attention_dim // self.h
    for name, shape in (('W', (in_features, self.attention_dim)), ('Wout', (self.attention_dim, in_features)),
                        ('a', (2 * self.d_k, 1, 1))):
      p = nn.Parameter(torch.zeros(size=shape))
      nn.init.xavier_normal_(p.data, gain=1.414)
      setattr(self, name, p)
    self.leakyrelu = nn.LeakyReLU(self.alpha)
    self._plan_cache = None

  def native_weights(self):
    """What engine.AttentionSpec needs: (scorer, wq, bq, wk, bk, kernel_scalars) -- Q and K are both x W."""
    w = self.W.detach().t().contiguous()                                   # nn.Linear layout [ad, d]
    zero = torch.zeros(self.attention_dim, dtype=w.dtype, device=w.device)
    slope = torch.full((1,), float(self.alpha), dtype=w.dtype, device=w.device)
    return 'gat', w, zero, w, zero, torch.cat([self.a.detach().reshape(-1), slope])

  def spec(self, engine):
    return AttentionSpec(self, engine.d, None)

  def _engine_for(self, x, edge):
    key = (edge.data_ptr(), tuple(edge.shape), x.shape[0], x.shape[1], str(x.device))
    if self._plan_cache is None or self._plan_cache[0] != key:
      self._plan_cache = (key, RhsEngine(GraphPlan(edge, num_nodes=x.shape[0], device=x.device), x.shape[1]), edge)
    return self._plan_cache[1]

  def forward(self, x, edge, engine=None):
    """-> (attention[E, H] in the order of `edge`, wx = x W)   (reference :105-116)"""
    N.require_cuda(x, 'x')
    if x.requires_grad and torch.is_grad_enabled():
      raise NotImplementedError('grand_b200: the GAT scorer is forward-only; call under torch.no_grad()')
    engine = engine if engine is not None else self._engine_for(x, edge)
    attention, _, _ = engine.attention_weights(self.spec(engine), x.contiguous())
    return attention, torch.mm(x, self.W.detach())

  def __repr__(self):
    return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class ODEFuncAtt(ODEFunc):

  def __init__(self, in_features, out_features, opt, data, device):
    super(ODEFuncAtt, self).__init__(opt, data, device)
    if opt['self_loop_weight'] > 0:
      self.edge_index, self.edge_weight = add_remaining_self_loops(data.edge_index, data.edge_attr,
                                                                   fill_value=opt['self_loop_weight'])
    else:
      self.edge_index, self.edge_weight = data.edge_index, data.edge_attr
    self.multihead_att_layer = SpGraphAttentionLayer(in_features, out_features, opt, device).to(device)
    self.attention_dim = self.multihead_att_layer.attention_dim
    self.d_k = self.multihead_att_layer.d_k
    self.in_features, self.out_features = in_features, out_features

  def per_call_only(self):
    """mix_features (reference :31-37) is three launches around two dense layers per evaluation -- attention weights,
    aggregation of x W, product with Wout -- not the single fused RHS launch: the solvers then drive `forward` stage by
    stage (odeint.odeint) instead of the fused integration."""
    return bool(self.opt['mix_features'])

  def _mix_features_rhs(self, x):
    """f = alpha (mean_h(A_h (x W)) Wout - x) + beta x0.  By linearity the mean over the heads' spmm's is ONE aggregation
    with the head-mean attention (the weights the Laplacian kernel takes), run at width attention_dim."""
    from .linear import linear_forward
    layer = self.multihead_att_layer
    engine = self.engine(x)
    _, _, mean_csr = engine.attention_weights(layer.spec(engine), x, want_attention=False, want_prods=False,
                                              want_mean_csr=True)
    wx = linear_forward(x, layer.W.detach().t().contiguous(), None)                      # [N, ad]   (:106)
    cached = getattr(self, '_engine_ad', None)
    if cached is None or cached[0] is not engine:
      cached = (engine, RhsEngine(engine.plan, layer.attention_dim))
      self._engine_ad = cached
    one = torch.ones((), dtype=torch.float32, device=x.device)
    r = torch.empty_like(wx)
    cached[1].rhs_laplacian(mean_csr, wx, RhsEngine.epilogue(one, None, None, True, f_out=r))   # r = A wx - wx
    ax = linear_forward(r + wx, layer.Wout.detach().t().contiguous(), None)              # [N, d]    (:36)
    alpha = self.alpha_train.detach() if self.opt['no_alpha_sigmoid'] else torch.sigmoid(self.alpha_train.detach())
    f = alpha * (ax - x)
    if self.opt['add_source']:
      f = f + self.beta_train.detach() * self.x0
    return f

  def residual_epilogue(self, engine, x0_buffer, **outputs):
    x0 = x0_buffer if self.opt['add_source'] else None
    return RhsEngine.epilogue(self.alpha_train.detach(), self.beta_train.detach(), x0, self.opt['no_alpha_sigmoid'],
                              **outputs)

  def native_description(self, engine):
    assert not self.opt['mix_features'], 'mix_features runs stage by stage (per_call_only)'
    return self.multihead_att_layer.spec(engine), None

  def stage_x0(self, engine):
    if not self.opt['add_source']:
      return None
    buf = engine.buffer('x0', rows=engine.plan.n_rows)
    buf.copy_(self.x0)
    return buf

  def forward(self, t, x):  # t is needed when called by the integrator
    self._bump_nfe()
    N.require_cuda(x, 'x')
    if needs_grad(self, x):
      from .autograd_rhs import GeneralRhsFn
      return GeneralRhsFn.apply(x, self.alpha_train, self.beta_train, self, *self.multihead_att_layer.parameters())
    return self.native_forward(x.contiguous())

  def native_forward(self, x):
    """The evaluation itself on the native library (no autograd graph)."""
    if self.opt['mix_features']:
      return self._mix_features_rhs(x)
    engine = self.engine(x)
    f = torch.empty_like(x)
    engine.rhs_attention(self.multihead_att_layer.spec(engine), x,
                         self.residual_epilogue(engine, self.stage_x0(engine), f_out=f))
    return f

  def torch_rhs(self, x, x0):
    """Differentiable torch restatement of the same evaluation (the GAT scorer has no backward kernel)."""
    from .attention_autograd import gat_rhs
    return gat_rhs(x, self, x0)

  def __repr__(self):
    return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'
