"""ODEFunc / ODEblock / BaseGNN with the reference's constructor signatures, attribute names and state_dict keys
(reference src/base_classes.py:32-146), rebuilt around the CUDA engine.

Differences a caller can observe: none on the API; the classes are plain nn.Modules (the reference subclasses PyG's
MessagePassing but never calls propagate) and `ODEFunc` owns a lazily built device `GraphPlan` that is
invalidated whenever `edge_index` is reassigned from outside (reference src/run_GNN.py:252-254 does that).
"""
import torch
from torch import nn

from .engine import RhsEngine
from .graph_plan import GraphPlan
from .linear import NativeLinear
from .utils import Meter


class RegularizedODEfunc(nn.Module):
  """Holder that keeps `ODEblock.reg_odefunc.odefunc` (and its parameters / state_dict keys) alive
  (reference src/regularized_ODE_function.py:8-33).  The RNODE regularisers themselves are out of the hot path's
  scope (all coefficients default to None, reference src/run_GNN.py:370-374)."""

  def __init__(self, odefunc, regularization_fns):
    super().__init__()
    self.odefunc = odefunc
    self.regularization_fns = regularization_fns

  def before_odeint(self, *args, **kwargs):
    return self.odefunc.before_odeint(*args, **kwargs) if hasattr(self.odefunc, 'before_odeint') else None

  def forward(self, t, state):
    if len(self.regularization_fns) > 0:
      raise NotImplementedError('RNODE regularisers are outside the grand_b200 hot path')
    return self.odefunc(t, state)

  @property
  def _num_evals(self):
    return self.odefunc._num_evals


REGULARIZATION_KEYS = ('kinetic_energy', 'jacobian_norm2', 'total_deriv', 'directional_penalty')


def create_regularization_fns(args):
  """reference src/base_classes.py:18-29; any non-None coefficient is refused loudly (out of scope)."""
  fns, coeffs = [], []
  for key in REGULARIZATION_KEYS:
    if args.get(key) is not None:
      raise NotImplementedError('regulariser %r is outside the grand_b200 hot path' % key)
  return fns, coeffs


class ODEFunc(nn.Module):
  """reference src/base_classes.py:77-95.  Mutable attributes assigned from outside: edge_index, edge_weight,
  attention_weights, x0, nfe."""

  def __init__(self, opt, data, device):
    super(ODEFunc, self).__init__()
    self.opt = opt
    self.device = device
    self._edge_index = None
    self._engine = None
    self.edge_weight = None
    self.attention_weights = None
    self.alpha_train = nn.Parameter(torch.tensor(0.0))
    self.beta_train = nn.Parameter(torch.tensor(0.0))
    self.x0 = None
    self.nfe = 0
    self.alpha_sc = nn.Parameter(torch.ones(1))
    self.beta_sc = nn.Parameter(torch.ones(1))

  # edge_index is a plain attribute in the reference; here assigning it drops the cached CSR plan
  @property
  def edge_index(self):
    return self._edge_index

  @edge_index.setter
  def edge_index(self, value):
    self._edge_index = value
    self._engine = None

  def engine(self, x):
    """The CUDA engine for the current graph and state width (built on first use).  The engines of the last TWO edge
    lists are kept: the hard-attention block alternates between the full graph (evaluation) and a freshly sampled one
    (every training forward, reference src/block_transformer_hard_attention.py:49-68), and the full graph's plan should not
    be rebuilt after every training step.  An entry is only reused for the very tensor object it was built from (the
    entry holds a reference to it, so its identity cannot be recycled)."""
    d = int(x.shape[1])
    if self._engine is None or self._engine.d != d or self._engine.device != x.device:
      cache = self.__dict__.setdefault('_engine_cache', [])
      for i, (ei, eng) in enumerate(cache):
        if ei is self._edge_index and eng.d == d and eng.device == x.device:
          cache.append(cache.pop(i))                      # most recently used last
          self._engine = eng
          break
      else:
        plan = GraphPlan(self._edge_index, num_nodes=x.shape[0], device=x.device)
        self._engine = RhsEngine(plan, d)
        cache.append((self._edge_index, self._engine))
        del cache[:-2]
    return self._engine

  def _bump_nfe(self, count=1):
    """`if self.nfe > max_nfe: raise` BEFORE the increment, once per evaluation
    (reference src/function_transformer_attention.py:39-42)."""
    from .utils import MaxNFEException
    for _ in range(count):
      if self.nfe > self.opt['max_nfe']:
        raise MaxNFEException
      self.nfe += 1

  def __repr__(self):
    return self.__class__.__name__


class ODEblock(nn.Module):
  """reference src/base_classes.py:32-74."""

  def __init__(self, odefunc, regularization_fns, opt, data, device, t):
    super(ODEblock, self).__init__()
    self.opt = opt
    self.t = t
    self.device = device
    self.aug_dim = 2 if opt['augment'] else 1
    self.odefunc = odefunc(self.aug_dim * opt['hidden_dim'], self.aug_dim * opt['hidden_dim'], opt, data, device)
    self.nreg = len(regularization_fns)
    self.reg_odefunc = RegularizedODEfunc(self.odefunc, regularization_fns)
    from .odeint import odeint, odeint_adjoint
    self.train_integrator = odeint_adjoint if opt['adjoint'] else odeint
    self.test_integrator = None
    self.set_tol()

  def set_x0(self, x0):
    self.odefunc.x0 = x0.clone().detach()
    self.reg_odefunc.odefunc.x0 = x0.clone().detach()

  def set_tol(self):
    self.atol = self.opt['tol_scale'] * 1e-7
    self.rtol = self.opt['tol_scale'] * 1e-9
    if self.opt['adjoint']:
      self.atol_adjoint = self.opt['tol_scale_adjoint'] * 1e-7
      self.rtol_adjoint = self.opt['tol_scale_adjoint'] * 1e-9

  def reset_tol(self):
    self.atol = 1e-7
    self.rtol = 1e-9
    self.atol_adjoint = 1e-7
    self.rtol_adjoint = 1e-9

  def set_time(self, time):
    self.t = torch.tensor([0, time]).to(self.device)

  # ---- helpers shared by the concrete blocks -------------------------------------------------------------------
  def rebuild_function(self, odefunc_cls, data):
    """The reference's blocks construct a second function after the base class made one; the first
    stays alive inside `reg_odefunc` with its own parameters and state_dict keys."""
    width = self.aug_dim * self.opt['hidden_dim']
    self.odefunc = odefunc_cls(width, width, self.opt, data, self.device)

  def install_graph(self, edge_index, edge_weight):
    """Hand the normalised graph to both function instances (reference src/block_constant.py:22-24)."""
    self.odefunc.edge_index = edge_index.to(self.device)
    self.odefunc.edge_weight = edge_weight.to(self.device)
    self.reg_odefunc.odefunc.edge_index = self.odefunc.edge_index
    self.reg_odefunc.odefunc.edge_weight = self.odefunc.edge_weight

  def use_default_integrators(self):
    from .odeint import odeint, odeint_adjoint
    chosen = odeint_adjoint if self.opt['adjoint'] else odeint
    self.train_integrator = chosen
    self.test_integrator = chosen
    self.set_tol()

  def integrate(self, x, solver_options, adjoint_options):
    """Tail of every block's forward (reference src/block_constant.py:35-70, src/block_transformer_attention.py:36-72):
    pick the integrator by mode, call it with the reference's keyword set, return the state at t[1].
    `adjoint_options` is a callable so that options the reference only reads in adjoint training are only read then."""
    if self.training and self.nreg > 0:
      raise NotImplementedError('RNODE regularisers are outside the grand_b200 hot path')
    integrator = self.train_integrator if self.training else self.test_integrator
    kwargs = dict(method=self.opt['method'], options=solver_options, atol=self.atol, rtol=self.rtol)
    if self.opt['adjoint'] and self.training:
      kwargs.update(adjoint_method=self.opt['adjoint_method'], adjoint_options=adjoint_options(),
                    adjoint_atol=self.atol_adjoint, adjoint_rtol=self.rtol_adjoint)
    trajectory = integrator(self.odefunc, x, self.t.type_as(x), **kwargs)
    return trajectory[1]

  def __repr__(self):
    return self.__class__.__name__ + '( Time Interval ' + str(self.t[0].item()) + ' -> ' + str(self.t[1].item()) \
           + ")"


class BaseGNN(nn.Module):
  """Encoder / decoder shell, reference src/base_classes.py:98-146 (state_dict keys m1, m2, mx, mp, ... kept)."""

  def __init__(self, opt, dataset, device=torch.device('cpu')):
    super(BaseGNN, self).__init__()
    self.opt = opt
    self.T = opt['time']
    self.num_classes = dataset.num_classes
    self.num_features = dataset.data.num_features
    self.num_nodes = dataset.data.num_nodes
    self.device = device
    self.fm = Meter()
    self.bm = Meter()
    # layers in the reference's creation order and under the reference's names (state_dict compatibility);
    # opt['hidden_dim'] is rewritten in place exactly where the reference rewrites it (src/base_classes.py:113,122)
    if opt['beltrami']:
      self._add_linear('mx', self.num_features, opt['feat_hidden_dim'])
      self._add_linear('mp', opt['pos_enc_dim'], opt['pos_enc_hidden_dim'])
      opt['hidden_dim'] = opt['feat_hidden_dim'] + opt['pos_enc_hidden_dim']
    else:
      self._add_linear('m1', self.num_features, opt['hidden_dim'])
    if opt['use_mlp']:
      for name in ('m11', 'm12'):
        self._add_linear(name, opt['hidden_dim'], opt['hidden_dim'])
    if opt['use_labels']:
      opt['hidden_dim'] += dataset.num_classes
    else:
      self.hidden_dim = opt['hidden_dim']
    if opt['fc_out']:
      self._add_linear('fc', opt['hidden_dim'], opt['hidden_dim'])
    self._add_linear('m2', opt['hidden_dim'], dataset.num_classes)
    if opt['batch_norm']:
      for name in ('bn_in', 'bn_out'):
        setattr(self, name, torch.nn.BatchNorm1d(opt['hidden_dim']))
    self.regularization_fns, self.regularization_coeffs = create_regularization_fns(self.opt)

  def _add_linear(self, name, n_in, n_out):
    setattr(self, name, NativeLinear(n_in, n_out))   # nn.Linear state_dict keys; forward on the native library

  def getNFE(self):
    return self.odeblock.odefunc.nfe + self.odeblock.reg_odefunc.odefunc.nfe

  def resetNFE(self):
    self.odeblock.odefunc.nfe = 0
    self.odeblock.reg_odefunc.odefunc.nfe = 0

  def reset(self):
    self.m1.reset_parameters()
    self.m2.reset_parameters()

  def __repr__(self):
    return self.__class__.__name__
