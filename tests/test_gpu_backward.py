"""Gradients of the CUDA path (RHS vector-Jacobian product kernel + solver loops) against autograd through the CPU
oracle in fp64.  Tolerance 2e-4 of the gradient tensor's max magnitude: the backward accumulates with fp32 atomics in a
different order than any reference, and fp32 autograd through the reference's own op graph differs from fp64 by the
same order."""
import pytest
import torch

import _cases as C
from __graft_entry__ import load_package
from oracle import grand_oracle as O

pytestmark = pytest.mark.gpu

load_package()
from graph_neural_pde_b200.block_constant import ConstantODEblock  # noqa: E402
from graph_neural_pde_b200.function_laplacian_diffusion import LaplacianODEFunc  # noqa: E402
from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt  # noqa: E402
from graph_neural_pde_b200.synthetic import Data, default_opt, make_graph  # noqa: E402

DEV = torch.device('cuda:0')
GTOL = 2e-4


def _make(function, n, d, ad, heads, method='rk4', adjoint=False, seed=0, **extra):
  edge_index = make_graph(n, 6 * n, max_degree=150, seed=seed)
  torch.manual_seed(seed + 1)
  x = torch.randn(n, d)
  opt = default_opt(function=function, block='constant', hidden_dim=d, attention_dim=ad, heads=heads, method=method,
                    step_size=0.5, time=1.0, add_source=True, adjoint=adjoint, adjoint_method='rk4',
                    adjoint_step_size=0.5, **extra)
  data = Data(x=x, edge_index=edge_index, edge_attr=None, num_nodes=n)
  cls = ODEFuncTransformerAtt if function == 'transformer' else LaplacianODEFunc
  block = ConstantODEblock(cls, (), opt, data, DEV, t=torch.tensor([0, opt['time']])).to(DEV)
  g = torch.Generator().manual_seed(seed + 2)
  with torch.no_grad():
    for p in block.parameters():
      if p.dim() == 2:
        p.copy_((1.5 * torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5).to(DEV))
    block.odefunc.alpha_train.fill_(0.3)
    block.odefunc.beta_train.fill_(-0.4)
  block.train()
  return block, opt, edge_index, x


def _oracle_params(block, dtype=torch.float64):
  sd = {k: v.detach().cpu().to(dtype).requires_grad_(v.dtype.is_floating_point) for k, v in block.state_dict().items()
        if v.dtype.is_floating_point}
  return sd


def _oracle_func(sd, opt, edge_index, n, x0, function):
  ei, ew = O.constant_block_graph(edge_index, None, n, opt, torch.float64)
  alpha, beta = sd['odefunc.alpha_train'], sd['odefunc.beta_train']
  if function == 'transformer':
    P = C.sub_state(sd, 'odefunc.multihead_att_layer.')
    return lambda t, y: O.transformer_rhs(y, ei, P, opt, alpha, beta, x0)
  return lambda t, y: O.laplacian_rhs(y, ei, ew, opt, alpha, beta, x0)


NAMES = {'transformer': ['odefunc.multihead_att_layer.Q.weight', 'odefunc.multihead_att_layer.Q.bias',
                         'odefunc.multihead_att_layer.K.weight', 'odefunc.multihead_att_layer.K.bias',
                         'odefunc.alpha_train', 'odefunc.beta_train'],
         'laplacian': ['odefunc.alpha_train', 'odefunc.beta_train']}


@pytest.mark.parametrize('function,d,ad,heads', [('transformer', 128, 16, 4), ('transformer', 32, 32, 8),
                                                 ('laplacian', 64, 16, 2)])
def test_rhs_vjp_vs_oracle_autograd(function, d, ad, heads):
  n = 400
  block, opt, edge_index, x = _make(function, n, d, ad, heads)
  func = block.odefunc
  xd = x.to(DEV).requires_grad_(True)
  block.set_x0(x.to(DEV))
  torch.manual_seed(9)
  gout = torch.randn(n, d)
  f = func(0.0, xd)
  f.backward(gout.to(DEV))
  # oracle
  sd = _oracle_params(block)
  xo = x.double().requires_grad_(True)
  fo = _oracle_func(sd, opt, edge_index, n, x.double(), function)(0.0, xo)
  fo.backward(gout.double())
  C.assert_close(f, fo, C.RTOL, 'forward value')
  C.assert_close(xd.grad, xo.grad, GTOL, 'dL/dx')
  got = dict(block.named_parameters())
  for name in NAMES[function]:
    C.assert_close(got[name].grad, sd[name].grad, GTOL, 'dL/d' + name)


@pytest.mark.parametrize('function', ['transformer', 'laplacian'])
@pytest.mark.parametrize('method', ['rk4', 'dopri5'])
def test_block_backward_vs_oracle_autograd(function, method):
  n, d = 300, 32
  block, opt, edge_index, x = _make(function, n, d, 16, 4, method=method, tol_scale=1000.0)
  xd = x.to(DEV).requires_grad_(True)
  block.set_x0(x.to(DEV))
  torch.manual_seed(3)
  w = torch.randn(n, d)
  z = block(xd)
  (z * w.to(DEV)).sum().backward()
  sd = _oracle_params(block)
  xo = x.double().requires_grad_(True)
  fo = _oracle_func(sd, opt, edge_index, n, x.double(), function)
  zo = O.block_forward(fo, xo, opt['time'], opt)
  (zo * w.double()).sum().backward()
  tol = GTOL if method == 'rk4' else 2e-2     # dopri5: different accepted-step sequences (see test_gpu_parity)
  C.assert_close(z, zo, C.RTOL if method == 'rk4' else 5e-3, 'z')
  C.assert_close(xd.grad, xo.grad, tol, 'dL/dx')
  got = dict(block.named_parameters())
  for name in NAMES[function]:
    C.assert_close(got[name].grad, sd[name].grad, tol, 'dL/d' + name)


def test_adjoint_gradients_close_to_direct_gradients():
  """odeint_adjoint (reverse-time rk4 of the augmented system) vs backprop through the forward stages: both
  discretise the same continuous sensitivity, so they agree to the step's truncation error, not to rounding."""
  n, d = 300, 32
  grads = {}
  for adjoint in (False, True):
    block, opt, edge_index, x = _make('transformer', n, d, 16, 4, method='rk4', adjoint=adjoint, seed=5)
    xd = x.to(DEV).requires_grad_(True)
    block.set_x0(x.to(DEV))
    torch.manual_seed(3)
    w = torch.randn(n, d).to(DEV)
    z = block(xd)
    (z * w).sum().backward()
    grads[adjoint] = [xd.grad.clone()] + [dict(block.named_parameters())[k].grad.clone() for k in NAMES['transformer']]
    nfe_after_backward = block.odefunc.nfe
    assert nfe_after_backward > 8 or not adjoint      # the adjoint pass evaluates the function again (bm meter)
  for a, b in zip(grads[False], grads[True]):
    C.assert_close(b, a, 5e-2, 'adjoint vs direct')
