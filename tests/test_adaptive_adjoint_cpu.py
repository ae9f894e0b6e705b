"""Adaptive adjoint solvers (graph-neural-pde_b200/autograd_rhs.py: integrate_system_adaptive, OdeintAdjointFn with
adjoint_method dopri5 / adaptive_heun / bosh3) on the CPU with a small differentiable stand-in for the RHS.  The
reference's default adjoint method is adaptive_heun (src/run_GNN.py:334).  Checked against: the torchdiffeq restatement
under oracle/shims (dopri5, single-tensor state: same accepted steps -> same NFE, same end state), closed-form
solutions, and fp64 autograd through a tight forward solve for the adjoint gradients."""
import math
import os
import sys

import pytest
import torch

from __graft_entry__ import load_package

load_package()
from graph_neural_pde_b200 import autograd_rhs as AG  # noqa: E402

SHIMS = os.path.join(os.path.dirname(__file__), '..', 'oracle', 'shims')


def _shim_odeint():
  if SHIMS not in sys.path:
    sys.path.insert(0, SHIMS)
  from torchdiffeq import odeint
  return odeint


class Counting:
  def __init__(self, a):
    self.a, self.n = a, 0

  def __call__(self, t, y):
    self.n += 1
    return torch.tanh(y @ self.a) - 0.3 * y


def test_dopri5_single_tensor_matches_torchdiffeq_restatement():
  torch.manual_seed(0)
  a = torch.randn(6, 6, dtype=torch.float64) * 0.7
  y0 = torch.randn(5, 6, dtype=torch.float64)
  ref_f = Counting(a)
  want = _shim_odeint()(ref_f, y0, torch.tensor([0.0, 2.5], dtype=torch.float64), rtol=1e-6, atol=1e-8, method='dopri5')[1]
  f = Counting(a)
  stats = {}
  got = AG.integrate_system_adaptive(lambda t, s: [f(t, s[0])], [y0], 0.0, 2.5, 'dopri5', 1e-6, 1e-8, stats=stats)[0]
  assert f.n == ref_f.n == stats['n_eval']
  assert float((got - want).abs().max()) < 1e-12


@pytest.mark.parametrize('method,tol', [('dopri5', 2e-6), ('bosh3', 2e-4), ('adaptive_heun', 2e-3)])
def test_adaptive_methods_solve_a_linear_system_in_both_directions(method, tol):
  lam = torch.tensor([-0.5, -2.0, 0.25], dtype=torch.float64)
  y0 = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
  f = lambda t, s: [lam * s[0], -s[1]]        # two components: the mixed norm path
  for ta, tb in ((0.0, 1.5), (1.5, 0.0)):
    stats = {}
    out = AG.integrate_system_adaptive(f, [y0, y0.clone()], ta, tb, method, 1e-7, 1e-9, stats=stats)
    want0, want1 = y0 * torch.exp(lam * (tb - ta)), y0 * math.exp(-(tb - ta))
    assert float(((out[0] - want0) / want0).abs().max()) < tol, (method, ta, tb)
    assert float(((out[1] - want1) / want1).abs().max()) < tol
    assert stats['n_steps'] >= 1


class Func(torch.nn.Module):
  def __init__(self):
    super().__init__()
    torch.manual_seed(1)
    self.w = torch.nn.Parameter(torch.randn(4, 4, dtype=torch.float64) * 0.5)
    self.b = torch.nn.Parameter(torch.randn(4, dtype=torch.float64) * 0.1)
    self.nfe = 0

  def forward(self, t, y):
    self.nfe += 1
    return torch.tanh(y @ self.w + self.b) - 0.2 * y


def _rk4_fine(func, y, t0, t1, n):
  h = (t1 - t0) / n
  for _ in range(n):
    k1 = func(0, y)
    k2 = func(0, y + 0.5 * h * k1)
    k3 = func(0, y + 0.5 * h * k2)
    k4 = func(0, y + h * k3)
    y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
  return y


@pytest.mark.parametrize('method,tol', [('dopri5', 1e-6), ('adaptive_heun', 2e-4), ('bosh3', 2e-5), ('rk4', 1e-5)])
def test_adjoint_gradients_vs_autograd_through_a_fine_solve(method, tol):
  func = Func()
  torch.manual_seed(2)
  y0 = torch.randn(7, 4, dtype=torch.float64, requires_grad=True)
  t = torch.tensor([0.0, 1.2], dtype=torch.float64)
  c = torch.randn(7, 4, dtype=torch.float64)
  # arbiter: backprop through 400 classical rk4 steps
  (c * _rk4_fine(func, y0, 0.0, 1.2, 400)).sum().backward()
  want = [y0.grad.clone(), func.w.grad.clone(), func.b.grad.clone()]
  y0.grad = None
  func.zero_grad()

  def fwd(y):
    return torch.stack([y, _rk4_fine(func, y, 0.0, 1.2, 400)], 0)

  params = tuple(func.parameters())
  sol = AG.OdeintAdjointFn.apply(y0, func, t, fwd, method, {'step_size': 0.05, 'max_iters': 100}, 1e-8, 1e-10, *params)
  nfe0 = func.nfe
  (c * sol[1]).sum().backward()
  assert func.nfe > nfe0
  got = [y0.grad, func.w.grad, func.b.grad]
  for g, w in zip(got, want):
    assert float((g - w).abs().max() / w.abs().max()) < tol, method


def test_adjoint_resets_the_state_to_the_forward_solution_at_interior_times():
  """torchdiffeq's adjoint loop: `aug_state[1] = y[i - 1]` after every interval; gradients of a loss on an interior
  time and the end time."""
  func = Func()
  torch.manual_seed(3)
  y0 = torch.randn(3, 4, dtype=torch.float64, requires_grad=True)
  t = torch.tensor([0.0, 0.5, 1.2], dtype=torch.float64)
  ya = _rk4_fine(func, y0, 0.0, 0.5, 200)
  yb = _rk4_fine(func, ya, 0.5, 1.2, 200)
  (ya.sum() + 2 * yb.sum()).backward()
  want = [y0.grad.clone(), func.w.grad.clone()]
  y0.grad = None
  func.zero_grad()

  def fwd(y):
    a = _rk4_fine(func, y, 0.0, 0.5, 200)
    return torch.stack([y, a, _rk4_fine(func, a, 0.5, 1.2, 200)], 0)

  sol = AG.OdeintAdjointFn.apply(y0, func, t, fwd, 'dopri5', {}, 1e-9, 1e-11, *tuple(func.parameters()))
  (sol[1].sum() + 2 * sol[2].sum()).backward()
  assert float((y0.grad - want[0]).abs().max() / want[0].abs().max()) < 1e-6
  assert float((func.w.grad - want[1]).abs().max() / want[1].abs().max()) < 1e-6


def test_unknown_adjoint_method_raises_like_torchdiffeq():
  with pytest.raises(ValueError, match='Invalid'):
    AG._integrate_system(lambda t, s: s, [torch.zeros(2)], 0.0, 1.0, 'tsit5', {}, 1e-3, 1e-3)
