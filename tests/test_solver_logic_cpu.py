"""The adaptive solver's HOST logic (graph-neural-pde_b200/odeint.py: Dopri5 buffer rotation, controller, dense
output) run on the CPU with the device primitives (RHS launch, lincomb, error norm) replaced by torch stand-ins, and
compared with the oracle's dopri5.  This isolates controller / bookkeeping bugs from kernel bugs without a GPU."""
import copy
import math

import torch

import _cases as C
from __graft_entry__ import load_package
from oracle import grand_oracle as O

load_package()
from graph_neural_pde_b200 import odeint as OD  # noqa: E402


class FakeEngine:
  def __init__(self, n, d):
    self.n, self.d, self.bufs = n, d, {}

  def buffer(self, name, rows=None):
    rows = rows or self.n
    if name not in self.bufs or self.bufs[name].shape[0] != rows:
      self.bufs[name] = torch.zeros(rows, self.d)
    return self.bufs[name]

  def lincomb(self, out, terms):
    acc = torch.zeros_like(out)
    for t, c in terms:
      acc = acc + torch.tensor(c, dtype=torch.float32) * t
    out.copy_(acc)
    return out

  def error_sumsq(self, acc, terms, y0, y1, atol, rtol):
    err = torch.zeros_like(y0)
    for t, c in terms:
      err = err + torch.tensor(c, dtype=torch.float32) * t
    tol = torch.tensor(atol, dtype=torch.float32) + torch.tensor(rtol, dtype=torch.float32) * torch.max(y0.abs(), y1.abs())
    acc += ((err / tol).double() ** 2).sum()


class FakeProblem:
  def __init__(self, f, y0):
    self.f = f
    self.engine = FakeEngine(*y0.shape)
    self.n = y0.shape[0]

  def rhs(self, x, f_out=None, y_out=None, coef_f=0.0, terms=()):
    fx = self.f(0.0, x.clone())
    if y_out is not None:
      acc = torch.tensor(coef_f, dtype=torch.float32) * fx
      for t, c in terms:
        acc = acc + torch.tensor(c, dtype=torch.float32) * t
      y_out.copy_(acc)
    if f_out is not None:
      f_out.copy_(fx)


class FakeFunc:
  nfe = 0

  def _bump_nfe(self, count=1):
    self.nfe += count


def run_host_dopri5(f, y0, t, rtol, atol, monkeypatch):
  func = FakeFunc()
  monkeypatch.setattr(OD, 'NativeProblem', lambda fn, y: FakeProblem(f, y))
  out, solver = OD.integrate_dopri5(func, y0, t, rtol, atol)
  return out, solver, func


def test_host_dopri5_matches_oracle_on_golden_blocks(golden, monkeypatch):
  names = sorted(k for k, v in golden.items() if v['kind'] == 'block' and v['opt']['method'] == 'dopri5')
  assert len(names) >= 6
  for name in names:
    c = golden[name]
    opt = c['opt']
    # rebuild the oracle RHS closure exactly as _cases.oracle_block does
    holder = {}
    real_block_forward = O.block_forward

    def capture_block_forward(func, x, T, o, _real=real_block_forward):
      holder['func'] = func
      return _real(func, x, T, o)

    monkeypatch.setattr(O, 'block_forward', capture_block_forward)
    rhs, z, nfe, _ = C.oracle_block(c)
    monkeypatch.undo()
    f = holder['func']
    x = c['inputs']['x']
    atol, rtol = O.block_tolerances(opt)
    t = torch.tensor([0, opt['time']]).type_as(x)
    stats = O.Dopri5Stats()
    want = O.odeint_dopri5(f, x, t, rtol, atol, stats=stats)
    out, solver, func = run_host_dopri5(f, x, t, rtol, atol, monkeypatch)
    monkeypatch.undo()
    assert func.nfe == 2 + 6 * solver.attempts, name
    # The controller feeds on an error estimate that is rounding noise when the step is far inside the tolerance
    # (ratio ~1e-5), so two correct implementations may pick step sizes a few % apart and then land on different
    # dense-output points; the solution they return agrees to the solver's own accuracy, not to 1e-5.  Judge both
    # against a tight fp64 integration of the same ODE.
    exact = C.exact_block_solution(c, name)
    ref_err = C.rel_err(c['outputs']['z'], exact)
    got_err = C.rel_err(out[1], exact)
    assert got_err <= 1.5 * ref_err + 1e-5, '%s: host dopri5 %.2e vs exact, reference %.2e' % (name, got_err, ref_err)
    assert abs(solver.attempts - (stats.accepted + stats.rejected)) <= 2, name


def test_host_dopri5_interior_times_and_rejections(monkeypatch):
  # a stiff-ish linear system forces rejected steps; requested interior output times exercise the dense output
  torch.manual_seed(0)
  A = -torch.diag(torch.linspace(0.5, 60.0, 12))
  f = lambda t, y: y @ A.T
  y0 = torch.randn(5, 12)
  t = torch.tensor([0.0, 0.13, 0.5, 1.0])
  stats = O.Dopri5Stats()
  want = O.odeint_dopri5(f, y0, t, 1e-6, 1e-8, stats=stats)
  out, solver, func = run_host_dopri5(f, y0, t, 1e-6, 1e-8, monkeypatch)
  assert stats.rejected > 0 and solver.attempts == stats.accepted + stats.rejected
  C.assert_close(out, want, 5e-6, 'dense output')


def test_early_stop_device_scoring_matches_host_logic():
  """The sync-free scorer (running best kept in a device tensor, one read at the end) must reproduce the reference's
  per-step host logic (src/early_stop_solver.py:203-218, :172-176): same accuracies as python int/int, strict '>'."""
  from graph_neural_pde_b200.early_stop_solver import EarlyStopRK4

  class Split:
    pass

  torch.manual_seed(0)
  n, classes, d = 200, 5, 16
  for dataset, yshape in (('Cora', (n,)), ('ogbn-arxiv', (n, 1))):
    data = Split()
    data.y = torch.randint(0, classes, yshape)
    r = torch.rand(n)
    data.train_mask, data.val_mask, data.test_mask = r < 0.3, (r >= 0.3) & (r < 0.6), r >= 0.6
    solver = EarlyStopRK4(None, None, 1e-3, 1e-3, {'dataset': dataset, 'max_test_steps': 100}, step_size=1.0)
    solver.data, solver.m2_weight, solver.m2_bias = data, torch.randn(classes, d), torch.randn(classes)
    best = (0, 0, 0, 0)
    for step in range(12):
      y = torch.randn(n, d)
      accs = solver.evaluate(y)            # the reference's host path
      if accs[1] > best[1]:
        best = (accs[0], accs[1], accs[2], float(step + 1))
      solver._score(y, step + 1)
    solver._publish()
    assert best == (solver.best_train, solver.best_val, solver.best_test, solver.best_time), dataset
