import collections
import torch
from .interp import _interp_evaluate, _interp_fit
from .misc import _compute_error_ratio, _select_initial_step, _optimal_step_size, _rms_norm
from .solvers import AdaptiveStepsizeODESolver

_ButcherTableau = collections.namedtuple('_ButcherTableau', 'alpha, beta, c_sol, c_error')
_RungeKuttaState = collections.namedtuple('_RungeKuttaState', 'y1, f1, t0, t1, dt, interp_coeff')

_one_third = 1 / 3
_two_thirds = 2 / 3
_one_sixth = 1 / 6


def _runge_kutta_step(func, y0, f0, t0, dt, t1, tableau):
  t0 = t0.to(y0.dtype)
  dt = dt.to(y0.dtype)
  t1 = t1.to(y0.dtype)
  k = torch.empty(*f0.shape, len(tableau.alpha) + 1, dtype=y0.dtype, device=y0.device)
  k[..., 0] = f0
  for i, (alpha_i, beta_i) in enumerate(zip(tableau.alpha, tableau.beta)):
    ti = t1 if alpha_i == 1. else t0 + alpha_i * dt
    yi = y0 + k[..., :i + 1].matmul(beta_i * dt).view_as(f0)
    f = func(ti, yi)
    k[..., i + 1] = f
  if not (tableau.c_sol[-1] == 0 and (tableau.c_sol[:-1] == tableau.beta[-1]).all()):
    yi = y0 + k.matmul(dt * tableau.c_sol).view_as(f0)
  y1 = yi
  f1 = k[..., -1]
  y1_error = k.matmul(dt * tableau.c_error)
  return y1, f1, y1_error, k


def rk4_step_func(func, t0, dt, t1, y0, f0=None, perturb=False):
  k1 = f0 if f0 is not None else func(t0, y0)
  half_dt = dt * 0.5
  k2 = func(t0 + half_dt, y0 + half_dt * k1)
  k3 = func(t0 + half_dt, y0 + half_dt * k2)
  k4 = func(t1, y0 + dt * k3)
  return (k1 + 2 * (k2 + k3) + k4) * dt * _one_sixth


def rk4_alt_step_func(func, t0, dt, t1_or_y0, y0=None, f0=None, perturb=False):
  """Smaller error with slightly more compute: the 3/8 rule.  Accepts both the 0.2.1 (func, t0, dt, y0) and the
  >= 0.2.2 (func, t0, dt, t1, y0) call forms used at src/early_stop_solver.py:150-155."""
  if y0 is None:
    y0 = t1_or_y0
    t1 = t0 + dt
  else:
    t1 = t1_or_y0
  k1 = f0 if f0 is not None else func(t0, y0)
  k2 = func(t0 + dt * _one_third, y0 + dt * k1 * _one_third)
  k3 = func(t0 + dt * _two_thirds, y0 + dt * (k2 - k1 * _one_third))
  k4 = func(t1, y0 + dt * (k1 - k2 + k3))
  return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


class RKAdaptiveStepsizeODESolver(AdaptiveStepsizeODESolver):
  order: int
  tableau: _ButcherTableau
  mid: torch.Tensor

  def __init__(self, func, y0, rtol, atol, first_step=None, safety=0.9, ifactor=10.0, dfactor=0.2,
               max_num_steps=2 ** 31 - 1, dtype=torch.float64, norm=_rms_norm, **kwargs):
    kwargs.pop('step_size', None)
    kwargs.pop('max_iters', None)
    super().__init__(dtype=dtype, y0=y0, norm=norm)
    dtype = torch.promote_types(dtype, y0.dtype)
    device = y0.device
    self.func = func
    self.rtol = torch.as_tensor(rtol, dtype=dtype, device=device)
    self.atol = torch.as_tensor(atol, dtype=dtype, device=device)
    self.first_step = None if first_step is None else torch.as_tensor(first_step, dtype=dtype, device=device)
    self.safety = torch.as_tensor(safety, dtype=dtype, device=device)
    self.ifactor = torch.as_tensor(ifactor, dtype=dtype, device=device)
    self.dfactor = torch.as_tensor(dfactor, dtype=dtype, device=device)
    self.max_num_steps = torch.as_tensor(max_num_steps, dtype=torch.int32, device=device)
    self.dtype = dtype
    self.tableau = _ButcherTableau(alpha=self.tableau.alpha.to(device=device, dtype=y0.dtype),
                                   beta=[b.to(device=device, dtype=y0.dtype) for b in self.tableau.beta],
                                   c_sol=self.tableau.c_sol.to(device=device, dtype=y0.dtype),
                                   c_error=self.tableau.c_error.to(device=device, dtype=y0.dtype))
    self.mid = self.mid.to(device=device, dtype=y0.dtype)

  def _before_integrate(self, t):
    f0 = self.func(t[0], self.y0)
    if self.first_step is None:
      first_step = _select_initial_step(self.func, t[0], self.y0, self.order - 1, self.rtol, self.atol,
                                        self.norm, f0=f0)
    else:
      first_step = self.first_step
    self.rk_state = _RungeKuttaState(self.y0, f0, t[0], t[0], first_step, [self.y0] * 5)

  def _advance(self, next_t):
    n_steps = 0
    while next_t > self.rk_state.t1:
      assert n_steps < self.max_num_steps, 'max_num_steps exceeded ({}>={})'.format(n_steps, self.max_num_steps)
      self.rk_state = self._adaptive_step(self.rk_state)
      n_steps += 1
    return _interp_evaluate(self.rk_state.interp_coeff, self.rk_state.t0, self.rk_state.t1, next_t)

  def _adaptive_step(self, rk_state):
    y0, f0, _, t0, dt, interp_coeff = rk_state
    t1 = t0 + dt
    assert t0 + dt > t0, 'underflow in dt {}'.format(dt.item())
    assert torch.isfinite(y0).all(), 'non-finite values in state `y`: {}'.format(y0)
    y1, f1, y1_error, k = _runge_kutta_step(self.func, y0, f0, t0, dt, t1, tableau=self.tableau)
    error_ratio = _compute_error_ratio(y1_error, self.rtol, self.atol, y0, y1, self.norm)
    accept_step = error_ratio <= 1
    if accept_step:
      t_next = t1
      y_next = y1
      interp_coeff = self._interp_fit(y0, y_next, k, dt)
      f_next = f1
    else:
      t_next = t0
      y_next = y0
      f_next = f0
    dt_next = _optimal_step_size(dt, error_ratio, self.safety, self.ifactor, self.dfactor, self.order)
    return _RungeKuttaState(y_next, f_next, t0, t_next, dt_next, interp_coeff)

  def _interp_fit(self, y0, y1, k, dt):
    dt = dt.type_as(y0)
    y_mid = y0 + k.matmul(dt * self.mid).view_as(y0)
    f0 = k[..., 0]
    f1 = k[..., -1]
    return _interp_fit(y0, y1, y_mid, f0, f1, dt)
