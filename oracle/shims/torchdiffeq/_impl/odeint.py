import warnings
from .dopri5 import Dopri5Solver
from .fixed_grid import Euler, Midpoint, RK4
from .misc import _check_inputs

SOLVERS = {'dopri5': Dopri5Solver, 'euler': Euler, 'midpoint': Midpoint, 'rk4': RK4}


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
  shapes, func, y0, t, rtol, atol, method, options, event_fn, _ = _check_inputs(func, y0, t, rtol, atol, method,
                                                                                options, event_fn, SOLVERS)
  solver = SOLVERS[method](func=func, y0=y0, rtol=rtol, atol=atol, **options)
  return solver.integrate(t)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None,
                   adjoint_rtol=None, adjoint_atol=None, adjoint_method=None, adjoint_options=None,
                   adjoint_params=None):
  warnings.warn('shim: odeint_adjoint forward == odeint forward; the adjoint backward is not restated here')
  return odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
