from .solvers import FixedGridODESolver
from .rk_common import rk4_alt_step_func


class Euler(FixedGridODESolver):
  order = 1

  def _step_func(self, func, t0, dt, t1, y0):
    return dt * func(t0, y0)


class Midpoint(FixedGridODESolver):
  order = 2

  def _step_func(self, func, t0, dt, t1, y0):
    half_dt = 0.5 * dt
    y_mid = y0 + func(t0, y0) * half_dt
    return dt * func(t0 + half_dt, y_mid)


class RK4(FixedGridODESolver):
  order = 4

  def _step_func(self, func, t0, dt, t1, y0):
    return rk4_alt_step_func(func, t0, dt, t1, y0)
