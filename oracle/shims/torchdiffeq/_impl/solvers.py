import abc
import torch
from .misc import _rms_norm


class AdaptiveStepsizeODESolver(metaclass=abc.ABCMeta):
  def __init__(self, dtype, y0, norm, **unused_kwargs):
    self.y0 = y0
    self.dtype = dtype
    self.norm = norm

  def _before_integrate(self, t):
    pass

  @abc.abstractmethod
  def _advance(self, next_t):
    raise NotImplementedError

  def integrate(self, t):
    solution = torch.empty(len(t), *self.y0.shape, dtype=self.y0.dtype, device=self.y0.device)
    solution[0] = self.y0
    t = t.to(self.dtype)
    self._before_integrate(t)
    for i in range(1, len(t)):
      solution[i] = self._advance(t[i])
    return solution


class FixedGridODESolver(metaclass=abc.ABCMeta):
  order: int

  def __init__(self, func, y0, step_size=None, grid_constructor=None, interp="linear", perturb=False,
               **unused_kwargs):
    unused_kwargs.pop('rtol', None)
    unused_kwargs.pop('atol', None)
    unused_kwargs.pop('norm', None)
    self.func = func
    self.y0 = y0
    self.dtype = y0.dtype
    self.device = y0.device
    self.step_size = step_size
    self.interp = interp
    self.perturb = perturb
    if step_size is None:
      self.grid_constructor = (lambda f, y0, t: t) if grid_constructor is None else grid_constructor
    else:
      assert grid_constructor is None, "step_size and grid_constructor are mutually exclusive arguments."
      self.grid_constructor = self._grid_constructor_from_step_size(step_size)

  @staticmethod
  def _grid_constructor_from_step_size(step_size):
    def _grid_constructor(func, y0, t):
      start_time = t[0]
      end_time = t[-1]
      niters = torch.ceil((end_time - start_time) / step_size + 1).item()
      t_infer = torch.arange(0, niters, dtype=t.dtype, device=t.device) * step_size + start_time
      t_infer[-1] = t[-1]
      return t_infer
    return _grid_constructor

  @abc.abstractmethod
  def _step_func(self, func, t0, dt, t1, y0):
    pass

  def integrate(self, t):
    time_grid = self.grid_constructor(self.func, self.y0, t)
    assert time_grid[0] == t[0] and time_grid[-1] == t[-1]
    solution = torch.empty(len(t), *self.y0.shape, dtype=self.y0.dtype, device=self.y0.device)
    solution[0] = self.y0
    j = 1
    y0 = self.y0
    for t0, t1 in zip(time_grid[:-1], time_grid[1:]):
      dt = t1 - t0
      dy = self._step_func(self.func, t0, dt, t1, y0)
      y1 = y0 + dy
      while j < len(t) and t1 >= t[j]:
        solution[j] = self._linear_interp(t0, t1, y0, y1, t[j])
        j += 1
      y0 = y1
    return solution

  def _linear_interp(self, t0, t1, y0, y1, t):
    if t == t0:
      return y0
    if t == t1:
      return y1
    slope = (t - t0) / (t1 - t0)
    return y0 + slope * (y1 - y0)
