import warnings
import torch


def _rms_norm(tensor):
  return tensor.abs().pow(2).mean().sqrt()


class _PerturbFunc(torch.nn.Module):
  def __init__(self, base_func):
    super().__init__()
    self.base_func = base_func

  def forward(self, t, y, *, perturb=None):
    # the reference's RHS never depends on t, so the nextafter perturbation of t is dropped
    return self.base_func(t.to(y.dtype) if torch.is_tensor(t) else t, y)


def _select_initial_step(func, t0, y0, order, rtol, atol, norm, f0=None):
  dtype, device, t_dtype = y0.dtype, y0.device, t0.dtype
  t0 = t0.to(dtype)
  if f0 is None:
    f0 = func(t0, y0)
  scale = atol + torch.abs(y0) * rtol
  d0 = norm(y0 / scale).abs()
  d1 = norm(f0 / scale).abs()
  if d0 < 1e-5 or d1 < 1e-5:
    h0 = torch.tensor(1e-6, dtype=dtype, device=device)
  else:
    h0 = 0.01 * d0 / d1
  h0 = h0.abs()
  y1 = y0 + h0 * f0
  f1 = func(t0 + h0, y1)
  d2 = torch.abs(norm((f1 - f0) / scale) / h0)
  if d1 <= 1e-15 and d2 <= 1e-15:
    h1 = torch.max(torch.tensor(1e-6, dtype=dtype, device=device), h0 * 1e-3)
  else:
    h1 = (0.01 / max(d1, d2)) ** (1. / float(order + 1))
  h1 = h1.abs()
  return torch.min(100 * h0, h1).to(t_dtype)


def _compute_error_ratio(error_estimate, rtol, atol, y0, y1, norm):
  error_tol = atol + rtol * torch.max(y0.abs(), y1.abs())
  return norm(error_estimate / error_tol).abs()


@torch.no_grad()
def _optimal_step_size(last_step, error_ratio, safety, ifactor, dfactor, order):
  if error_ratio == 0:
    return last_step * ifactor
  if error_ratio < 1:
    dfactor = torch.ones((), dtype=last_step.dtype, device=last_step.device)
  error_ratio = error_ratio.type_as(last_step)
  exponent = torch.tensor(order, dtype=last_step.dtype, device=last_step.device).reciprocal()
  factor = torch.min(ifactor, torch.max(safety / error_ratio ** exponent, dfactor))
  return last_step * factor


def _check_inputs(func, y0, t, rtol, atol, method, options, event_fn, SOLVERS):
  shapes = None
  assert torch.is_tensor(y0), 'tuple states are outside the hot path'
  if options is None:
    options = {}
  else:
    options = options.copy()
  if method is None:
    method = 'dopri5'
  if method not in SOLVERS:
    raise ValueError('Invalid method "{}". Must be one of {}'.format(method, sorted(SOLVERS.keys())))
  if not torch.is_tensor(t):
    t = torch.tensor(t, dtype=torch.float64)
  assert (t[1:] > t[:-1]).all(), 't must be strictly increasing'
  if t.device != y0.device:
    warnings.warn("t is not on the same device as y0. Coercing to y0.device.")
    t = t.to(y0.device)
  func = _PerturbFunc(func)
  return shapes, func, y0, t, rtol, atol, method, options, event_fn, False


def _flat_to_shape(tensor, length, shapes):
  raise NotImplementedError
