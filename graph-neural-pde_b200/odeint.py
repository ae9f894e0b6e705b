"""torchdiffeq-compatible integrator front end driving the CUDA RHS engine.

`odeint(func, y0, t, rtol=, atol=, method=, options=)` has the call signature the reference's blocks use
(reference src/block_constant.py:46-62) and returns `[len(t), N, d]`.  `func` must be an ODEFunc of this package:
the solver does not call `func.forward` per stage -- it asks the function for its device description (attention spec
or CSR weights + residual parameters) and launches the stages natively:

  * euler / midpoint / rk4 (torchdiffeq's rk4 = 3/8 rule): the whole grid in ONE C call (grand_integrate_fixed),
    replayed from a CUDA Graph; the Runge-Kutta combine is the epilogue of each RHS launch.
  * dopri5: Dormand-Prince 5(4) with torchdiffeq's controller (rms error norm, safety 0.9, factors [0.2, 10], Hairer
    initial step), 6 RHS launches per attempted step with fused stage combines, one small error-norm kernel and one
    8-byte device->host read per attempt for the accept / step-size decision (the reference syncs there as well).

NFE accounting (`func.nfe`, MaxNFEException) matches the reference evaluation for evaluation.
"""
import math

import torch

from . import _native as N
from . import autograd_rhs as AG
from .engine import DP_BETA, DP_C_ERROR, DP_C_MID, RhsEngine, _f32, fixed_grid_dts

FIXED_METHODS = ('euler', 'midpoint', 'rk4')
EVALS_PER_STEP = {'euler': 1, 'midpoint': 2, 'rk4': 4}


def _require_native(func):
  if not hasattr(func, 'native_description'):
    raise TypeError('grand_b200.odeint integrates the ODEFunc modules of this package (got %r); arbitrary Python '
                    'callables would need the per-stage Python loop this engine exists to remove' % type(func))


class NativeProblem:
  """Everything the kernels need for `func` at state shape of `y0`, resolved once per integration."""

  def __init__(self, func, y0):
    _require_native(func)
    N.require_cuda(y0, 'y0')
    self.func = func
    self.engine = func.engine(y0)
    self.spec, self.w_csr = func.native_description(self.engine)
    self.x0 = func.stage_x0(self.engine)
    self.n = y0.shape[0]

  def epilogue(self, **outputs):
    return self.func.residual_epilogue(self.engine, self.x0, **outputs)

  def rhs(self, x, **outputs):
    ep = self.epilogue(**outputs)
    if self.spec is not None:
      self.engine.rhs_attention(self.spec, x, ep)
    else:
      self.engine.rhs_laplacian(self.w_csr, x, ep)


def _planes(buf, n, count):
  return [buf[i * n:(i + 1) * n] for i in range(count)]


def integrate_fixed(func, y0, t, method, step_size, on_step=None, use_graph=True):
  """FixedGridODESolver.integrate (in-repo witness reference src/early_stop_solver.py:163-184)."""
  prob = NativeProblem(func, y0)
  eng = prob.engine
  dts, grid = fixed_grid_dts(float(t[0]), float(t[-1]), step_size)
  assert grid[0] == float(t[0]) and grid[-1] == float(t[-1])
  y = eng.buffer('y')
  y.copy_(y0)
  per_step = EVALS_PER_STEP[method]
  out = torch.empty((len(t),) + tuple(y0.shape), dtype=y0.dtype, device=y0.device)
  out[0] = y0
  if len(t) == 2 and on_step is None:
    func._bump_nfe(per_step * len(dts))
    eng.integrate_fixed(prob.spec, prob.w_csr, prob.epilogue(), method, dts, y, use_graph=use_graph)
    out[1] = y
    return out
  # general case: requested interior times and / or a per-step hook -> step by step, linear interpolation
  tt = [float(v) for v in t]
  prev = eng.buffer('y_prev')
  j = 1
  for s, h in enumerate(dts):
    prev.copy_(y)
    func._bump_nfe(per_step)
    eng.integrate_fixed(prob.spec, prob.w_csr, prob.epilogue(), method, [h], y, use_graph=False)
    ta, tb = float(grid[s]), float(grid[s + 1])
    if on_step is not None:
      on_step(ta, tb, prev, y)
    while j < len(tt) and tb >= tt[j]:
      if tb == tt[j]:
        out[j] = y
      else:
        slope = _f32((tt[j] - ta) / (tb - ta))
        eng.lincomb(out[j], [(prev, 1.0 - slope), (y, slope)])
      j += 1
  return out


class Dopri5:
  """RKAdaptiveStepsizeODESolver for the DP5(4) tableau on device buffers (torchdiffeq 0.2.x semantics, witnessed by
  reference src/early_stop_solver.py:30-86).  Time-like scalars are python floats (= float64, as in torchdiffeq)."""

  order = 5
  safety, ifactor, dfactor = 0.9, 10.0, 0.2

  def __init__(self, func, y0, rtol, atol):
    self.prob = NativeProblem(func, y0)
    self.func = func
    self.rtol, self.atol = float(rtol), float(atol)
    eng = self.prob.engine
    n = self.prob.n
    planes = _planes(eng.buffer('dopri', rows=11 * n), n, 11)
    self.k = planes[0:7]
    self.ycur, self.ya, self.yb, self.ymid = planes[7], planes[8], planes[9], planes[10]
    self.ycur.copy_(y0)
    self.acc = torch.zeros(1, dtype=torch.float64, device=y0.device)
    self.numel = y0.numel()
    self.attempts = 0
    self.accepted = 0

  def _rms(self, terms, y0, y1, atol, rtol):
    self.acc.zero_()
    self.prob.engine.error_sumsq(self.acc, terms, y0, y1, atol, rtol)
    return math.sqrt(float(self.acc.item()) / self.numel)

  def before_integrate(self, t0):
    p, k = self.prob, self.k
    self.func._bump_nfe()
    p.rhs(self.ycur, f_out=k[0])
    # _select_initial_step (Hairer), order - 1 = 4
    d0 = self._rms([(self.ycur, 1.0)], self.ycur, self.ycur, self.atol, self.rtol)
    d1 = self._rms([(k[0], 1.0)], self.ycur, self.ycur, self.atol, self.rtol)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = _f32(abs(h0))
    p.engine.lincomb(self.ya, [(self.ycur, 1.0), (k[0], h0)])
    self.func._bump_nfe()
    p.rhs(self.ya, f_out=k[1])
    d2 = self._rms([(k[1], 1.0), (k[0], -1.0)], self.ycur, self.ycur, self.atol, self.rtol) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
      h1 = max(1e-6, h0 * 1e-3)
    else:
      h1 = _f32((0.01 / max(d1, d2)) ** (1. / 5.))
    self.dt = float(min(100 * h0, abs(h1)))
    self.t0 = self.t1 = float(t0)
    self.have_interp = False

  def adaptive_step(self, force_accept=None):
    """One attempted step; returns True when accepted.  On accept the state (ycur, k[0]) moves to t1.  `force_accept`
    overrides the controller's decision (replay of another implementation's step sequence in the parity tests)."""
    p, k, func = self.prob, self.k, self.func
    t0, dt = self.t1, self.dt
    assert t0 + dt > t0, 'underflow in dt {}'.format(dt)
    h = _f32(dt)
    p.engine.lincomb(self.ya, [(self.ycur, 1.0), (k[0], _f32(_f32(DP_BETA[0][0]) * h))])
    inp, out = self.ya, self.yb
    for i in range(6):
      func._bump_nfe()
      if i < 5:
        beta = DP_BETA[i + 1]
        terms = [(self.ycur, 1.0)] + [(k[j], _f32(_f32(beta[j]) * h)) for j in range(i + 1) if beta[j] != 0]
        p.rhs(inp, f_out=k[i + 1], y_out=out, coef_f=_f32(_f32(beta[i + 1]) * h), terms=terms)
        inp, out = out, inp
      else:
        p.rhs(inp, f_out=k[6])
    y1 = inp   # the 6th stage input is the 5th-order solution (c_sol == last beta row, FSAL)
    err_terms = [(k[j], _f32(h * _f32(DP_C_ERROR[j]))) for j in range(7) if DP_C_ERROR[j] != 0]
    ratio = self._rms(err_terms, self.ycur, y1, self.atol, self.rtol)
    accept = ratio <= 1 if force_accept is None else bool(force_accept)
    self.last_ratio = ratio
    self.attempts += 1
    if accept:
      self.accepted += 1
      # dense output data of this step: y0 (old ycur), y1, y_mid, f0 = k0, f1 = k6, dt
      mid_terms = [(self.ycur, 1.0)] + [(k[j], _f32(h * _f32(DP_C_MID[j]))) for j in range(7) if DP_C_MID[j] != 0]
      p.engine.lincomb(self.ymid, mid_terms)
      self.interp = dict(y0=self.ycur, y1=y1, f0=k[0], f1=k[6], dt=h, t0=t0, t1=t0 + dt)
      self.have_interp = True
      # advance: ycur <- y1 buffer, k0 <- k6 (buffer swaps; the old buffers stay valid for the interpolation
      # until the next attempt overwrites them)
      free = self.ycur
      self.ycur = y1
      if y1 is self.ya:
        self.ya = free
      else:
        self.yb = free
      k[0], k[6] = k[6], k[0]
      self.interp['f0'] = k[6]
      self.interp['f1'] = k[0]
      self.t0, self.t1 = t0, t0 + dt
    else:
      self.t0 = self.t1 = t0
    self.dt = self._optimal_step_size(dt, ratio)
    return accept

  def _optimal_step_size(self, last_step, error_ratio):
    if error_ratio == 0:
      return last_step * self.ifactor
    dfactor = 1.0 if error_ratio < 1 else self.dfactor
    factor = min(self.ifactor, max(self.safety / error_ratio ** (1.0 / self.order), dfactor))
    return last_step * factor

  def interpolate(self, t, out):
    """_interp_evaluate(_interp_fit(y0, y1, y_mid, f0, f1, dt), t0, t1, t) as ONE linear combination."""
    I = self.interp
    if t == I['t1']:
      out.copy_(I['y1'])
      return out
    x = _f32((t - I['t0']) / (I['t1'] - I['t0']))
    dt = I['dt']
    x2, x3, x4 = x * x, x * x * x, x * x * x * x
    # coefficient rows [e, d, c, b, a] over the basis (y0, y1, y_mid, f0, f1)
    w_y0 = 1.0 + x2 * (-11.0) + x3 * 18.0 + x4 * (-8.0)
    w_y1 = x2 * (-5.0) + x3 * 14.0 + x4 * (-8.0)
    w_ym = x2 * 16.0 + x3 * (-32.0) + x4 * 16.0
    w_f0 = dt * (x + x2 * (-4.0) + x3 * 5.0 + x4 * (-2.0))
    w_f1 = dt * (x2 * 1.0 + x3 * (-3.0) + x4 * 2.0)
    self.prob.engine.lincomb(out, [(I['y0'], _f32(w_y0)), (I['y1'], _f32(w_y1)), (self.ymid, _f32(w_ym)),
                                   (I['f0'], _f32(w_f0)), (I['f1'], _f32(w_f1))])
    return out


class Dopri5Device:
  """dopri5 with the controller on the device (grand_dopri5_create / launch, csrc/dopri5_device.cu): one CUDA Graph
  with a conditional WHILE node per (function, shape, tolerances, time span); the host launches it and reads the
  24-byte counters once per integration (NFE accounting).  Cached on the engine."""

  def __init__(self, func, y0, t0, t1, rtol, atol, max_attempts):
    import ctypes as C
    self.prob = NativeProblem(func, y0)
    eng = self.prob.engine
    n = self.prob.n
    self.planes = eng.buffer('dopri_dev', rows=13 * n)
    self.y0 = eng.buffer('dopri_dev_y0', rows=n)
    self.out = eng.buffer('dopri_dev_out', rows=n)
    self.state = torch.zeros(C.sizeof(N.GrandDopri5State), dtype=torch.uint8, device=y0.device)
    self.ep = self.prob.epilogue()
    plan = C.c_void_p()
    ws = eng.workspace(self.prob.spec)
    N.check(N.lib().grand_dopri5_create(
      C.byref(eng.plan.struct()), C.byref(self.prob.spec.struct) if self.prob.spec is not None else None,
      N.ptr(self.prob.w_csr), eng.d, C.byref(ws), C.byref(self.ep), self.y0.data_ptr(), self.out.data_ptr(),
      self.planes.data_ptr(), self.state.data_ptr(), float(t0), float(t1), float(rtol), float(atol), int(max_attempts),
      C.byref(plan)))
    self.plan = plan
    self.key = None

  def __del__(self):
    try:
      if getattr(self, 'plan', None):
        N.lib().grand_dopri5_destroy(self.plan)
    except Exception:
      pass

  def run(self, y0):
    self.y0.copy_(y0)
    N.check(N.lib().grand_dopri5_launch(self.plan, N.current_stream(y0.device)))
    return self.out

  def counters(self):
    st = N.GrandDopri5State.from_buffer_copy(self.state.cpu().numpy().tobytes())
    return st


def integrate_dopri5_device(func, y0, t, rtol, atol):
  """y(t[0]) -> y(t[1]) with the device-side controller; returns (out [2, N, d], counters)."""
  prob_engine = func.engine(y0)
  t0, t1 = float(t[0]), float(t[-1])
  budget = int(func.opt.get('max_nfe', 1000)) - int(func.nfe) - 2
  max_attempts = max(1, budget // 6 + 1)
  ptrs = tuple(p.data_ptr() for p in func.parameters())
  key = ('dopri5_dev', tuple(y0.shape), t0, t1, float(rtol), float(atol), max_attempts, ptrs, id(prob_engine.plan),
         None if getattr(func, 'x0', None) is None else func.x0.data_ptr())
  cache = prob_engine._state.setdefault('_dopri5_dev', {})
  solver = cache.get(key)
  if solver is None:
    cache.clear()                     # one plan per engine: buffers are shared
    solver = Dopri5Device(func, y0, t0, t1, rtol, atol, max_attempts)
    cache[key] = solver
  z = solver.run(y0)
  st = solver.counters()              # the one device -> host read of the integration
  if st.status != 0:
    raise AssertionError('underflow in dt {}'.format(st.dt))
  func._bump_nfe(int(st.nfe))
  out = torch.empty((2,) + tuple(y0.shape), dtype=y0.dtype, device=y0.device)
  out[0] = y0
  out[1] = z
  return out, st


def integrate_dopri5(func, y0, t, rtol, atol, on_step=None, max_steps=None):
  solver = Dopri5(func, y0, rtol, atol)
  tt = [float(v) for v in t]
  out = torch.empty((len(tt),) + tuple(y0.shape), dtype=y0.dtype, device=y0.device)
  out[0] = y0
  solver.before_integrate(tt[0])
  for i in range(1, len(tt)):
    n_steps = 0
    while tt[i] > solver.t1 and (max_steps is None or n_steps < max_steps):
      accepted = solver.adaptive_step()
      n_steps += 1
      if on_step is not None:
        on_step(solver, accepted)
    if max_steps is not None and n_steps >= max_steps:
      # EarlyStopDopri5.advance (reference src/early_stop_solver.py:76-86): `if n_steps < max_test_steps: interpolate else:
      # the state at t1` -- also when the last allowed step happened to reach the target time
      out[i] = solver.ycur
    elif not solver.have_interp:
      out[i] = solver.ycur
    else:
      solver.interpolate(tt[i], out[i])
  return out, solver


def _check_call(y0, method, options, event_fn):
  if event_fn is not None:
    raise NotImplementedError('event handling is not used by the reference hot path')
  if not torch.is_tensor(y0):
    raise NotImplementedError('tuple states (RNODE regulariser states) are outside the hot path')
  options = dict(options or {})
  method = method or 'dopri5'
  if method in FIXED_METHODS:
    if 'step_size' not in options or options['step_size'] is None:
      raise NotImplementedError('fixed-grid solvers need options["step_size"] (every reference call passes it)')
  elif method != 'dopri5':
    raise ValueError('Invalid method "{}". Must be one of {}'.format(method, sorted(FIXED_METHODS + ('dopri5',))))
  return method, options


def _native_solution(func, y0, t, method, options, rtol, atol):
  if method in FIXED_METHODS:
    return integrate_fixed(func, y0, t, method, options['step_size'])
  import os
  if len(t) == 2 and float(t[-1]) > float(t[0]) and os.environ.get('GRAND_DOPRI5', 'device') == 'device':
    out, _ = integrate_dopri5_device(func, y0, t, rtol, atol)     # controller on the device: one graph launch
    return out
  out, _ = integrate_dopri5(func, y0, t, rtol, atol)
  return out


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
  """torchdiffeq.odeint call-compatible.  No gradients required: the fused engine path.  Gradients required
  (training with adjoint=False): the same solvers with every RHS evaluation as one autograd node."""
  method, options = _check_call(y0, method, options, event_fn)
  _require_native(func)
  per_call = getattr(func, 'per_call_only', None)
  if AG.needs_grad(func, y0) or (per_call is not None and per_call()):
    # every RHS evaluation is one call of func.forward (an autograd node when gradients are required; several launches
    # for the functions that have no single fused RHS launch, e.g. GAT with mix_features)
    if method in FIXED_METHODS:
      _, grid = fixed_grid_dts(float(t[0]), float(t[-1]), options['step_size'])
      return AG.odeint_fixed_autograd(func, y0, t, method, grid)
    return AG.odeint_dopri5_autograd(func, y0, t, rtol, atol)
  return _native_solution(func, y0, t, method, options, rtol, atol)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None,
                   adjoint_rtol=None, adjoint_atol=None, adjoint_method=None, adjoint_options=None,
                   adjoint_params=None):
  """torchdiffeq.odeint_adjoint call-compatible: forward on the fused engine without an autograd graph, backward by
  integrating the adjoint system (reference call site src/block_constant.py:45-55)."""
  method, options = _check_call(y0, method, options, event_fn)
  _require_native(func)
  if not AG.needs_grad(func, y0):
    return _native_solution(func, y0, t, method, options, rtol, atol)
  adjoint_method = adjoint_method or method
  adjoint_options = dict(adjoint_options if adjoint_options is not None else options)
  params = tuple(p for p in (adjoint_params if adjoint_params is not None else func.parameters()) if p.requires_grad)

  def fwd(y):
    return _native_solution(func, y, t, method, options, rtol, atol)

  return AG.OdeintAdjointFn.apply(y0, func, t, fwd, adjoint_method, adjoint_options,
                                  adjoint_rtol if adjoint_rtol is not None else rtol,
                                  adjoint_atol if adjoint_atol is not None else atol, *params)
