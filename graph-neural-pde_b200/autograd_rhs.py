"""Training support: the RHS as a differentiable op (forward and vector-Jacobian product both on the CUDA engine) and
the solver loops used when gradients are required.

The reference differentiates through ~30 small PyTorch / torch_scatter ops per RHS evaluation, either directly
(`adjoint=False`: autograd through every stored solver stage, reference src/run_GNN.py:92) or inside
torchdiffeq's adjoint (`adjoint=True`, reference src/block_constant.py:45-55).  Here one RHS evaluation is
ONE autograd node:

    forward   grand_rhs_attention / grand_rhs_laplacian                      (the inference kernels)
    backward  grand_rhs_attention_backward  -> abar, dQ, dK                    (edge phase, csrc/rhs_backward.cu)
              grand_rhs_laplacian on the FLIPPED graph with abar             -> alpha (Abar^T g - g)
              dense tails dX += dQ Wq + dK Wk, dWq = dQ^T X, ... (cuBLAS GEMMs via torch.matmul)

The stage algebra of the solvers (a handful of axpy's per step) runs as torch ops in this mode so that autograd can see
it; the no-grad path (odeint.py) keeps it fused in the kernel epilogues.  Built for the fused attention configuration
(scaled_dot, softmax, attention_norm_idx 0) and for the Laplacian function with fixed weights; everything else raises.
"""
import torch

from . import _native as N
from .engine import DP_BETA, DP_C_ERROR, DP_C_MID, RhsEngine
from .graph_plan import GraphPlan


def _transposed_engine(func, x):
  """Engine over the flipped edge list (row <-> col), cached next to the forward engine."""
  eng = func.engine(x)
  cached = getattr(func, '_engine_T', None)
  if cached is None or cached[0] is not eng:
    plan_t = GraphPlan(func.edge_index.flip(0), num_nodes=x.shape[0], device=x.device)
    cached = (eng, RhsEngine(plan_t, eng.d))
    func._engine_T = cached
  return cached[1]


def _same_engine(ctx, func, x):
  """The engine (graph plan) captured at forward time; a graph that was re-assigned between forward and backward
  (hard-attention re-selection, an outside `edge_index` assignment) is an error, as a version-counter mismatch is for
  tensors in plain autograd."""
  eng = func.engine(x)
  if eng is not ctx.engine:
    raise RuntimeError('grand_b200 backward: the graph of this ODE function changed between forward and backward')
  return eng


def _sigmoid_grad(alpha_train, no_alpha_sigmoid):
  if no_alpha_sigmoid:
    return torch.ones_like(alpha_train)
  s = torch.sigmoid(alpha_train)
  return s * (1 - s)


class _ConstOne:
  """alpha_train stand-in (no_alpha_sigmoid = True, value 1) to evaluate r = A x - x with the forward kernels."""
  _cache = {}

  @classmethod
  def get(cls, device):
    t = cls._cache.get(device)
    if t is None:
      t = torch.ones((), dtype=torch.float32, device=device)
      cls._cache[device] = t
    return t


class TransformerRhsFn(torch.autograd.Function):
  """f = ODEFuncTransformerAtt.forward(t, x) as one differentiable node."""

  @staticmethod
  def forward(ctx, x, wq, bq, wk, bk, alpha_train, beta_train, func):
    x = x.contiguous()
    eng = func.engine(x)
    spec = func.multihead_att_layer.spec(eng)
    if not spec.fused:
      raise NotImplementedError('grand_b200 backward is built for attention_type=scaled_dot, softmax, '
                                'attention_norm_idx=0 (the other configurations run forward-only)')
    if x.shape[1] % 4 != 0:
      raise NotImplementedError('grand_b200 backward needs a state width that is a multiple of 4 (got %d)' % x.shape[1])
    f = torch.empty_like(x)
    ep = func.residual_epilogue(eng, func.stage_x0(eng), f_out=f)
    eng.rhs_attention(spec, x, ep)
    ctx.func = func
    ctx.engine = eng                      # the graph plan and source term this evaluation was computed with
    ctx.x0 = func.x0 if func.opt['add_source'] else None
    ctx.save_for_backward(x, wq, bq, wk, bk, alpha_train, beta_train)
    return f

  @staticmethod
  def backward(ctx, g):
    x, wq, bq, wk, bk, alpha_train, beta_train = ctx.saved_tensors
    func = ctx.func
    opt = func.opt
    g = g.contiguous()
    eng = _same_engine(ctx, func, x)
    spec = func.multihead_att_layer.spec(eng)
    no_sig = bool(opt['no_alpha_sigmoid'])
    # 1. edge phase
    abar, dq, dk = eng.attention_backward(spec, x, g, alpha_train.detach(), no_sig)
    # 2. transposed aggregation + residual: alpha (Abar^T g - g)
    eng_t = _transposed_engine(func, x)
    gx = torch.empty_like(x)
    ep_t = RhsEngine.epilogue(alpha_train.detach(), None, None, no_sig, f_out=gx)
    eng_t.rhs_laplacian(eng_t.edge_values_to_csr(abar), g, ep_t)
    # 3. dense tails (plain GEMMs)
    gx = gx + dq @ wq + dk @ wk
    dwq, dbq = dq.t() @ x, dq.sum(0)
    dwk, dbk = dk.t() @ x, dk.sum(0)
    # 4. scalars: d alpha needs r = A x - x (forward kernel with alpha := 1, no source term)
    dalpha = dbeta = None
    if ctx.needs_input_grad[5]:
      r = torch.empty_like(x)
      ep_r = RhsEngine.epilogue(_ConstOne.get(x.device), None, None, True, f_out=r)
      eng.rhs_attention(spec, x, ep_r)
      dalpha = (g * r).sum() * _sigmoid_grad(alpha_train.detach(), no_sig)
    if ctx.needs_input_grad[6] and opt['add_source']:
      dbeta = (g * ctx.x0).sum()
    return gx, dwq, dbq, dwk, dbk, dalpha, dbeta, None


class GeneralRhsFn(torch.autograd.Function):
  """One RHS evaluation of a function WITHOUT a fused backward kernel as an autograd node: the transformer function in
  the configurations outside scaled_dot / softmax / norm idx 0 / width % 4 == 0 (cosine, pearson, exp kernel, squareplus,
  idx 1, reweighting), and the GAT function (incl. mix_features).  Forward: the function's native evaluation
  (`func.native_forward`), as without gradients.  Backward: autograd through the torch restatement of the same
  evaluation (`func.torch_rhs`, attention_autograd.py) -- edge-sized torch kernels per evaluation, the price of a
  configuration outside the fused path; the reference pays it for every configuration."""

  @staticmethod
  def forward(ctx, x, alpha_train, beta_train, func, *layer_params):
    x = x.contiguous()
    f = func.native_forward(x)
    ctx.func, ctx.engine = func, func.engine(x)
    ctx.x0 = func.x0 if func.opt['add_source'] else None
    ctx.n_layer = len(layer_params)
    ctx.save_for_backward(x)
    return f

  @staticmethod
  def backward(ctx, g):
    (x,) = ctx.saved_tensors
    func = ctx.func
    _same_engine(ctx, func, x)
    layer_params = list(func.multihead_att_layer.parameters())
    assert len(layer_params) == ctx.n_layer
    with torch.enable_grad():
      x_ = x.detach().requires_grad_(True)
      f_ = func.torch_rhs(x_, ctx.x0)
      wanted = [x_, func.alpha_train, func.beta_train] + layer_params
      live = [w for w in wanted if w.requires_grad]
      got = iter(torch.autograd.grad(f_, live, g.contiguous(), allow_unused=True))
    grads = [next(got) if w.requires_grad else None for w in wanted]
    return (grads[0], grads[1], grads[2], None) + tuple(grads[3:])


TransformerRhsGeneralFn = GeneralRhsFn


class LaplacianRhsFn(torch.autograd.Function):
  """f = LaplacianODEFunc.forward(t, x) with weights that do not depend on x (constant block)."""

  @staticmethod
  def forward(ctx, x, alpha_train, beta_train, func):
    x = x.contiguous()
    eng = func.engine(x)
    f = torch.empty_like(x)
    ep = func.residual_epilogue(eng, func.stage_x0(eng), f_out=f)
    eng.rhs_laplacian(func.weights_csr(eng), x, ep)
    ctx.func = func
    ctx.engine = eng
    ctx.x0 = func.x0 if func.opt['add_source'] else None
    ctx.edge_values = func.edge_values().detach()
    ctx.save_for_backward(x, alpha_train, beta_train)
    return f

  @staticmethod
  def backward(ctx, g):
    x, alpha_train, beta_train = ctx.saved_tensors
    func = ctx.func
    opt = func.opt
    g = g.contiguous()
    no_sig = bool(opt['no_alpha_sigmoid'])
    eng = _same_engine(ctx, func, x)
    eng_t = _transposed_engine(func, x)
    gx = torch.empty_like(x)
    ep_t = RhsEngine.epilogue(alpha_train.detach(), None, None, no_sig, f_out=gx)
    eng_t.rhs_laplacian(eng_t.edge_values_to_csr(ctx.edge_values), g, ep_t)
    dalpha = dbeta = None
    if ctx.needs_input_grad[1]:
      r = torch.empty_like(x)
      ep_r = RhsEngine.epilogue(_ConstOne.get(x.device), None, None, True, f_out=r)
      eng.rhs_laplacian(eng.edge_values_to_csr(ctx.edge_values), x, ep_r)
      dalpha = (g * r).sum() * _sigmoid_grad(alpha_train.detach(), no_sig)
    if ctx.needs_input_grad[2] and opt['add_source']:
      dbeta = (g * ctx.x0).sum()
    return gx, dalpha, dbeta, None


class FrozenAttentionFn(torch.autograd.Function):
  """AttODEblock.get_attention_weights as a differentiable node (reference src/block_transformer_attention.py:28-39):
  (x0, layer parameters...) -> (attention[E, H] in caller order, not differentiable; mean_h attention[E] in CSR order,
  differentiable).  The Laplacian function integrates with the second output (FrozenLaplacianRhsFn); its gradient sums
  over the RHS evaluations and comes back here ONCE per training step.  Forward: always the native kernel.  Backward:
  the fused configuration (scaled_dot, softmax, norm idx 0) runs grand_attention_weights_backward + the dense tails;
  every other scorer / normaliser differentiates the torch restatement of the layer at x0 (attention_autograd.py)."""

  @staticmethod
  def forward(ctx, x0, layer, eng, edge_index, *params):
    x0 = x0.contiguous()
    spec = layer.spec(eng)
    att, _, mean_csr = eng.attention_weights(spec, x0, want_attention=True, want_prods=False, want_mean_csr=True)
    ctx.layer, ctx.engine, ctx.fused, ctx.edge_index = layer, eng, bool(spec.fused), edge_index
    ctx.names = [n for n, _ in layer.named_parameters()]
    ctx.save_for_backward(x0, *params)
    ctx.mark_non_differentiable(att)
    return att, mean_csr

  @staticmethod
  def backward(ctx, _g_att, c_csr):
    x0, *params = ctx.saved_tensors
    eng, layer = ctx.engine, ctx.layer
    by_name = dict(zip(ctx.names, params))
    grads = {n: None for n in ctx.names}
    if ctx.fused and x0.shape[1] % 4 == 0:
      wq, wk = by_name['Q.weight'], by_name['K.weight']
      dq, dk = eng.attention_weights_backward(layer.spec(eng), x0, c_csr)
      gx = dq @ wq + dk @ wk
      grads.update({'Q.weight': dq.t() @ x0, 'Q.bias': dq.sum(0), 'K.weight': dk.t() @ x0, 'K.bias': dk.sum(0)})
    else:
      from .attention_autograd import mean_attention
      with torch.enable_grad():
        x0_ = x0.detach().requires_grad_(True)
        mean = mean_attention(x0_, ctx.edge_index, layer, eng.plan.n_rows)       # caller order, live parameters
        wanted = [(n, p) for n, p in layer.named_parameters() if p.requires_grad]
        # c is indexed by CSR position: bring it to the caller's edge order
        c_caller = torch.zeros_like(mean).index_copy_(0, eng.plan.eid_long, c_csr.contiguous())
        out = torch.autograd.grad(mean, [x0_] + [p for _, p in wanted], c_caller, allow_unused=True)
      gx = out[0]
      for (n, _), g in zip(wanted, out[1:]):
        grads[n] = g
    return (gx, None, None, None) + tuple(grads[n] for n in ctx.names)


class FrozenLaplacianRhsFn(torch.autograd.Function):
  """f = LaplacianODEFunc.forward(t, x) with differentiable per-edge weights `w_csr` (CSR order; the head mean of the
  frozen attention).  `w_csr_t` is the same weight vector in the CSR order of the flipped graph (for A^T g)."""

  @staticmethod
  def forward(ctx, x, w_csr, alpha_train, beta_train, func, w_csr_t):
    x = x.contiguous()
    eng = func.engine(x)
    f = torch.empty_like(x)
    ep = func.residual_epilogue(eng, func.stage_x0(eng), f_out=f)
    eng.rhs_laplacian(w_csr, x, ep)
    ctx.func, ctx.engine, ctx.w_csr_t = func, eng, w_csr_t
    ctx.x0 = func.x0 if func.opt['add_source'] else None
    ctx.save_for_backward(x, w_csr, alpha_train, beta_train)
    return f

  @staticmethod
  def backward(ctx, g):
    x, w_csr, alpha_train, beta_train = ctx.saved_tensors
    func = ctx.func
    opt = func.opt
    g = g.contiguous()
    no_sig = bool(opt['no_alpha_sigmoid'])
    eng = _same_engine(ctx, func, x)
    eng_t = _transposed_engine(func, x)
    gx = torch.empty_like(x)
    ep_t = RhsEngine.epilogue(alpha_train.detach(), None, None, no_sig, f_out=gx)
    eng_t.rhs_laplacian(ctx.w_csr_t, g, ep_t)
    gw = eng.edge_dot(x, g, alpha_train.detach(), no_sig) if ctx.needs_input_grad[1] else None
    dalpha = dbeta = None
    if ctx.needs_input_grad[2]:
      r = torch.empty_like(x)
      ep_r = RhsEngine.epilogue(_ConstOne.get(x.device), None, None, True, f_out=r)
      eng.rhs_laplacian(w_csr, x, ep_r)
      dalpha = (g * r).sum() * _sigmoid_grad(alpha_train.detach(), no_sig)
    if ctx.needs_input_grad[3] and opt['add_source']:
      dbeta = (g * ctx.x0).sum()
    return gx, gw, dalpha, dbeta, None, None


def needs_grad(func, *tensors):
  if not torch.is_grad_enabled():
    return False
  if any(t is not None and t.requires_grad for t in tensors):
    return True
  return any(p.requires_grad for p in func.parameters()) and func.training


# ---------------------------------------------------------------------------------------------------------------
# solver loops in torch ops (autograd-visible); algorithms as in odeint.py / torchdiffeq
# ---------------------------------------------------------------------------------------------------------------
def _rk4_38(func, t0, dt, t1, y):
  k1 = func(t0, y)
  k2 = func(t0 + dt / 3, y + dt * k1 * (1 / 3))
  k3 = func(t0 + dt * (2 / 3), y + dt * (k2 - k1 * (1 / 3)))
  k4 = func(t1, y + dt * (k1 - k2 + k3))
  return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


def _euler(func, t0, dt, t1, y):
  return dt * func(t0, y)


def _midpoint(func, t0, dt, t1, y):
  half = 0.5 * dt
  return dt * func(t0 + half, y + func(t0, y) * half)


FIXED_STEPS = {'rk4': _rk4_38, 'euler': _euler, 'midpoint': _midpoint}


def odeint_fixed_autograd(func, y0, t, method, grid):
  """Fixed-grid integration whose every RHS evaluation is an autograd node (`func(t, y)`)."""
  step = FIXED_STEPS[method]
  out = [y0]
  y = y0
  j = 1
  tt = [float(v) for v in t]
  for s in range(len(grid) - 1):
    ta, tb = float(grid[s]), float(grid[s + 1])
    y1 = y + step(func, ta, tb - ta, tb, y)
    while j < len(tt) and tb >= tt[j]:
      out.append(y1 if tb == tt[j] else y + ((tt[j] - ta) / (tb - ta)) * (y1 - y))
      j += 1
    y = y1
  return torch.stack(out, dim=0)


def _rms(t):
  return t.abs().pow(2).mean().sqrt()


def odeint_dopri5_autograd(func, y0, t, rtol, atol):
  """dopri5 with torchdiffeq's controller; step acceptance is decided on detached values."""
  tt = [float(v) for v in t]
  f0 = func(tt[0], y0)
  with torch.no_grad():
    scale = atol + y0.abs() * rtol
    d0, d1 = float(_rms(y0 / scale)), float(_rms(f0 / scale))
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = func(tt[0] + h0, y0 + h0 * f0)
    d2 = float(_rms((f1 - f0) / scale)) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** 0.2
    dt = min(100 * h0, h1)
  out = [y0]
  y, f, t0 = y0, f0, tt[0]
  interp = None
  for target in tt[1:]:
    while interp is None or target > interp[1]:
      ks = [f]
      yi = y
      for beta in DP_BETA:
        yi = y + dt * sum(b * k for b, k in zip(beta, ks) if b != 0)
        ks.append(func(t0, yi))
      err = dt * sum(c * k for c, k in zip(DP_C_ERROR, ks) if c != 0)
      with torch.no_grad():
        tol = atol + rtol * torch.max(y.abs(), yi.abs())
        ratio = float(_rms(err / tol))
      accept = ratio <= 1
      if accept:
        y_mid = y + dt * sum(c * k for c, k in zip(DP_C_MID, ks) if c != 0)
        interp = (t0, t0 + dt, y, yi, y_mid, ks[0], ks[-1], dt)
        y, f, t0 = yi, ks[-1], t0 + dt
      if ratio == 0:
        dt = dt * 10.0
      else:
        dt = dt * min(10.0, max(0.9 / ratio ** 0.2, 1.0 if ratio < 1 else 0.2))
    ta, tb, ya, yb, ym, fa, fb, h = interp
    if target == tb:
      out.append(yb)
    else:
      x = (target - ta) / (tb - ta)
      a = 2 * h * (fb - fa) - 8 * (yb + ya) + 16 * ym
      b = h * (5 * fa - 3 * fb) + 18 * ya + 14 * yb - 32 * ym
      c = h * (fb - 4 * fa) - 11 * ya - 5 * yb + 16 * ym
      out.append(ya + x * (h * fa) + x ** 2 * c + x ** 3 * b + x ** 4 * a)
  return torch.stack(out, dim=0)


# ---------------------------------------------------------------------------------------------------------------
# adjoint sensitivity (torchdiffeq odeint_adjoint semantics): forward without a graph on the fused engine, backward =
# integration of (y, a_y, a_theta) from T to 0 with `adjoint_method` / `adjoint_options`
# ---------------------------------------------------------------------------------------------------------------
class OdeintAdjointFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, y0, func, t, fwd, adjoint_method, adjoint_options, adjoint_rtol, adjoint_atol, *params):
    with torch.no_grad():
      sol = fwd(y0.detach())
    ctx.func, ctx.adjoint_method, ctx.adjoint_options = func, adjoint_method, dict(adjoint_options or {})
    ctx.adjoint_rtol, ctx.adjoint_atol = adjoint_rtol, adjoint_atol
    ctx.t = [float(v) for v in t]
    ctx.save_for_backward(sol, *params)
    return sol

  @staticmethod
  def backward(ctx, grad_sol):
    sol, *params = ctx.saved_tensors
    func = ctx.func
    t = ctx.t
    n_t = len(t)
    a = grad_sol[n_t - 1].clone()
    y = sol[n_t - 1]
    g_params = [torch.zeros_like(p) for p in params]

    def aug_dynamics(tt, state):
      yy, aa = state[0], state[1]
      with torch.enable_grad():
        yy_ = yy.detach().requires_grad_(True)
        ff = func(tt, yy_)
        grads = torch.autograd.grad(ff, [yy_] + list(params), -aa, allow_unused=True)
      gy = grads[0] if grads[0] is not None else torch.zeros_like(yy)
      gp = [gi if gi is not None else torch.zeros_like(p) for gi, p in zip(grads[1:], params)]
      return [ff.detach(), gy] + gp

    method = ctx.adjoint_method
    for i in range(n_t - 1, 0, -1):
      state = [y, a] + g_params
      state = _integrate_system(aug_dynamics, state, t[i], t[i - 1], method, ctx.adjoint_options, ctx.adjoint_rtol,
                                ctx.adjoint_atol)
      a, g_params = state[1], state[2:]
      y = sol[i - 1]                      # torchdiffeq: "update to use our forward-pass estimate of the state"
      a = a + grad_sol[i - 1]
    return (a, None, None, None, None, None, None, None) + tuple(g_params)


# torchdiffeq 0.2.1's adaptive Runge-Kutta family (rk_common.py / dopri5.py / bosh3.py / adaptive_heun.py): order, stage
# rows, solution / error / mid-point weights.  `adaptive_heun` is the reference's default adjoint method
# (reference src/run_GNN.py:334), `dopri5` what its Computers configuration uses (src/best_params.py).
ADAPTIVE_TABLEAUS = {
  'dopri5': dict(order=5, alpha=[1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.], beta=DP_BETA, c_sol=DP_BETA[-1] + [0.],
                 c_error=DP_C_ERROR, c_mid=DP_C_MID),
  'bosh3': dict(order=3, alpha=[1 / 2, 3 / 4, 1.], beta=[[1 / 2], [0., 3 / 4], [2 / 9, 1 / 3, 4 / 9]],
                c_sol=[2 / 9, 1 / 3, 4 / 9, 0.], c_error=[2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8],
                c_mid=[0., 0.5, 0., 0.]),
  'adaptive_heun': dict(order=2, alpha=[1.], beta=[[1.]], c_sol=[0.5, 0.5], c_error=[0.5, -0.5], c_mid=[0.5, 0.]),
}


def _mixed_norm(parts):
  """torchdiffeq's default norm of a tuple state: max over the components of their RMS norms (ONE host read)."""
  return float(torch.stack([_rms(p) for p in parts]).max())


def _lin(base, dt, coefs, ks):
  """base + dt * sum_j coefs[j] * ks[j] for a list-of-tensors state (zero weights skipped)."""
  out = []
  for c in range(len(base)):
    acc = base[c]
    for w, k in zip(coefs, ks):
      if w != 0:
        acc = acc + (dt * w) * k[c]
    out.append(acc)
  return out


def integrate_system_adaptive(f, state, ta, tb, method, rtol, atol, first_step=None, max_num_steps=2 ** 31 - 1,
                              stats=None):
  """RKAdaptiveStepsizeODESolver.integrate over [ta, tb] for a list-of-tensors state, returning the state at tb
  (torchdiffeq 0.2.1 semantics: Hairer initial step with `order - 1`, error ratio in the mixed norm, step factor
  clipped to [0.2, 10] with safety 0.9, the last step overshoots tb and the quartic Hermite fit through
  (y0, y_mid, y1, f0, f1) is evaluated at tb).  tb < ta integrates the time-reversed system, as `_check_inputs` does for a
  decreasing `t`.  As in the library, the derivative handed to the next step is the LAST STAGE's (`f1 = k[..., -1]`),
  also for tableaus that are not first-same-as-last (adaptive_heun)."""
  tab = ADAPTIVE_TABLEAUS[method]
  order, betas = tab['order'], tab['beta']
  fsal = tab['c_sol'][-1] == 0 and list(tab['c_sol'][:-1]) == list(betas[-1])
  sign = 1.0 if tb >= ta else -1.0
  s0, s1 = sign * ta, sign * tb

  def g(s, st):
    ks = f(sign * s, st)
    return ks if sign > 0 else [-k for k in ks]

  n_eval = 1
  f0 = g(s0, state)
  if first_step is None:
    scale = [atol + y.abs() * rtol for y in state]
    d0 = _mixed_norm([y / sc for y, sc in zip(state, scale)])
    d1 = _mixed_norm([k / sc for k, sc in zip(f0, scale)])
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = g(s0 + h0, [y + h0 * k for y, k in zip(state, f0)])
    n_eval += 1
    d2 = _mixed_norm([(b - a) / sc for a, b, sc in zip(f0, f1, scale)]) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / order)
    dt = min(100 * h0, h1)
  else:
    dt = float(first_step)
  y, t0, t1 = state, s0, s0
  interp = None
  n_steps = 0
  while s1 > t1:
    assert n_steps < max_num_steps, 'max_num_steps exceeded ({}>={})'.format(n_steps, max_num_steps)
    assert t1 + dt > t1, 'underflow in dt {}'.format(dt)
    ks = [f0]
    yi = y
    for alpha_i, beta_i in zip(tab['alpha'], betas):
      yi = _lin(y, dt, beta_i, ks)
      ks.append(g(t1 + dt if alpha_i == 1. else t1 + alpha_i * dt, yi))
      n_eval += 1
    y1 = yi if fsal else _lin(y, dt, tab['c_sol'], ks)
    err = _lin([torch.zeros_like(p) for p in y], dt, tab['c_error'], ks)
    ratio = _mixed_norm([e / (atol + rtol * torch.max(a.abs(), b.abs())) for e, a, b in zip(err, y, y1)])
    n_steps += 1
    if ratio <= 1:
      y_mid = _lin(y, dt, tab['c_mid'], ks)
      interp = (t1, t1 + dt, y, y1, y_mid, ks[0], ks[-1], dt)
      y, f0, t0, t1 = y1, ks[-1], t1, t1 + dt
    if ratio == 0:
      dt = dt * 10.0
    else:
      dt = dt * min(10.0, max(0.9 / ratio ** (1.0 / order), 1.0 if ratio < 1 else 0.2))
  if stats is not None:
    stats['n_eval'], stats['n_steps'] = n_eval, n_steps
  if interp is None:
    return list(state)
  ia, ib, ya, yb, ym, fa, fb, h = interp
  if s1 == ib:
    return list(yb)
  x = (s1 - ia) / (ib - ia)
  out = []
  for c in range(len(ya)):
    a = 2 * h * (fb[c] - fa[c]) - 8 * (yb[c] + ya[c]) + 16 * ym[c]
    b = h * (5 * fa[c] - 3 * fb[c]) + 18 * ya[c] + 14 * yb[c] - 32 * ym[c]
    cc = h * (fb[c] - 4 * fa[c]) - 11 * ya[c] - 5 * yb[c] + 16 * ym[c]
    out.append(ya[c] + x * (h * fa[c]) + x ** 2 * cc + x ** 3 * b + x ** 4 * a)
  return out


def _integrate_system(f, state, ta, tb, method, options, rtol, atol):
  """Integrate a list-of-tensors system from ta to tb (tb may be < ta) with torch ops under no_grad."""
  with torch.no_grad():
    if method in ADAPTIVE_TABLEAUS:
      # adaptive solvers ignore the fixed-grid keys the reference always passes (step_size, max_iters)
      return integrate_system_adaptive(f, state, ta, tb, method, rtol, atol, first_step=options.get('first_step'))
    if method in FIXED_STEPS:
      h = float(options.get('step_size', abs(tb - ta)))
      n = max(1, int(torch.ceil(torch.tensor(abs(tb - ta) / h - 1e-9)).item()))
      sign = 1.0 if tb >= ta else -1.0
      cur = ta
      for s in range(n):
        nxt = tb if s == n - 1 else cur + sign * h
        dt = nxt - cur
        if method == 'euler':
          k1 = f(cur, state)
          state = [y + dt * k for y, k in zip(state, k1)]
        elif method == 'midpoint':
          k1 = f(cur, state)
          mid = [y + 0.5 * dt * k for y, k in zip(state, k1)]
          k2 = f(cur + 0.5 * dt, mid)
          state = [y + dt * k for y, k in zip(state, k2)]
        else:
          k1 = f(cur, state)
          k2 = f(cur + dt / 3, [y + dt * a * (1 / 3) for y, a in zip(state, k1)])
          k3 = f(cur + dt * 2 / 3, [y + dt * (b - a * (1 / 3)) for y, a, b in zip(state, k1, k2)])
          k4 = f(nxt, [y + dt * (a - b + c) for y, a, b, c in zip(state, k1, k2, k3)])
          state = [y + (a + 3 * (b + c) + e) * dt * 0.125 for y, a, b, c, e in zip(state, k1, k2, k3, k4)]
        cur = nxt
      return state
    raise ValueError('Invalid adjoint method "{}". Must be one of {}'.format(
      method, sorted(list(FIXED_STEPS) + list(ADAPTIVE_TABLEAUS))))
