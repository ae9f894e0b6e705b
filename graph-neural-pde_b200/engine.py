"""Host-side driver of the CUDA RHS kernels for one ODE function: owns the graph plan, the workspace and the static
state buffers, builds the C structs, launches through the C ABI on torch's current stream, and replays whole
fixed-grid integrations from a CUDA Graph.

Nothing here computes: every arithmetic step of the hot path is a kernel of libgrand_b200.so.  torch is used for
device memory, streams and graph capture.
"""
import ctypes as C
import math

import torch

from . import _native as N
from .graph_plan import GraphPlan

# Dormand-Prince 5(4), torchdiffeq/_impl/dopri5.py (the tableau reference src/early_stop_solver.py:32 imports)
DP_BETA = [
  [1 / 5],
  [3 / 40, 9 / 40],
  [44 / 45, -56 / 15, 32 / 9],
  [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
  [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
  [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_C_ERROR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
              -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.]
DP_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
            187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def _f32(v):
  """Round a python float to fp32 the way torch does when a python scalar meets an fp32 tensor."""
  return float(torch.tensor(v, dtype=torch.float32))


class AttentionSpec:
  """Device view of a SpGraphTransAttentionLayer for the kernels (weights by reference, never copied, except for the
  beltrami exp_kernel whose four Linear layers are packed into one block-diagonal [2 ad, d] pair)."""

  def __init__(self, layer, d, edge_weight_csr=None):
    opt = layer.opt
    self.keep = []
    a = N.GrandAttention()
    a.d, a.ad, a.heads = int(d), int(layer.attention_dim), int(layer.h)
    att_type = opt['attention_type']
    if hasattr(layer, 'native_weights'):
      # layers with their own scorer (GAT: function_GAT_attention.SpGraphAttentionLayer) describe themselves
      score, wq, bq, wk, bk, scalars = layer.native_weights()
    elif opt['beltrami'] and att_type == 'exp_kernel':
      score = 'exp_kernel2'
      wq, bq, wk, bk = layer.packed_beltrami_weights()
      scalars = torch.stack([layer.output_var_x.detach().reshape(()), layer.lengthscale_x.detach().reshape(()),
                             layer.output_var_p.detach().reshape(()), layer.lengthscale_p.detach().reshape(())])
    else:
      if att_type not in ('scaled_dot', 'cosine_sim', 'pearson', 'exp_kernel'):
        raise UnboundLocalError("attention_type {!r} is not defined".format(att_type))  # reference leaves prods unbound
      score = att_type
      wq, bq, wk, bk = layer.Q.weight, layer.Q.bias, layer.K.weight, layer.K.bias
      scalars = None
      if att_type == 'exp_kernel':
        scalars = torch.stack([layer.output_var.detach().reshape(()), layer.lengthscale.detach().reshape(())])
    for t in (wq, bq, wk, bk):
      N.require_cuda(t, 'attention weight')
    a.score_type = N.SCORE_TYPES[score]
    a.norm_type = N.NORM_SQUAREPLUS if (opt['square_plus'] and score != 'gat') else N.NORM_SOFTMAX   # GAT: softmax only
    a.norm_idx = int(opt['attention_norm_idx'])
    self.keep += [wq, bq, wk, bk, scalars, edge_weight_csr]
    a.wq, a.bq, a.wk, a.bk = (N.ptr(t.detach().contiguous() if not t.is_contiguous() else t.detach())
                              for t in (wq, bq, wk, bk))
    a.kernel_scalars = N.ptr(scalars.contiguous()) if scalars is not None else None
    if scalars is not None:
      self.keep.append(scalars)
    a.edge_weight_csr = N.ptr(edge_weight_csr)
    self.struct = a
    self.ad_eff = 2 * a.ad if score == 'exp_kernel2' else a.ad
    self.heads = a.heads
    self.fused = score == 'scaled_dot' and a.norm_type == N.NORM_SOFTMAX and a.norm_idx == 0


class RhsEngine:
  """One per (graph, state width).  Allocation happens here, once; the C side never allocates."""

  def __init__(self, plan: GraphPlan, d: int):
    self.plan = plan
    self.d = int(d)
    self.device = plan.device
    self._ws = None
    self._ws_key = None
    self._ws_keep = None
    self._state = {}
    self._graphs = {}

  # ---- workspace -------------------------------------------------------------------------------------------
  def workspace(self, spec=None):
    key = (spec.ad_eff, spec.heads) if spec is not None else (0, 0)
    if self._ws is not None and self._ws_key is not None and self._ws_key[0] >= key[0] and self._ws_key[1] >= key[1]:
      return self._ws
    p, dev = self.plan, self.device
    ad_eff, heads = key
    f = lambda n: torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)
    keep = dict(q=f(p.n_rows * ad_eff), k=f(p.n_cols * ad_eff), scores=f(p.n_edges * heads),
                seg_stats=f(p.n_cols * heads * 2), split_stats=f(p.n_split_tiles * max(heads, 1) * 2),
                row_stats=f(p.n_split_rows * max(heads, 1) * 2),
                split_partial=f(p.n_split_tiles * self.d),
                split_counter=torch.zeros(max(p.n_split_rows, 1), dtype=torch.int32, device=dev), gmax=f(1),
                lrow_stats=f(p.n_lunits * max(heads, 1) * 2), lpartial=f(p.n_lunits * max(heads, 1) * self.d),
                lcounter=torch.zeros(max(p.n_lrows, 1), dtype=torch.int32, device=dev),
                coop_sync=torch.zeros(4, dtype=torch.int32, device=dev))
    ws = N.GrandWorkspace()
    for name, t in keep.items():
      setattr(ws, name, t.data_ptr())
    self._ws, self._ws_key, self._ws_keep = ws, key, keep
    self._graphs.clear()
    return ws

  def buffer(self, name, rows=None):
    """Persistent [rows, d] fp32 state buffer (stable address across calls => CUDA-Graph replayable)."""
    rows = self.plan.n_cols if rows is None else rows
    t = self._state.get(name)
    if t is None or t.shape[0] != rows:
      t = torch.empty((rows, self.d), dtype=torch.float32, device=self.device)
      self._state[name] = t
      self._graphs.clear()
    return t

  # ---- struct builders -------------------------------------------------------------------------------------
  @staticmethod
  def epilogue(alpha_train, beta_train, x0, no_alpha_sigmoid, f_out=None, y_out=None, coef_f=0.0, terms=()):
    ep = N.GrandEpilogue()
    ep.alpha_train = alpha_train.data_ptr()
    ep.beta_train = beta_train.data_ptr() if beta_train is not None else None
    ep.x0 = N.ptr(x0)
    ep.no_alpha_sigmoid = 1 if no_alpha_sigmoid else 0
    ep.f_out, ep.y_out = N.ptr(f_out), N.ptr(y_out)
    ep.coef_f = coef_f
    assert len(terms) <= N.GRAND_MAX_TERMS
    ep.n_terms = len(terms)
    for i, (t, c) in enumerate(terms):
      ep.term[i] = t.data_ptr()
      ep.coef[i] = c
    return ep

  def _check_state(self, x, what='x'):
    N.require_cuda(x, what)
    if x.dim() != 2 or x.shape[1] != self.d or x.shape[0] != self.plan.n_cols:
      raise N.GrandError('%s has shape %s, the plan expects [%d, %d]' % (what, tuple(x.shape), self.plan.n_cols, self.d))
    if not x.is_contiguous():
      raise N.GrandError('%s must be contiguous' % what)

  # ---- single launches -------------------------------------------------------------------------------------
  def rhs_attention(self, spec, x, ep):
    self._check_state(x)
    ws = self.workspace(spec)
    N.check(N.lib().grand_rhs_attention(C.byref(self.plan.struct()), C.byref(spec.struct), x.data_ptr(), C.byref(ws),
                                        C.byref(ep), N.current_stream(self.device)))

  def rhs_laplacian(self, w_csr, x, ep):
    self._check_state(x)
    ws = self.workspace(None)
    N.check(N.lib().grand_rhs_laplacian(C.byref(self.plan.struct()), self.d, x.data_ptr(), N.ptr(w_csr), C.byref(ws),
                                        C.byref(ep), N.current_stream(self.device)))

  def attention_weights(self, spec, x, want_attention=True, want_prods=True, want_mean_csr=False):
    self._check_state(x)
    ws = self.workspace(spec)
    p = self.plan
    att = torch.empty((p.n_edges, spec.heads), dtype=torch.float32, device=self.device) if want_attention else None
    prods = torch.empty((p.n_edges, spec.heads), dtype=torch.float32, device=self.device) if want_prods else None
    mean_csr = torch.empty((p.n_edges,), dtype=torch.float32, device=self.device) if want_mean_csr else None
    N.check(N.lib().grand_attention_weights(C.byref(p.struct()), C.byref(spec.struct), x.data_ptr(), C.byref(ws),
                                            N.ptr(att), N.ptr(prods), N.ptr(mean_csr),
                                            N.current_stream(self.device)))
    return att, prods, mean_csr

  def attention_backward(self, spec, x, grad_f, alpha_train, no_alpha_sigmoid):
    """Edge phase of the RHS backward (grand_rhs_attention_backward): -> (abar[E] caller order, dQ, dK)."""
    self._check_state(x)
    self._check_state(grad_f, 'grad_f')
    ws = self.workspace(spec)
    p, dev = self.plan, self.device
    abar = torch.empty((p.n_edges,), dtype=torch.float32, device=dev)
    dq = torch.zeros((p.n_rows, spec.ad_eff), dtype=torch.float32, device=dev)
    dk = torch.zeros((p.n_cols, spec.ad_eff), dtype=torch.float32, device=dev)
    row_dot = torch.zeros((max(p.n_split_rows, 1) * spec.heads,), dtype=torch.float32, device=dev)
    N.check(N.lib().grand_rhs_attention_backward(
      C.byref(p.struct()), C.byref(spec.struct), x.data_ptr(), grad_f.data_ptr(), C.byref(ws),
      alpha_train.data_ptr(), 1 if no_alpha_sigmoid else 0, row_dot.data_ptr(), abar.data_ptr(), dq.data_ptr(),
      dk.data_ptr(), N.current_stream(dev)))
    return abar, dq, dk

  def edge_dot(self, x, grad_f, alpha_train, no_alpha_sigmoid):
    """c[p] = alpha grad_f[row_p] . x[col_p] in CSR order (grand_edge_dot): gradient of one Laplacian RHS evaluation
    w.r.t. its per-edge weights."""
    self._check_state(x)
    self._check_state(grad_f, 'grad_f')
    c = torch.empty((self.plan.n_edges,), dtype=torch.float32, device=self.device)
    N.check(N.lib().grand_edge_dot(C.byref(self.plan.struct()), self.d, x.data_ptr(), grad_f.data_ptr(),
                                   alpha_train.data_ptr(), 1 if no_alpha_sigmoid else 0, c.data_ptr(),
                                   N.current_stream(self.device)))
    return c

  def attention_weights_backward(self, spec, x0, c_csr):
    """Softmax + head-mean backward of the attention evaluated at x0, given c = dL/d(mean_h attention) in CSR order
    (grand_attention_weights_backward): -> (dQ, dK)."""
    self._check_state(x0, 'x0')
    ws = self.workspace(spec)
    p, dev = self.plan, self.device
    dq = torch.zeros((p.n_rows, spec.ad_eff), dtype=torch.float32, device=dev)
    dk = torch.zeros((p.n_cols, spec.ad_eff), dtype=torch.float32, device=dev)
    row_dot = torch.zeros((max(p.n_split_rows, 1) * spec.heads,), dtype=torch.float32, device=dev)
    c_csr = c_csr.contiguous()
    N.check(N.lib().grand_attention_weights_backward(
      C.byref(p.struct()), C.byref(spec.struct), x0.data_ptr(), C.byref(ws), c_csr.data_ptr(), row_dot.data_ptr(),
      dq.data_ptr(), dk.data_ptr(), N.current_stream(dev)))
    return dq, dk

  def edge_values_to_csr(self, w):
    N.require_cuda(w, 'edge values')
    w = w.contiguous()
    heads = 1 if w.dim() == 1 else int(w.shape[1])
    out = torch.empty((self.plan.n_edges,), dtype=torch.float32, device=self.device)
    N.check(N.lib().grand_edge_values_to_csr(C.byref(self.plan.struct()), w.data_ptr(), heads, out.data_ptr(),
                                             N.current_stream(self.device)))
    return out

  def lincomb(self, out, terms):
    n = out.numel()
    arr = (C.c_void_p * len(terms))(*[t.data_ptr() for t, _ in terms])
    coef = (C.c_float * len(terms))(*[c for _, c in terms])
    N.check(N.lib().grand_lincomb(n, len(terms), arr, coef, out.data_ptr(), N.current_stream(self.device)))
    return out

  def error_sumsq(self, acc, terms, y0, y1, atol, rtol):
    """acc (device float64 scalar) += sum ((sum_t c_t k_t) / (atol + rtol max(|y0|,|y1|)))^2"""
    n = y0.numel()
    arr = (C.c_void_p * len(terms))(*[t.data_ptr() for t, _ in terms])
    coef = (C.c_float * len(terms))(*[c for _, c in terms])
    N.check(N.lib().grand_error_sumsq(n, len(terms), arr, coef, y0.data_ptr(), y1.data_ptr(), atol, rtol,
                                      acc.data_ptr(), N.current_stream(self.device)))

  # ---- fixed-grid integration ------------------------------------------------------------------------------
  def integrate_fixed(self, spec, w_csr, residual_ep, method, dts, y, use_graph=True):
    """y <- y(t_end) in place (y is the persistent 'y' buffer).  One C call launches every stage; with `use_graph`
    the launch sequence is captured once per (method, grid, buffers) and replayed.  Wrapped in an NVTX range
    (`grand.integrate_fixed`) so that a timeline shows the solver spans."""
    torch.cuda.nvtx.range_push('grand.integrate_fixed[%s x%d]' % (method, len(dts)))
    try:
      return self._integrate_fixed(spec, w_csr, residual_ep, method, dts, y, use_graph)
    finally:
      torch.cuda.nvtx.range_pop()

  def _integrate_fixed(self, spec, w_csr, residual_ep, method, dts, y, use_graph=True):
    self._check_state(y, 'y')
    n_steps = len(dts)
    if n_steps == 0:
      return y
    ws = self.workspace(spec)
    stage = self.buffer('stages', rows=5 * self.plan.n_cols)
    dt_arr = (C.c_float * n_steps)(*dts)

    def launch():
      N.check(N.lib().grand_integrate_fixed(
        C.byref(self.plan.struct()), C.byref(spec.struct) if spec is not None else None, N.ptr(w_csr), self.d,
        C.byref(ws), C.byref(residual_ep), N.METHODS[method], dt_arr, n_steps, y.data_ptr(), stage.data_ptr(),
        N.current_stream(self.device)))

    if not use_graph:
      launch()
      return y
    key = ('fixed', method, tuple(dts), y.data_ptr(), stage.data_ptr(), N.ptr(w_csr),
           residual_ep.x0, residual_ep.alpha_train, residual_ep.beta_train, residual_ep.no_alpha_sigmoid,
           None if spec is None else (spec.struct.wq, spec.struct.wk, spec.struct.bq, spec.struct.bk,
                                      spec.struct.score_type, spec.struct.norm_type, spec.struct.norm_idx,
                                      spec.struct.heads, spec.struct.ad, spec.struct.edge_weight_csr,
                                      spec.struct.kernel_scalars))
    entry = self._graphs.get(key)
    if entry is None:
      # first call: run eagerly (also the warm-up cudaFuncSetAttribute etc. need), capture on the next one
      launch()
      while len(self._graphs) >= 8:       # keys hold device addresses of temporaries: bound the cache, oldest out
        self._graphs.pop(next(iter(self._graphs)))
      self._graphs[key] = 'warm'
      return y
    if entry == 'warm':
      saved = y.clone()
      torch.cuda.synchronize(self.device)
      graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(graph):
        launch()
      y.copy_(saved)
      self._graphs[key] = graph
      entry = graph
    entry.replay()
    return y


def fixed_grid_dts(t0, t1, step_size):
  """Increments of torchdiffeq's fixed grid (FixedGridODESolver._grid_constructor_from_step_size), computed in fp32
  like the reference does (`t = self.t.type_as(x)`, reference src/block_constant.py:36)."""
  t = torch.tensor([t0, t1], dtype=torch.float32)
  niters = torch.ceil((t[1] - t[0]) / step_size + 1).item()
  grid = torch.arange(0, niters, dtype=torch.float32) * step_size + t[0]
  grid[-1] = t[1]
  return [float(v) for v in (grid[1:] - grid[:-1])], grid
