"""CPU oracle for the GRAND / BLEND ODE right-hand side and its integrators.

TEST INFRASTRUCTURE ONLY.  This module is a pure-PyTorch (CPU, fp32 or fp64)
restatement of the arithmetic the reference executes on the hot path.  It is
imported by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` as the *checker / CPU baseline* and by
nothing else; the product path (`graph-neural-pde_b200/`) never routes through
it and fails loudly when the CUDA library is missing.

Parity pinning.  The reference's own in-tree modules cannot be imported with
their real third-party dependencies here (torch_geometric, torch_scatter,
torch_sparse, torchdiffeq are absent from the image).
The restatement below is pinned two ways:
  * against every known-answer test the reference holds for this path
    (`tests/test_oracle_known_answers.py`; reference test/test_transformer_attention.py:92-99,
    :54-72, :110-121, test/test_function_laplacian_diffusion.py:58-85,
    test/test_utils.py:62-77);
  * against golden vectors produced by importing the reference's OWN modules
    (`/root/reference/src/function_transformer_attention.py`, ... unmodified)
    over `oracle/shims` stand-ins for the absent third-party leaves
    (`tests/golden/make_golden.py`, fixtures committed under `tests/golden/`).
The third-party leaves themselves (PyG softmax / add_remaining_self_loops,
torch_sparse.spmm, torch_scatter.scatter_add, torchdiffeq 0.2.x solvers) are
restated from their published algorithms: "third-party leaves: parity pinned on
the reference's call sites, not on the wheels".

Every function cites the reference file:line (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

# ---------------------------------------------------------------------------------------------
# third-party leaves (published algorithms)
# ---------------------------------------------------------------------------------------------


def maybe_num_nodes(index: Tensor, num_nodes: Optional[int] = None) -> int:
  """torch_geometric.utils.num_nodes.maybe_num_nodes (used at src/utils.py:56,106,200)."""
  if num_nodes is not None:
    return int(num_nodes)
  return int(index.max()) + 1 if index.numel() > 0 else 0


def scatter_add(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
  """torch_scatter.scatter_add(src, index, dim=0, dim_size) (call sites src/utils.py:69,120)."""
  out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
  return out.index_add_(0, index, src)


def scatter_max(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
  """torch_scatter.scatter(src, index, dim=0, reduce='max'); empty segments are never read back."""
  out = torch.full((dim_size,) + tuple(src.shape[1:]), -float('inf'), dtype=src.dtype, device=src.device)
  idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
  return out.scatter_reduce_(0, idx, src, reduce='amax', include_self=True)


def add_remaining_self_loops(edge_index: Tensor, edge_weight: Optional[Tensor] = None, fill_value: float = 1.,
                             num_nodes: Optional[int] = None) -> Tuple[Tensor, Optional[Tensor]]:
  """PyG 1.6/1.7 `add_remaining_self_loops` (call sites src/function_transformer_attention.py:18,
  src/utils.py:63,113): existing loops are dropped from the list, N loops are appended at the END,
  an existing loop's weight survives, the others get `fill_value`; weights stay None if None."""
  n = maybe_num_nodes(edge_index, num_nodes)
  row, col = edge_index[0], edge_index[1]
  mask = row != col
  loop_index = torch.arange(0, n, dtype=row.dtype, device=row.device).unsqueeze(0).repeat(2, 1)
  new_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
  if edge_weight is not None:
    inv_mask = ~mask
    loop_weight = torch.full((n,), fill_value, dtype=edge_weight.dtype, device=edge_weight.device)
    remaining = edge_weight[inv_mask]
    if remaining.numel() > 0:
      loop_weight[row[inv_mask]] = remaining
    edge_weight = torch.cat([edge_weight[mask], loop_weight], dim=0)
  return new_index, edge_weight


def segment_softmax(src: Tensor, index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
  """torch_geometric.utils.softmax(src[E,H], index[E]) (call site src/function_transformer_attention.py:213;
  in-repo structural witness src/utils.py:196-208): per segment and per head column
  exp(s - max_seg) / (sum_seg + 1e-16)."""
  n = maybe_num_nodes(index, num_nodes)
  out = src - scatter_max(src, index, n)[index]
  out = out.exp()
  out_sum = scatter_add(out, index, n)[index]
  return out / (out_sum + 1e-16)


def spmm(index: Tensor, value: Tensor, m: int, n: int, matrix: Tensor) -> Tensor:
  """torch_sparse.spmm (call sites src/function_transformer_attention.py:35,
  src/function_laplacian_diffusion.py:31-35): out[r] = sum_{e: row_e = r} value_e * matrix[col_e]."""
  row, col = index[0], index[1]
  out = matrix.index_select(0, col) * value.unsqueeze(-1)
  return scatter_add(out, row, m)


# ---------------------------------------------------------------------------------------------
# in-repo graph normalisation and squareplus
# ---------------------------------------------------------------------------------------------


def squareplus(src: Tensor, index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
  """src/utils.py:179-208: GLOBAL max subtraction, (o + sqrt(o^2 + 4)) / 2, segment sum-normalise, +1e-16."""
  out = src - src.max()
  out = (out + torch.sqrt(out ** 2 + 4)) / 2
  n = maybe_num_nodes(index, num_nodes)
  out_sum = scatter_add(out, index, n)[index]
  return out / (out_sum + 1e-16)


def get_rw_adj(edge_index: Tensor, edge_weight: Optional[Tensor] = None, norm_dim: int = 1, fill_value: float = 0.,
               num_nodes: Optional[int] = None, dtype=None) -> Tuple[Tensor, Tensor]:
  """src/utils.py:105-123: optional self loops, degree over `row` (norm_dim 0) or `col` (norm_dim 1),
  weight / degree."""
  n = maybe_num_nodes(edge_index, num_nodes)
  if edge_weight is None:
    edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
  if not fill_value == 0:
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill_value, n)
  row, col = edge_index[0], edge_index[1]
  indices = row if norm_dim == 0 else col
  deg = scatter_add(edge_weight, indices, n)
  deg_inv = deg.pow(-1)
  edge_weight = deg_inv[indices] * edge_weight if norm_dim == 0 else edge_weight * deg_inv[indices]
  return edge_index, edge_weight


def gcn_norm_fill_val(edge_index: Tensor, edge_weight: Optional[Tensor] = None, fill_value: float = 0.,
                      num_nodes: Optional[int] = None, dtype=None) -> Tuple[Tensor, Tensor]:
  """src/utils.py:55-72: loops only when int(fill_value) != 0; D^-1/2 (A) D^-1/2 with degree over `col`."""
  n = maybe_num_nodes(edge_index, num_nodes)
  if edge_weight is None:
    edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
  if not int(fill_value) == 0:
    edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill_value, n)
  row, col = edge_index[0], edge_index[1]
  deg = scatter_add(edge_weight, col, n)
  deg_inv_sqrt = deg.pow(-0.5)
  deg_inv_sqrt = deg_inv_sqrt.masked_fill(deg_inv_sqrt == float('inf'), 0)
  return edge_index, deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]


# ---------------------------------------------------------------------------------------------
# attention layer (src/function_transformer_attention.py:59-217)
# ---------------------------------------------------------------------------------------------


def _linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
  return torch.nn.functional.linear(x, w, b)


def _cosine(a: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
  """torch.nn.CosineSimilarity(dim=1, eps) with the semantics of the torch installed here (2.x):
  x.y / (max(|x|, eps) * max(|y|, eps))   (SURVEY.md A.4 notes the version dependence)."""
  return torch.nn.functional.cosine_similarity(a, b, dim=1, eps=eps)


def attention_scores(x: Tensor, edge: Tensor, P: Dict[str, Tensor], opt: dict) -> Tensor:
  """Unnormalised per-edge, per-head scores `prods[E,H]`; src/function_transformer_attention.py:128-206.

  P holds the layer parameters under the reference's names: 'Q.weight','Q.bias','K.weight','K.bias'
  (and for exp_kernel 'output_var','lengthscale'; for beltrami+exp_kernel 'Qx.*','Kx.*','Qp.*','Kp.*',
  'output_var_x','lengthscale_x','output_var_p','lengthscale_p')."""
  h = int(opt['heads'])
  att_type = opt['attention_type']
  if opt['beltrami'] and att_type == 'exp_kernel':
    # :133-169 -- feature part and positional part scored by two exp kernels whose product is the score
    label_index = opt['feat_hidden_dim'] + opt['pos_enc_hidden_dim']
    p = x[:, opt['feat_hidden_dim']: label_index]
    xf = torch.cat((x[:, :opt['feat_hidden_dim']], x[:, label_index:]), dim=1)
    ad = P['Qx.weight'].shape[0]
    dk = ad // h
    qx = _linear(xf, P['Qx.weight'], P['Qx.bias']).view(-1, h, dk).transpose(1, 2)
    kx = _linear(xf, P['Kx.weight'], P['Kx.bias']).view(-1, h, dk).transpose(1, 2)
    qp = _linear(p, P['Qp.weight'], P['Qp.bias']).view(-1, h, dk).transpose(1, 2)
    kp = _linear(p, P['Kp.weight'], P['Kp.bias']).view(-1, h, dk).transpose(1, 2)
    src_x, dst_x = qx[edge[0, :], :, :], kx[edge[1, :], :, :]
    src_p, dst_p = qp[edge[0, :], :, :], kp[edge[1, :], :, :]
    return P['output_var_x'] ** 2 * torch.exp(
      -torch.sum((src_x - dst_x) ** 2, dim=1) / (2 * P['lengthscale_x'] ** 2)) \
      * P['output_var_p'] ** 2 * torch.exp(
      -torch.sum((src_p - dst_p) ** 2, dim=1) / (2 * P['lengthscale_p'] ** 2))

  ad = P['Q.weight'].shape[0]
  dk = ad // h
  q = _linear(x, P['Q.weight'], P['Q.bias']).view(-1, h, dk).transpose(1, 2)   # [N, d_k, H]  (:174-188)
  k = _linear(x, P['K.weight'], P['K.bias']).view(-1, h, dk).transpose(1, 2)
  src = q[edge[0, :], :, :]                                                     # :190
  dst_k = k[edge[1, :], :, :]                                                   # :191
  if not opt['beltrami'] and att_type == 'exp_kernel':                          # :193-194
    prods = P['output_var'] ** 2 * torch.exp(-(torch.sum((src - dst_k) ** 2, dim=1) / (2 * P['lengthscale'] ** 2)))
  elif att_type == 'scaled_dot':                                                # :195-196
    prods = torch.sum(src * dst_k, dim=1) / np.sqrt(dk)
  elif att_type == 'cosine_sim':                                                # :197-199
    prods = _cosine(src, dst_k)
  elif att_type == 'pearson':                                                   # :200-206
    src = src - torch.mean(src, dim=1, keepdim=True)
    dst_k = dst_k - torch.mean(dst_k, dim=1, keepdim=True)
    prods = _cosine(src, dst_k)
  else:
    # the reference leaves `prods` unbound here (e.g. beltrami + exp_kernel handled above)
    raise UnboundLocalError("attention_type {!r} not defined for this configuration".format(att_type))
  return prods


def attention_forward(x: Tensor, edge: Tensor, P: Dict[str, Tensor], opt: dict,
                      edge_weights: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
  """SpGraphTransAttentionLayer.forward -> (attention[E,H], prods[E,H]); :128-214."""
  prods = attention_scores(x, edge, P, opt)
  if opt['reweight_attention'] and edge_weights is not None:                    # :208-209
    prods = prods * edge_weights.unsqueeze(dim=1)
  index = edge[opt['attention_norm_idx']]
  if opt['square_plus']:                                                        # :210-211
    attention = squareplus(prods, index)
  else:                                                                         # :212-213
    attention = segment_softmax(prods, index)
  return attention, prods


# ---------------------------------------------------------------------------------------------
# ODE functions (src/function_transformer_attention.py:12-56, src/function_laplacian_diffusion.py:15-51)
# ---------------------------------------------------------------------------------------------


def _alpha(alpha_train: Tensor, opt: dict) -> Tensor:
  return alpha_train if opt['no_alpha_sigmoid'] else torch.sigmoid(alpha_train)


def residual(ax: Tensor, x: Tensor, alpha_train: Tensor, beta_train: Tensor, x0: Optional[Tensor], opt: dict) -> Tensor:
  """f = alpha (ax - x) [+ beta x0]; src/function_transformer_attention.py:46-53,
  src/function_laplacian_diffusion.py:43-51."""
  f = _alpha(alpha_train, opt) * (ax - x)
  if opt['add_source']:
    f = f + beta_train * x0
  return f


def transformer_rhs(x: Tensor, edge_index: Tensor, P: Dict[str, Tensor], opt: dict, alpha_train: Tensor,
                    beta_train: Tensor, x0: Optional[Tensor] = None, edge_weights: Optional[Tensor] = None) -> Tensor:
  """ODEFuncTransformerAtt.forward without the nfe bookkeeping; default (mix_features=False) branch of
  multiply_attention (:34-35): head-mean first, then one spmm over the whole state width."""
  attention, _ = attention_forward(x, edge_index, P, opt, edge_weights)
  mean_attention = attention.mean(dim=1)
  ax = spmm(edge_index, mean_attention, x.shape[0], x.shape[0], x)
  return residual(ax, x, alpha_train, beta_train, x0, opt)


def gat_attention_forward(x: Tensor, edge: Tensor, P: Dict[str, Tensor], opt: dict) -> Tuple[Tensor, Tensor]:
  """SpGraphAttentionLayer.forward -> (attention[E,H], wx[N,ad]); src/function_GAT_attention.py:105-116.
  P: 'W' [d, ad], 'a' [2 d_k, 1, 1] (shared by the heads), 'Wout' [ad, d] (mix_features only)."""
  h = int(opt['heads'])
  ad = P['W'].shape[1]
  dk = ad // h
  wx = torch.mm(x, P['W'])                                                        # :106
  hh = wx.view(-1, h, dk).transpose(1, 2)                                         # [N, d_k, H]  :107-108
  edge_h = torch.cat((hh[edge[0, :], :, :], hh[edge[1, :], :, :]), dim=1).transpose(0, 1)   # [2 d_k, E, H]  :111-112
  edge_e = torch.nn.functional.leaky_relu(torch.sum(P['a'] * edge_h, dim=0), opt['leaky_relu_slope'])   # :113
  attention = segment_softmax(edge_e, edge[opt['attention_norm_idx']])            # :114
  return attention, wx


def gat_rhs(x: Tensor, edge_index: Tensor, P: Dict[str, Tensor], opt: dict, alpha_train: Tensor, beta_train: Tensor,
            x0: Optional[Tensor] = None) -> Tensor:
  """ODEFuncAtt.forward without the nfe bookkeeping; src/function_GAT_attention.py:31-64 (per-head spmm, then the
  mean over heads -- unlike the transformer function, which averages the attention first)."""
  attention, wx = gat_attention_forward(x, edge_index, P, opt)
  n = x.shape[0]
  if opt['mix_features']:                                                         # :32-37
    agg = torch.mean(torch.stack([spmm(edge_index, attention[:, i], n, n, wx) for i in range(opt['heads'])], dim=0),
                     dim=0)
    ax = torch.mm(agg, P['Wout'])
  else:                                                                           # :38-42
    ax = torch.mean(torch.stack([spmm(edge_index, attention[:, i], n, n, x) for i in range(opt['heads'])], dim=0),
                    dim=0)
  return residual(ax, x, alpha_train, beta_train, x0, opt)


def hard_attention_sample(x: Tensor, data_edge_index: Tensor, attention: Tensor, opt: dict, num_nodes: int,
                          training: bool) -> Tuple[Tensor, Tensor]:
  """The edge selection of HardAttODEblock.forward (src/block_transformer_hard_attention.py:49-68):
  -> (odefunc.edge_index, odefunc.attention_weights[E']).  Eval mode keeps every edge with the head-mean attention;
  training keeps the edges whose (optionally flux-weighted) mean attention exceeds the (1 - att_samp_pct) quantile
  and renormalises the kept values over edge_index[attention_norm_idx] (:43-46)."""
  mean_att = attention.mean(dim=1, keepdim=False)
  if not training:
    return data_edge_index, mean_att
  if opt['use_flux']:
    delta = torch.linalg.norm(x[data_edge_index[0, :], :] - x[data_edge_index[1, :], :], dim=1)
    mean_att = mean_att * delta
  threshold = torch.quantile(mean_att, 1 - opt['att_samp_pct'])
  mask = mean_att > threshold
  edge_index = data_edge_index[:, mask]
  kept = mean_att[mask]
  index = edge_index[opt['attention_norm_idx']]
  sums = scatter_add(kept, index, num_nodes)[index]
  return edge_index, kept / (sums + 1e-16)


def laplacian_weights(opt: dict, edge_weight: Optional[Tensor], attention_weights: Optional[Tensor]) -> Tensor:
  """Which per-edge weight LaplacianODEFunc.sparse_multiply uses, by block name; :28-36."""
  if opt['block'] in ['attention']:
    return attention_weights.mean(dim=1)
  if opt['block'] in ['mixed', 'hard_attention']:
    return attention_weights
  return edge_weight


def laplacian_rhs(x: Tensor, edge_index: Tensor, weight: Tensor, opt: dict, alpha_train: Tensor, beta_train: Tensor,
                  x0: Optional[Tensor] = None) -> Tensor:
  """LaplacianODEFunc.forward without the nfe bookkeeping; :38-51."""
  ax = spmm(edge_index, weight, x.shape[0], x.shape[0], x)
  return residual(ax, x, alpha_train, beta_train, x0, opt)


# ---------------------------------------------------------------------------------------------
# integrators: torchdiffeq 0.2.x semantics (in-repo witnesses src/early_stop_solver.py)
# ---------------------------------------------------------------------------------------------


def fixed_grid(t0: float, t1: float, step_size: float, dtype=torch.float32) -> Tensor:
  """FixedGridODESolver._grid_constructor_from_step_size: niters = ceil((t1 - t0)/h + 1),
  grid = arange(niters) * h + t0, last entry set to t1.  Computed in the dtype of `t` (the blocks pass
  `t.type_as(x)`, src/block_constant.py:36)."""
  t = torch.tensor([t0, t1], dtype=dtype)
  niters = torch.ceil((t[1] - t[0]) / step_size + 1).item()
  grid = torch.arange(0, niters, dtype=dtype) * step_size + t[0]
  grid[-1] = t[1]
  return grid


def rk4_alt_step(func: Callable, t0, dt, t1, y0: Tensor) -> Tensor:
  """torchdiffeq rk4_alt_step_func = the 3/8 rule; witnessed at src/early_stop_solver.py:150-155."""
  k1 = func(t0, y0)
  k2 = func(t0 + dt * (1 / 3), y0 + dt * k1 * (1 / 3))
  k3 = func(t0 + dt * (2 / 3), y0 + dt * (k2 - k1 * (1 / 3)))
  k4 = func(t1, y0 + dt * (k1 - k2 + k3))
  return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


def euler_step(func: Callable, t0, dt, t1, y0: Tensor) -> Tensor:
  return dt * func(t0, y0)


def midpoint_step(func: Callable, t0, dt, t1, y0: Tensor) -> Tensor:
  half_dt = 0.5 * dt
  y_mid = y0 + func(t0, y0) * half_dt
  return dt * func(t0 + half_dt, y_mid)


_FIXED_STEPS = {'rk4': rk4_alt_step, 'euler': euler_step, 'midpoint': midpoint_step}


def odeint_fixed(func: Callable, y0: Tensor, t: Tensor, method: str, step_size: float,
                 on_step: Optional[Callable] = None) -> Tensor:
  """FixedGridODESolver.integrate (witness: EarlyStopRK4.integrate, src/early_stop_solver.py:163-184):
  step over the grid, linear interpolation to the requested times."""
  step = _FIXED_STEPS[method]
  grid = fixed_grid(float(t[0]), float(t[-1]), step_size, dtype=t.dtype)
  assert grid[0] == t[0] and grid[-1] == t[-1]
  solution = torch.empty((len(t),) + tuple(y0.shape), dtype=y0.dtype)
  solution[0] = y0
  j = 1
  for ta, tb in zip(grid[:-1], grid[1:]):
    dt = tb - ta
    y1 = y0 + step(func, ta, dt, tb, y0)
    if on_step is not None:
      on_step(ta, tb, y0, y1)
    while j < len(t) and tb >= t[j]:
      if tb == t[j]:
        solution[j] = y1
      else:  # _linear_interp
        slope = (t[j] - ta) / (tb - ta)
        solution[j] = y0 + slope * (y1 - y0)
      j += 1
    y0 = y1
  return solution


# Dormand-Prince 5(4) tableau, torchdiffeq/_impl/dopri5.py (imported by src/early_stop_solver.py:2,32-33)
DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1., 1.]
DP_BETA = [
  [1 / 5],
  [3 / 40, 9 / 40],
  [44 / 45, -56 / 15, 32 / 9],
  [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
  [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
  [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
DP_C_ERROR = [
  35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
  -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1. / 60.,
]
DP_C_MID = [
  6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
  187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2
]


def rms_norm(tensor: Tensor) -> Tensor:
  return tensor.abs().pow(2).mean().sqrt()


def _select_initial_step(func, t0, y0, order, rtol, atol, f0):
  """Hairer's rule as in torchdiffeq/_impl/misc.py:_select_initial_step (one probe eval)."""
  scale = atol + torch.abs(y0) * rtol
  d0 = rms_norm(y0 / scale).abs()
  d1 = rms_norm(f0 / scale).abs()
  if d0 < 1e-5 or d1 < 1e-5:
    h0 = torch.tensor(1e-6, dtype=y0.dtype)
  else:
    h0 = 0.01 * d0 / d1
  h0 = h0.abs()
  y1 = y0 + h0 * f0
  f1 = func(t0 + h0, y1)
  d2 = torch.abs(rms_norm((f1 - f0) / scale) / h0)
  if d1 <= 1e-15 and d2 <= 1e-15:
    h1 = torch.max(torch.tensor(1e-6, dtype=y0.dtype), h0 * 1e-3)
  else:
    h1 = (0.01 / max(d1, d2)) ** (1. / float(order + 1))
  h1 = h1.abs()
  return torch.min(100 * h0, h1).to(torch.float64)


def _dopri5_rk_step(func, y0, f0, t0, dt, t1):
  """_runge_kutta_step with the DP tableau; 6 new evals; FSAL (f1 = k[-1]); y1 = last stage."""
  dty = dt.to(y0.dtype)
  ks = [f0]
  yi = y0
  for alpha_i, beta_i in zip(DP_ALPHA, DP_BETA):
    ti = t1 if alpha_i == 1. else t0 + alpha_i * dt
    kmat = torch.stack(ks, dim=-1)
    coeff = torch.tensor(beta_i, dtype=torch.float64).to(y0.dtype) * dty
    yi = y0 + kmat.matmul(coeff).view_as(f0)
    ks.append(func(ti, yi))
  k = torch.stack(ks, dim=-1)
  y1 = yi
  f1 = ks[-1]
  y1_error = k.matmul(dty * torch.tensor(DP_C_ERROR, dtype=torch.float64).to(y0.dtype))
  return y1, f1, y1_error, k


def _interp_fit(y0, y1, y_mid, f0, f1, dt):
  a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
  b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
  c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
  d = dt * f0
  e = y0
  return [e, d, c, b, a]


def _interp_evaluate(coefficients, t0, t1, t):
  x = ((t - t0) / (t1 - t0)).to(coefficients[0].dtype)
  total = coefficients[0] + x * coefficients[1]
  x_power = x
  for coefficient in coefficients[2:]:
    x_power = x_power * x
    total = total + x_power * coefficient
  return total


def _optimal_step_size(last_step, error_ratio, safety=0.9, ifactor=10.0, dfactor=0.2, order=5):
  if error_ratio == 0:
    return last_step * ifactor
  if error_ratio < 1:
    dfactor = 1.0
  error_ratio = error_ratio.to(last_step.dtype)
  exponent = 1.0 / order
  factor = min(ifactor, max(safety / float(error_ratio) ** exponent, dfactor))
  return last_step * factor


class Dopri5Stats:
  def __init__(self):
    self.accepted = 0
    self.rejected = 0
    self.nfe = 0
    self.step_sizes = []


def odeint_dopri5(func: Callable, y0: Tensor, t: Tensor, rtol: float, atol: float,
                  on_step: Optional[Callable] = None, max_steps: int = 2 ** 31 - 1,
                  stats: Optional[Dopri5Stats] = None) -> Tensor:
  """RKAdaptiveStepsizeODESolver.integrate for dopri5 (witness: EarlyStopDopri5.integrate/advance,
  src/early_stop_solver.py:58-86).  Time-like scalars in float64, state in y0.dtype."""
  t = t.to(torch.float64)
  stats = stats if stats is not None else Dopri5Stats()

  def f(tt, yy):
    stats.nfe += 1
    return func(tt, yy)

  f0 = f(t[0], y0)
  dt = _select_initial_step(f, t[0], y0, 4, rtol, atol, f0)
  state = dict(y1=y0, f1=f0, t0=t[0], t1=t[0], dt=dt, interp=[y0] * 5)
  solution = torch.empty((len(t),) + tuple(y0.shape), dtype=y0.dtype)
  solution[0] = y0
  for i in range(1, len(t)):
    next_t = t[i]
    n_steps = 0
    while next_t > state['t1'] and n_steps < max_steps:
      y0s, f0s, t0s, dts = state['y1'], state['f1'], state['t1'], state['dt']
      t1s = t0s + dts
      assert t0s + dts > t0s, 'underflow in dt {}'.format(float(dts))
      y1, f1, y1_error, k = _dopri5_rk_step(f, y0s, f0s, t0s, dts, t1s)
      error_tol = atol + rtol * torch.max(y0s.abs(), y1.abs())
      error_ratio = rms_norm(y1_error / error_tol).abs()
      accept = bool(error_ratio <= 1)
      stats.step_sizes.append((float(t0s), float(dts), float(error_ratio), accept))
      if accept:
        dty = dts.to(y0.dtype)
        y_mid = y0s + k.matmul(dty * torch.tensor(DP_C_MID, dtype=torch.float64).to(y0.dtype)).view_as(y0s)
        interp = _interp_fit(y0s, y1, y_mid, k[..., 0], k[..., -1], dty)
        state = dict(y1=y1, f1=f1, t0=t0s, t1=t1s, dt=None, interp=interp)
        stats.accepted += 1
      else:
        state = dict(y1=y0s, f1=f0s, t0=t0s, t1=t0s, dt=None, interp=state['interp'])
        stats.rejected += 1
      state['dt'] = _optimal_step_size(dts, error_ratio)
      n_steps += 1
      if on_step is not None and accept:
        on_step(state['t0'], state['t1'], y0s, y1)
    solution[i] = _interp_evaluate(state['interp'], state['t0'], state['t1'], next_t)
  return solution


def odeint(func: Callable, y0: Tensor, t: Tensor, rtol: float = 1e-7, atol: float = 1e-9, method: Optional[str] = None,
           options: Optional[dict] = None, **unused) -> Tensor:
  """torchdiffeq.odeint(func, y0, t, rtol, atol, method, options) -> [len(t), *y0.shape]
  (call sites src/block_constant.py:46-62, src/block_transformer_attention.py:46-64)."""
  options = dict(options or {})
  method = method or 'dopri5'
  if method in _FIXED_STEPS:
    return odeint_fixed(func, y0, t, method, options['step_size'])
  if method == 'dopri5':
    return odeint_dopri5(func, y0, t, rtol, atol)
  raise ValueError('Invalid method "{}"'.format(method))


# ---------------------------------------------------------------------------------------------
# block-level forward: what ConstantODEblock / AttODEblock compute (src/block_constant.py:35-70,
# src/block_transformer_attention.py:36-72)
# ---------------------------------------------------------------------------------------------


def block_tolerances(opt: dict) -> Tuple[float, float]:
  """ODEblock.set_tol: atol = tol_scale * 1e-7, rtol = tol_scale * 1e-9 (src/base_classes.py:56-61)."""
  return opt['tol_scale'] * 1e-7, opt['tol_scale'] * 1e-9


def constant_block_graph(edge_index: Tensor, edge_attr: Optional[Tensor], num_nodes: int, opt: dict, dtype):
  """ConstantODEblock.__init__ graph setup (src/block_constant.py:12-23)."""
  if opt['data_norm'] == 'rw':
    return get_rw_adj(edge_index, edge_weight=edge_attr, norm_dim=1, fill_value=opt['self_loop_weight'],
                      num_nodes=num_nodes, dtype=dtype)
  return gcn_norm_fill_val(edge_index, edge_weight=edge_attr, fill_value=opt['self_loop_weight'],
                           num_nodes=num_nodes, dtype=dtype)


def block_forward(func: Callable, x: Tensor, T: float, opt: dict) -> Tensor:
  """state_dt = odeint(func, x, t=[0,T], method, options{step_size}, atol, rtol); z = state_dt[1]."""
  t = torch.tensor([0, T]).type_as(x)
  atol, rtol = block_tolerances(opt)
  sol = odeint(func, x, t, method=opt['method'], options=dict(step_size=opt['step_size']), atol=atol, rtol=rtol)
  return sol[1]
