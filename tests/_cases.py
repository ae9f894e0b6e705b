"""Replay helpers shared by the oracle (CPU) and CUDA (-m gpu) parity tests."""
import torch

from oracle import grand_oracle as O

# Parity tolerance (BASELINE.json north_star: "within 1e-5 relative fp32").  Relative is taken against the
# tensor's max magnitude: |got - want| <= RTOL * max|want| element-wise (summation order differs between the
# reference's scatter-adds and any other implementation, so a per-element relative bound on near-zero entries
# is not meaningful).
RTOL = 1e-5


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
  want = want.detach().double().cpu()
  got = got.detach().double().cpu()
  denom = want.abs().max().clamp_min(1e-30)
  return float((got - want).abs().max() / denom)


def assert_close(got, want, rtol=RTOL, what=''):
  assert tuple(got.shape) == tuple(want.shape), '%s shape %s vs %s' % (what, tuple(got.shape), tuple(want.shape))
  err = rel_err(got, want)
  assert err <= rtol, '%s: max-normalised error %.3e > %.1e' % (what, err, rtol)


def sub_state(state_dict, prefix):
  return {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}


def cast(tree, dtype):
  if torch.is_tensor(tree):
    return tree.to(dtype) if tree.is_floating_point() else tree
  if isinstance(tree, dict):
    return {k: cast(v, dtype) for k, v in tree.items()}
  return tree


def oracle_attention_layer(case, dtype=torch.float32):
  inp = cast(case['inputs'], dtype)
  P = cast(case['state_dict'], dtype)
  return O.attention_forward(inp['x'], inp['edge'], P, case['opt'], inp['edge_weights'])


def oracle_block(case, dtype=torch.float32, want_extra=False):
  """Returns (rhs at x, z = block(x), nfe of the block forward, (edge_index, edge_weight) of the function) computed by
  the oracle; with want_extra also a dict (hard attention: the sampled 'attention_weights')."""
  opt = case['opt']
  inp = cast(case['inputs'], dtype)
  S = cast(case['state_dict'], dtype)
  x = inp['x']
  n = inp['num_nodes']
  if opt['block'] in ('attention', 'hard_attention'):
    ei, ew = O.get_rw_adj(inp['edge_index'], edge_weight=inp['edge_attr'], norm_dim=1,
                          fill_value=opt['self_loop_weight'], num_nodes=n, dtype=dtype)
  else:
    ei, ew = O.constant_block_graph(inp['edge_index'], inp['edge_attr'], n, opt, dtype)
  alpha, beta = S['odefunc.alpha_train'], S['odefunc.beta_train']
  x0 = x.clone()
  counter = [0]
  extra = {}
  # the function's own layer keeps the ctor-time edge weights: add_remaining_self_loops(data.edge_attr)
  # (src/function_transformer_attention.py:17-23), None when the data has no edge_attr
  if opt['self_loop_weight'] > 0:
    _, lw = O.add_remaining_self_loops(inp['edge_index'], inp['edge_attr'], fill_value=opt['self_loop_weight'])
  else:
    lw = inp['edge_attr']
  if opt['block'] == 'hard_attention':
    # src/block_transformer_hard_attention.py:36-68: attention at x(0) on the full graph (the block's own layer for the
    # Laplacian function, the function's layer otherwise), then the edge selection
    if opt['function'] == 'laplacian':
      att, _ = O.attention_forward(x, ei, sub_state(S, 'multihead_att_layer.'), opt, ew)
    elif opt['function'] == 'GAT':
      att, _ = O.gat_attention_forward(x, ei, sub_state(S, 'odefunc.multihead_att_layer.'), opt)
    else:
      att, _ = O.attention_forward(x, ei, sub_state(S, 'odefunc.multihead_att_layer.'), opt, lw)
    ei, hard_w = O.hard_attention_sample(x, ei, att, opt, n, bool(inp['train']))
    extra['attention_weights'] = hard_w
  if opt['function'] == 'transformer':
    P = sub_state(S, 'odefunc.multihead_att_layer.')

    def func(t, y):
      counter[0] += 1
      return O.transformer_rhs(y, ei, P, opt, alpha, beta, x0, edge_weights=lw)
  elif opt['function'] == 'GAT':
    P = sub_state(S, 'odefunc.multihead_att_layer.')

    def func(t, y):
      counter[0] += 1
      return O.gat_rhs(y, ei, P, opt, alpha, beta, x0)
  else:
    if opt['block'] == 'attention':
      P = sub_state(S, 'multihead_att_layer.')
      att, _ = O.attention_forward(x, ei, P, opt, ew)
      w = O.laplacian_weights(opt, ew, att)
    elif opt['block'] == 'hard_attention':
      w = O.laplacian_weights(opt, ew, hard_w)
    else:
      w = ew

    def func(t, y):
      counter[0] += 1
      return O.laplacian_rhs(y, ei, w, opt, alpha, beta, x0)
  rhs = func(0.0, x)
  counter[0] = 0
  z = O.block_forward(func, x, opt['time'], opt)
  if want_extra:
    return rhs, z, counter[0], (ei, ew), extra
  return rhs, z, counter[0], (ei, ew)


_EXACT_CACHE = None


def exact_block_solution(case, name=None):
  """z(T) of the case's ODE from a tight fp64 dopri5 run of the oracle (the arbiter for adaptive-solver parity); read
  from tests/golden/exact_dopri5.pt (make_exact.py) when the case is in it."""
  global _EXACT_CACHE
  if _EXACT_CACHE is None:
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'exact_dopri5.pt')
    _EXACT_CACHE = torch.load(path, map_location='cpu', weights_only=False) if os.path.exists(path) else {}
  if name is not None and name in _EXACT_CACHE:
    return _EXACT_CACHE[name]
  return compute_exact_block_solution(case)


def compute_exact_block_solution(case):
  holder = {}
  real = O.block_forward

  def capture(func, x, T, opt):
    holder['f'] = func
    return real(func, x, T, {**opt, 'method': 'euler', 'step_size': opt['time'] if opt['time'] > 0 else 1})

  O.block_forward = capture
  try:
    oracle_block(case, torch.float64)
  finally:
    O.block_forward = real
  x = case['inputs']['x'].double()
  t = torch.tensor([0, case['opt']['time']], dtype=torch.float64)
  return O.odeint_dopri5(holder['f'], x, t, 1e-13, 1e-13)[1]
