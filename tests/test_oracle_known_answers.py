"""The reference's own known-answer / property tests for the hot path, restated on the oracle without PyG imports.

  test/test_transformer_attention.py:92-99   K3 + constant init -> attention == 0.5 exactly
  test/test_transformer_attention.py:54-72   segment sums == 1 along `not attention_norm_idx`, 0 < a <= 1
  test/test_transformer_attention.py:110-121 mean of per-head spmm == spmm of head-mean
  test/test_function_laplacian_diffusion.py:58-85, test/test_utils.py:62-77   get_rw_adj == l1-normalise(A + wI)
"""
import numpy as np
import torch
from sklearn.preprocessing import normalize

from oracle import grand_oracle as O

EDGE = torch.tensor([[0, 2, 2, 1], [1, 0, 1, 2]])
X = torch.tensor([[1., 2.], [3., 2.], [4., 5.]])
EDGE1 = torch.tensor([[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]])
X1 = torch.ones((3, 2))

OPT = {'heads': 2, 'attention_dim': 32, 'attention_type': 'scaled_dot', 'beltrami': False, 'square_plus': False,
       'attention_norm_idx': 0, 'reweight_attention': False}


def const_init_params(in_features, ad, seed=0):
  """A.7: Linear weights = 1e-5, biases at the PyTorch default U(+-1/sqrt(in))."""
  g = torch.Generator().manual_seed(seed)
  bound = 1 / np.sqrt(in_features)
  P = {}
  for n in 'QK':
    P[n + '.weight'] = torch.full((ad, in_features), 1e-5)
    P[n + '.bias'] = (torch.rand(ad, generator=g) * 2 - 1) * bound
  return P


def dense(edge, val, n):
  a = torch.zeros(n, n, dtype=val.dtype)
  a.index_put_((edge[0], edge[1]), val, accumulate=True)
  return a


def test_symmetric_attention_is_half():
  P = const_init_params(2, 32)
  att, _ = O.attention_forward(X1, EDGE1, P, OPT)
  assert att.shape == (6, 2)
  assert torch.all(torch.eq(att, 0.5 * torch.ones((6, 2))))


def test_segment_sums_and_range():
  P = const_init_params(2, 32)
  for idx in (0, 1):
    for sp in (False, True):
      opt = {**OPT, 'attention_norm_idx': idx, 'square_plus': sp}
      att, _ = O.attention_forward(X, EDGE, P, opt)
      assert att.shape == (4, 2)
      for h in range(2):
        d = dense(EDGE, att[:, h], 3)
        s = d.sum(dim=int(not idx))
        assert torch.allclose(s, torch.ones(3), atol=1e-6)
      assert torch.all(att > 0.) and torch.all(att <= 1.)


def test_head_aggregation_linearity():
  P = const_init_params(2, 32, seed=3)
  att, _ = O.attention_forward(X, EDGE, P, OPT)
  ax1 = torch.mean(torch.stack([O.spmm(EDGE, att[:, i], 3, 3, X) for i in range(2)], dim=0), dim=0)
  ax2 = O.spmm(EDGE, att.mean(dim=1), 3, 3, X)
  assert torch.all(torch.isclose(ax1, ax2))


def test_get_rw_adj_matches_l1_normalise():
  base = dense(EDGE, torch.ones(4), 3).numpy()
  for self_loop in [0, 0.3, 1, 3.2]:
    for norm_dim in (0, 1):
      ei, ew = O.get_rw_adj(EDGE, norm_dim=norm_dim, fill_value=self_loop, num_nodes=3, dtype=torch.float32)
      got = dense(ei, ew, 3).numpy()
      want = normalize(base + self_loop * np.identity(3), norm='l1', axis=[1, 0][norm_dim])
      assert np.allclose(want, got)


def test_gcn_norm_is_symmetric_normalisation():
  ei, ew = O.gcn_norm_fill_val(EDGE1, None, 1, 3, dtype=torch.float32)
  a = dense(EDGE1, torch.ones(6), 3).numpy() + np.identity(3)
  dsq = np.sqrt(a.sum(1))
  assert np.allclose(dense(ei, ew, 3).numpy(), a / dsq[:, None] / dsq[None, :])


def test_fixed_grid_matches_survey_counts():
  assert len(O.fixed_grid(0, 1.0, 0.1)) - 1 == 10          # C1/C2: 10 steps -> 40 evals
  g = O.fixed_grid(0, 18.294754260552843, 1)               # Cora best_params: 19 steps -> 76 evals
  assert len(g) - 1 == 19 and abs(float(g[-1] - g[-2]) - 0.2947) < 1e-3
  g = O.fixed_grid(0, 5.08, 1)                             # arxiv: 6 steps -> 24 evals
  assert len(g) - 1 == 6


def test_rk4_order_on_linear_ode():
  # dy/dt = -y, y(0) = 1: the 3/8 rule is 4th order
  f = lambda t, y: -y
  y0 = torch.ones(1, dtype=torch.float64)
  t = torch.tensor([0., 1.], dtype=torch.float64)
  e1 = abs(float(O.odeint_fixed(f, y0, t, 'rk4', 0.1)[1]) - np.exp(-1))
  e2 = abs(float(O.odeint_fixed(f, y0, t, 'rk4', 0.05)[1]) - np.exp(-1))
  assert 12 < e1 / e2 < 20
  st = O.Dopri5Stats()
  z = O.odeint_dopri5(f, y0, t, 1e-7, 1e-9, stats=st)
  assert abs(float(z[1]) - np.exp(-1)) < 1e-6 and st.nfe == 2 + 6 * (st.accepted + st.rejected)
