"""Backward of the transformer function in the configurations outside the fused kernel (autograd_rhs.TransformerRhsGeneralFn
-> attention_autograd.transformer_rhs), on the CPU: the native forward is stood in for by the oracle (the engine cannot
run here), everything else -- which gradients are produced, in which order they reach the parameters, the values -- is
the product code, against fp64 autograd through the oracle (reference src/function_transformer_attention.py:38-53)."""
import types

import pytest
import torch

from __graft_entry__ import load_package
from oracle import grand_oracle as O

load_package()
from graph_neural_pde_b200 import _native as N  # noqa: E402
from graph_neural_pde_b200 import attention_autograd as AA  # noqa: E402
from graph_neural_pde_b200 import autograd_rhs as AG  # noqa: E402
from graph_neural_pde_b200 import function_transformer_attention as FT  # noqa: E402
from graph_neural_pde_b200.synthetic import Data, default_opt, make_graph  # noqa: E402


class FakeEngine:
  """rhs_attention writes the oracle's forward into the epilogue's f_out."""

  def __init__(self, func, x0):
    self.func, self.x0 = func, x0

  def rhs_attention(self, spec, x, ep):
    func = self.func
    P = {k: v.detach() for k, v in func.multihead_att_layer.state_dict().items()}
    with torch.no_grad():
      ep['f_out'].copy_(O.transformer_rhs(x, func.edge_index, P, func.opt, func.alpha_train.detach(),
                                           func.beta_train.detach(), self.x0, func.multihead_att_layer.edge_weights))


@pytest.mark.parametrize('attention_type,square_plus,norm_idx,reweight,d', [
  ('scaled_dot', True, 1, False, 12), ('cosine_sim', False, 0, True, 12), ('pearson', True, 0, False, 12),
  ('exp_kernel', False, 1, False, 12), ('scaled_dot', False, 0, False, 10),     # fused scorer, width not a multiple of 4
])
def test_general_transformer_backward_vs_oracle_autograd(monkeypatch, attention_type, square_plus, norm_idx, reweight, d):
  monkeypatch.setattr(N, 'require_cuda', lambda t, what: None)
  monkeypatch.setattr(AA.N, 'require_cuda', lambda t, what: None)
  n, ad, h = 70, 16, 4
  edge_index = make_graph(n, 260, max_degree=25, seed=5)
  torch.manual_seed(1)
  x = torch.randn(n, d, dtype=torch.float64)
  x0 = torch.randn(n, d, dtype=torch.float64)
  w = torch.rand(edge_index.shape[1], dtype=torch.float64) + 0.5
  opt = default_opt(function='transformer', block='constant', hidden_dim=d, attention_dim=ad, heads=h,
                    attention_type=attention_type, square_plus=square_plus, attention_norm_idx=norm_idx,
                    reweight_attention=reweight, add_source=True, self_loop_weight=0)
  data = Data(x=x, edge_index=edge_index, edge_attr=w if reweight else None, num_nodes=n)
  func = FT.ODEFuncTransformerAtt(d, d, opt, data, torch.device('cpu')).double()
  with torch.no_grad():
    for name, p in func.named_parameters():
      if p.dim() == 2:
        p.copy_(torch.randn(p.shape, dtype=torch.float64) / p.shape[1] ** 0.5)
      elif name.endswith('output_var'):
        p.fill_(1.2)
      elif name.endswith('lengthscale'):
        p.fill_(1.7)
    func.alpha_train.fill_(0.3)
    func.beta_train.fill_(-0.4)
  func.x0 = x0
  func.train()
  eng = FakeEngine(func, x0)
  monkeypatch.setattr(func, 'engine', lambda xx: eng)
  monkeypatch.setattr(func, 'stage_x0', lambda e: None)
  monkeypatch.setattr(func, 'residual_epilogue', lambda e, x0b, f_out=None: {'f_out': f_out})
  monkeypatch.setattr(func.multihead_att_layer, 'spec',
                      lambda e: types.SimpleNamespace(fused=(attention_type == 'scaled_dot' and not square_plus
                                                             and norm_idx == 0 and not reweight)))
  xg = x.clone().requires_grad_(True)
  g = torch.randn(n, d, dtype=torch.float64)
  f = func(torch.tensor(0.0), xg)
  assert isinstance(f.grad_fn, AG.TransformerRhsGeneralFn._backward_cls)
  (f * g).sum().backward()
  # arbiter: autograd through the oracle with detached copies of the same parameters
  P = {k: v.detach().clone().requires_grad_(True) for k, v in func.multihead_att_layer.state_dict().items()}
  alpha = func.alpha_train.detach().clone().requires_grad_(True)
  beta = func.beta_train.detach().clone().requires_grad_(True)
  xo = x.clone().requires_grad_(True)
  fo = O.transformer_rhs(xo, func.edge_index, P, opt, alpha, beta, x0, func.multihead_att_layer.edge_weights)
  assert float((f.detach() - fo.detach()).abs().max()) < 1e-12
  (fo * g).sum().backward()
  assert float((xg.grad - xo.grad).abs().max() / xo.grad.abs().max()) < 1e-10
  assert float((func.alpha_train.grad - alpha.grad).abs().max()) < 1e-10 * float(alpha.grad.abs().max() + 1)
  assert float((func.beta_train.grad - beta.grad).abs().max()) < 1e-10 * float(beta.grad.abs().max() + 1)
  got = dict(func.multihead_att_layer.named_parameters())
  checked = 0
  for k, want in P.items():
    if want.grad is None:
      assert got[k].grad is None or float(got[k].grad.abs().max()) == 0, k      # V / Wout: dead on this path
      continue
    assert got[k].grad is not None, k
    assert float((got[k].grad - want.grad).abs().max()) <= 1e-9 * float(want.grad.abs().max() + 1e-30) + 1e-14, k
    checked += 1
  assert checked >= 4


@pytest.mark.parametrize('mix_features,norm_idx', [(False, 0), (False, 1), (True, 0)])
def test_gat_backward_vs_oracle_autograd(monkeypatch, mix_features, norm_idx):
  """Gradients through the GAT function (reference src/function_GAT_attention.py:31-64), incl. mix_features: native forward
  stood in for by the oracle, backward = attention_autograd.gat_rhs through autograd_rhs.GeneralRhsFn."""
  from graph_neural_pde_b200 import function_GAT_attention as FG
  monkeypatch.setattr(N, 'require_cuda', lambda t, what: None)
  monkeypatch.setattr(AA.N, 'require_cuda', lambda t, what: None)
  n, d, ad, h = 60, 12, 16, 4
  edge_index = make_graph(n, 220, max_degree=25, seed=6)
  torch.manual_seed(2)
  x = torch.randn(n, d, dtype=torch.float64)
  x0 = torch.randn(n, d, dtype=torch.float64)
  opt = default_opt(function='GAT', block='constant', hidden_dim=d, attention_dim=ad, heads=h, mix_features=mix_features,
                    attention_norm_idx=norm_idx, add_source=True, self_loop_weight=0, leaky_relu_slope=0.2)
  data = Data(x=x, edge_index=edge_index, edge_attr=None, num_nodes=n)
  func = FG.ODEFuncAtt(d, d, opt, data, torch.device('cpu')).double()
  with torch.no_grad():
    for p in func.multihead_att_layer.parameters():
      p.copy_(torch.randn(p.shape, dtype=torch.float64) * 0.4)
    func.alpha_train.fill_(0.3)
    func.beta_train.fill_(-0.4)
  func.x0 = x0
  func.train()

  def oracle_forward(xx):
    P = {k: v.detach() for k, v in func.multihead_att_layer.state_dict().items()}
    with torch.no_grad():
      return O.gat_rhs(xx, func.edge_index, P, opt, func.alpha_train.detach(), func.beta_train.detach(), x0)

  monkeypatch.setattr(func, 'native_forward', oracle_forward)
  monkeypatch.setattr(func, 'engine', lambda xx: 'the engine')
  xg = x.clone().requires_grad_(True)
  g = torch.randn(n, d, dtype=torch.float64)
  f = func(torch.tensor(0.0), xg)
  (f * g).sum().backward()
  P = {k: v.detach().clone().requires_grad_(True) for k, v in func.multihead_att_layer.state_dict().items()}
  alpha = func.alpha_train.detach().clone().requires_grad_(True)
  beta = func.beta_train.detach().clone().requires_grad_(True)
  xo = x.clone().requires_grad_(True)
  fo = O.gat_rhs(xo, func.edge_index, P, opt, alpha, beta, x0)
  assert float((f.detach() - fo.detach()).abs().max()) < 1e-12
  (fo * g).sum().backward()
  assert float((xg.grad - xo.grad).abs().max() / xo.grad.abs().max()) < 1e-10
  assert float((func.alpha_train.grad - alpha.grad).abs().max()) < 1e-10 * float(alpha.grad.abs().max() + 1)
  got = dict(func.multihead_att_layer.named_parameters())
  for k in ('W', 'a') + (('Wout',) if mix_features else ()):
    assert got[k].grad is not None, k
    assert float((got[k].grad - P[k].grad).abs().max()) <= 1e-9 * float(P[k].grad.abs().max()) + 1e-14, k
