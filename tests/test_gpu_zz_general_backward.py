"""Backward of the transformer function outside the fused configuration, on the GPU: native forward (general pipeline),
backward through attention_autograd.transformer_rhs (autograd_rhs.TransformerRhsGeneralFn), against fp64 autograd through
the oracle.  (This file sorts last on purpose: it was written after the round's GPU budget was spent; its CPU twin with the
native forward stood in for is tests/test_general_backward_cpu.py.)"""
import pytest
import torch

import _cases as C
from __graft_entry__ import load_package
from oracle import grand_oracle as O

load_package()
import test_gpu_backward as B  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


@pytest.mark.parametrize('extra,d', [
  (dict(attention_type='scaled_dot', square_plus=True, attention_norm_idx=1), 32),
  (dict(attention_type='cosine_sim', square_plus=False, attention_norm_idx=0), 32),
  (dict(attention_type='pearson', square_plus=True, attention_norm_idx=0), 32),
  (dict(attention_type='scaled_dot', square_plus=False, attention_norm_idx=0), 30),     # width not a multiple of 4
])
def test_transformer_rhs_general_backward_vs_oracle_autograd(extra, d):
  n = 300
  block, opt, edge_index, x = B._make('transformer', n, d, 16, 4, **extra)
  func = block.odefunc
  xd = x.to(DEV).requires_grad_(True)
  block.set_x0(x.to(DEV))
  torch.manual_seed(9)
  gout = torch.randn(n, d)
  f = func(0.0, xd)
  f.backward(gout.to(DEV))
  sd = B._oracle_params(block)
  xo = x.double().requires_grad_(True)
  fo = B._oracle_func(sd, opt, edge_index, n, x.double(), 'transformer')(0.0, xo)
  fo.backward(gout.double())
  C.assert_close(f, fo, C.RTOL, 'forward value')
  got = dict(block.named_parameters())
  names = B.NAMES['transformer']
  B.assert_grads_close([xd.grad] + [got[k].grad for k in names], [xo.grad] + [sd[k].grad for k in names], B.GTOL,
                       ['x'] + names)


@pytest.mark.parametrize('mix_features', [False, True])
def test_gat_rhs_backward_vs_oracle_autograd(mix_features):
  """Gradients through the GAT function (reference src/function_GAT_attention.py:31-64): native forward, backward through
  attention_autograd.gat_rhs."""
  from graph_neural_pde_b200.block_constant import ConstantODEblock
  from graph_neural_pde_b200.function_GAT_attention import ODEFuncAtt
  from graph_neural_pde_b200.synthetic import Data, default_opt, make_graph
  n, d, ad, h = 300, 32, 16, 4
  edge_index = make_graph(n, 6 * n, max_degree=150, seed=31)
  torch.manual_seed(32)
  x = torch.randn(n, d)
  opt = default_opt(function='GAT', block='constant', hidden_dim=d, attention_dim=ad, heads=h, method='rk4', step_size=0.5,
                    time=1.0, add_source=True, mix_features=mix_features)
  data = Data(x=x, edge_index=edge_index, edge_attr=None, num_nodes=n)
  block = ConstantODEblock(ODEFuncAtt, (), opt, data, DEV, t=torch.tensor([0, opt['time']])).to(DEV)
  g = torch.Generator().manual_seed(33)
  with torch.no_grad():
    for p in block.odefunc.multihead_att_layer.parameters():
      p.copy_((0.4 * torch.randn(p.shape, generator=g)).to(DEV))
    block.odefunc.alpha_train.fill_(0.3)
    block.odefunc.beta_train.fill_(-0.4)
  block.train()
  func = block.odefunc
  xd = x.to(DEV).requires_grad_(True)
  block.set_x0(x.to(DEV))
  torch.manual_seed(9)
  gout = torch.randn(n, d)
  f = func(0.0, xd)
  f.backward(gout.to(DEV))
  sd = B._oracle_params(block)
  xo = x.double().requires_grad_(True)
  fo = O.gat_rhs(xo, func.edge_index.cpu(), C.sub_state(sd, 'odefunc.multihead_att_layer.'), opt,
                 sd['odefunc.alpha_train'], sd['odefunc.beta_train'], x.double())
  fo.backward(gout.double())
  C.assert_close(f, fo, C.RTOL, 'forward value')
  names = ['odefunc.multihead_att_layer.W', 'odefunc.multihead_att_layer.a', 'odefunc.alpha_train', 'odefunc.beta_train']
  if mix_features:
    names.append('odefunc.multihead_att_layer.Wout')
  got = dict(block.named_parameters())
  B.assert_grads_close([xd.grad] + [got[k].grad for k in names], [xo.grad] + [sd[k].grad for k in names], B.GTOL,
                       ['x'] + names)
