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
