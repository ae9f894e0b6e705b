"""Dense kernels through the C ABI: the tcgen05 / TMEM / TMA projection (csrc/project_tc.cu) and grand_linear (model-shell
encoder / decoder, reference src/GNN.py:23-31, 60-71; Q / K projection, reference
src/function_transformer_attention.py:174-175) against fp64 matmuls.  Bar: 1e-5 of max |reference| (BASELINE.json)."""
import ctypes as C

import pytest
import torch

from __graft_entry__ import load_package

load_package()
from graph_neural_pde_b200 import _native as N  # noqa: E402
from graph_neural_pde_b200.linear import NativeLinear  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
TOL = 1e-5


@pytest.fixture(autouse=True)
def _restore_impl():
  yield
  N.lib().grand_set_projection_impl(-1)


def _rel(got, want):
  return float((got.double() - want).abs().max() / want.abs().max())


@pytest.mark.parametrize('impl', [0, 1, 2, 3])
@pytest.mark.parametrize('n_rows,n_cols,d,ad', [
  (1000, 1000, 128, 32),      # ogbn-arxiv layout (4 heads x 8)
  (300, 517, 80, 16),         # fewer query rows than key rows (a shard: own rows | halo), d not a multiple of 32
  (2485, 2485, 80, 16), (7650, 7650, 64, 32), (5000, 5000, 100, 24), (4096, 4096, 16, 8), (3000, 3000, 128, 128),
  (129, 129, 36, 10),         # 2 ad = 20 is not a multiple of 16: the tensor-memory kernel hands over
  (0, 0, 32, 8),
])
def test_projection_kernels_vs_fp64(impl, n_rows, n_cols, d, ad):
  g = torch.Generator().manual_seed(n_cols + d)
  x = torch.randn(max(n_cols, 1), d, generator=g).to(DEV)
  wq, wk = (torch.randn(ad, d, generator=g) / d ** 0.5).to(DEV), (torch.randn(ad, d, generator=g) / d ** 0.5).to(DEV)
  bq, bk = torch.randn(ad, generator=g).to(DEV), torch.randn(ad, generator=g).to(DEV)
  a = N.GrandAttention()
  a.d, a.ad, a.heads, a.score_type, a.norm_type, a.norm_idx = d, ad, 1, 0, 0, 0
  a.wq, a.bq, a.wk, a.bk = wq.data_ptr(), bq.data_ptr(), wk.data_ptr(), bk.data_ptr()
  N.check(N.lib().grand_set_projection_impl(impl))
  q = torch.full((max(n_rows, 1), ad), float('nan'), device=DEV)
  k = torch.full((max(n_cols, 1), ad), float('nan'), device=DEV)
  N.check(N.lib().grand_project_qk(x.data_ptr(), n_rows, n_cols, C.byref(a), q.data_ptr(), k.data_ptr(),
                                   N.current_stream(DEV)))
  torch.cuda.synchronize()
  if n_cols == 0:
    return
  xd = x[:n_cols].double()
  assert _rel(q[:n_rows], xd[:n_rows] @ wq.double().t() + bq.double()) <= TOL
  assert _rel(k[:n_cols], xd @ wk.double().t() + bk.double()) <= TOL
  assert torch.isnan(q[n_rows:]).all() and torch.isnan(k[n_cols:]).all()      # nothing stored past the row counts


@pytest.mark.parametrize('impl', [1, 2, 3])
@pytest.mark.parametrize('rows,d_in,d_out,bias', [
  (1000, 128, 128, True), (1000, 128, 40, True), (777, 40, 128, False), (257, 1433, 80, True), (100, 80, 7, True),
  (130, 64, 64, True), (5, 8, 16, True),
])
def test_native_linear_forward_and_backward(impl, rows, d_in, d_out, bias):
  N.check(N.lib().grand_set_projection_impl(impl))
  torch.manual_seed(rows + d_in)
  layer = NativeLinear(d_in, d_out, bias=bias).to(DEV)
  ref = torch.nn.Linear(d_in, d_out, bias=bias).to(DEV).double()
  ref.load_state_dict({k: v.double() for k, v in layer.state_dict().items()})
  x = torch.randn(rows, d_in, device=DEV, requires_grad=True)
  xr = x.detach().double().requires_grad_(True)
  w = torch.randn(rows, d_out, device=DEV)
  y, yr = layer(x), ref(xr)
  assert _rel(y.detach(), yr.detach()) <= TOL
  (y * w).sum().backward()
  (yr * w.double()).sum().backward()
  assert _rel(x.grad, xr.grad) <= TOL
  assert _rel(layer.weight.grad, ref.weight.grad) <= 1e-4       # cuBLAS fp32 reduction over the node dimension
  if bias:
    assert _rel(layer.bias.grad, ref.bias.grad) <= 1e-4


def test_native_linear_refuses_cpu_tensors():
  with pytest.raises(N.GrandError, match='CUDA-only'):
    NativeLinear(4, 4)(torch.zeros(2, 4))
