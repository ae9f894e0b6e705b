"""Gradients of the CUDA path (RHS vector-Jacobian product kernel + solver loops) against autograd through the CPU
oracle in fp64.  Tolerance 2e-4 of the gradient tensor's max magnitude: the backward accumulates with fp32 atomics in a
different order than any reference, and fp32 autograd through the reference's own op graph differs from fp64 by the
same order."""
import pytest
import torch

import _cases as C
from __graft_entry__ import load_package
from oracle import grand_oracle as O

pytestmark = pytest.mark.gpu

load_package()
from graph_neural_pde_b200.block_constant import ConstantODEblock  # noqa: E402
from graph_neural_pde_b200.function_laplacian_diffusion import LaplacianODEFunc  # noqa: E402
from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt  # noqa: E402
from graph_neural_pde_b200.synthetic import Data, default_opt, make_graph  # noqa: E402

DEV = torch.device('cuda:0')
GTOL = 2e-4


def assert_grads_close(got, want, tol, what):
  """Per-tensor max-normalised comparison, except that a gradient which is identically zero in exact arithmetic
  (dL/dK.bias: a constant added to every key shifts all scores of a row equally and the softmax cancels it) is only
  rounding noise on both sides; such tensors are normalised by the largest gradient of the group instead."""
  scale = max(float(w.detach().abs().max()) for w in want)
  for g, w, name in zip(got, want, what):
    assert tuple(g.shape) == tuple(w.shape), name
    err = float((g.detach().double().cpu() - w.detach().double().cpu()).abs().max())
    denom = max(float(w.detach().abs().max()), 1e-3 * scale)
    assert err <= tol * denom, '%s: error %.3e vs tolerance %.1e * %.3e' % (name, err, tol, denom)


def _make(function, n, d, ad, heads, method='rk4', adjoint=False, seed=0, **extra):
  edge_index = make_graph(n, 6 * n, max_degree=150, seed=seed)
  torch.manual_seed(seed + 1)
  x = torch.randn(n, d)
  kw = dict(function=function, block='constant', hidden_dim=d, attention_dim=ad, heads=heads, method=method,
            step_size=0.5, time=1.0, add_source=True, adjoint=adjoint, adjoint_method='rk4', adjoint_step_size=0.5)
  kw.update(extra)
  opt = default_opt(**kw)
  data = Data(x=x, edge_index=edge_index, edge_attr=None, num_nodes=n)
  cls = ODEFuncTransformerAtt if function == 'transformer' else LaplacianODEFunc
  block = ConstantODEblock(cls, (), opt, data, DEV, t=torch.tensor([0, opt['time']])).to(DEV)
  g = torch.Generator().manual_seed(seed + 2)
  with torch.no_grad():
    for p in block.parameters():
      if p.dim() == 2:
        p.copy_((1.5 * torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5).to(DEV))
    block.odefunc.alpha_train.fill_(0.3)
    block.odefunc.beta_train.fill_(-0.4)
  block.train()
  return block, opt, edge_index, x


def _oracle_params(block, dtype=torch.float64):
  sd = {k: v.detach().cpu().to(dtype).requires_grad_(v.dtype.is_floating_point) for k, v in block.state_dict().items()
        if v.dtype.is_floating_point}
  return sd


def _oracle_func(sd, opt, edge_index, n, x0, function):
  ei, ew = O.constant_block_graph(edge_index, None, n, opt, torch.float64)
  alpha, beta = sd['odefunc.alpha_train'], sd['odefunc.beta_train']
  if function == 'transformer':
    P = C.sub_state(sd, 'odefunc.multihead_att_layer.')
    return lambda t, y: O.transformer_rhs(y, ei, P, opt, alpha, beta, x0)
  return lambda t, y: O.laplacian_rhs(y, ei, ew, opt, alpha, beta, x0)


NAMES = {'transformer': ['odefunc.multihead_att_layer.Q.weight', 'odefunc.multihead_att_layer.Q.bias',
                         'odefunc.multihead_att_layer.K.weight', 'odefunc.multihead_att_layer.K.bias',
                         'odefunc.alpha_train', 'odefunc.beta_train'],
         'laplacian': ['odefunc.alpha_train', 'odefunc.beta_train']}


@pytest.mark.parametrize('function,d,ad,heads', [('transformer', 128, 16, 4), ('transformer', 32, 32, 8),
                                                 ('laplacian', 64, 16, 2)])
def test_rhs_vjp_vs_oracle_autograd(function, d, ad, heads):
  n = 400
  block, opt, edge_index, x = _make(function, n, d, ad, heads)
  func = block.odefunc
  xd = x.to(DEV).requires_grad_(True)
  block.set_x0(x.to(DEV))
  torch.manual_seed(9)
  gout = torch.randn(n, d)
  f = func(0.0, xd)
  f.backward(gout.to(DEV))
  # oracle
  sd = _oracle_params(block)
  xo = x.double().requires_grad_(True)
  fo = _oracle_func(sd, opt, edge_index, n, x.double(), function)(0.0, xo)
  fo.backward(gout.double())
  C.assert_close(f, fo, C.RTOL, 'forward value')
  got = dict(block.named_parameters())
  assert_grads_close([xd.grad] + [got[k].grad for k in NAMES[function]],
                     [xo.grad] + [sd[k].grad for k in NAMES[function]], GTOL, ['x'] + NAMES[function])


@pytest.mark.parametrize('function', ['transformer', 'laplacian'])
@pytest.mark.parametrize('method', ['rk4', 'dopri5'])
def test_block_backward_vs_oracle_autograd(function, method):
  n, d = 300, 32
  block, opt, edge_index, x = _make(function, n, d, 16, 4, method=method, tol_scale=1000.0)
  xd = x.to(DEV).requires_grad_(True)
  block.set_x0(x.to(DEV))
  torch.manual_seed(3)
  w = torch.randn(n, d)
  z = block(xd)
  (z * w.to(DEV)).sum().backward()
  sd = _oracle_params(block)
  xo = x.double().requires_grad_(True)
  fo = _oracle_func(sd, opt, edge_index, n, x.double(), function)
  zo = O.block_forward(fo, xo, opt['time'], opt)
  (zo * w.double()).sum().backward()
  tol = GTOL if method == 'rk4' else 2e-2     # dopri5: different accepted-step sequences (see test_gpu_parity)
  C.assert_close(z, zo, C.RTOL if method == 'rk4' else 5e-3, 'z')
  got = dict(block.named_parameters())
  assert_grads_close([xd.grad] + [got[k].grad for k in NAMES[function]],
                     [xo.grad] + [sd[k].grad for k in NAMES[function]], tol, ['x'] + NAMES[function])


ATT_NAMES = ['multihead_att_layer.Q.weight', 'multihead_att_layer.Q.bias', 'multihead_att_layer.K.weight',
             'multihead_att_layer.K.bias', 'odefunc.alpha_train', 'odefunc.beta_train']


@pytest.mark.parametrize('method,max_degree', [('rk4', 150), ('euler', 1500), ('dopri5', 150)])
def test_attention_block_backward_vs_oracle_autograd(method, max_degree):
  """AttODEblock (GRAND-l): the attention is evaluated once at x(0) and frozen (reference
  src/block_transformer_attention.py:36-39); without the adjoint its gradient reaches the block's Q / K and x(0)
  through every RHS evaluation.  max_degree 1500 gives rows that are split across tiles in the backward kernel."""
  from graph_neural_pde_b200.block_transformer_attention import AttODEblock
  n, d, ad, heads = 1600 if max_degree > 1000 else 300, 32, 16, 4
  edge_index = make_graph(n, 6 * n, max_degree=max_degree, seed=11)
  torch.manual_seed(12)
  x = torch.randn(n, d)
  opt = default_opt(function='laplacian', block='attention', hidden_dim=d, attention_dim=ad, heads=heads, method=method,
                    step_size=0.5, time=1.0, add_source=True, adjoint=False, tol_scale=1000.0)
  data = Data(x=x, edge_index=edge_index, edge_attr=None, num_nodes=n)
  block = AttODEblock(LaplacianODEFunc, (), opt, data, DEV, t=torch.tensor([0, opt['time']])).to(DEV)
  g = torch.Generator().manual_seed(13)
  with torch.no_grad():
    for p_ in block.multihead_att_layer.parameters():
      if p_.dim() == 2:
        p_.copy_((1.5 * torch.randn(p_.shape, generator=g) / p_.shape[1] ** 0.5).to(DEV))
      else:
        p_.copy_((0.3 * torch.randn(p_.shape, generator=g)).to(DEV))
    block.odefunc.alpha_train.fill_(0.3)
    block.odefunc.beta_train.fill_(-0.4)
  block.train()
  xd = x.to(DEV).requires_grad_(True)
  block.set_x0(x.to(DEV))
  torch.manual_seed(3)
  w = torch.randn(n, d)
  z = block(xd)
  (z * w.to(DEV)).sum().backward()
  # oracle: autograd through attention_forward at x(0) + the integration, fp64
  sd = _oracle_params(block)
  xo = x.double().requires_grad_(True)
  ei, ew = O.constant_block_graph(edge_index, None, n, {**opt, 'data_norm': 'rw'}, torch.float64)
  att, _ = O.attention_forward(xo, ei, C.sub_state(sd, 'multihead_att_layer.'), opt, ew)
  wts = O.laplacian_weights(opt, ew, att)
  fo = lambda t, y: O.laplacian_rhs(y, ei, wts, opt, sd['odefunc.alpha_train'], sd['odefunc.beta_train'], x.double())
  zo = O.block_forward(fo, xo, opt['time'], opt)
  (zo * w.double()).sum().backward()
  tol = 2e-2 if method == 'dopri5' else GTOL
  C.assert_close(z, zo, 5e-3 if method == 'dopri5' else C.RTOL, 'z')
  got = dict(block.named_parameters())
  for k in ATT_NAMES[:3]:
    assert got[k].grad is not None and float(got[k].grad.abs().max()) > 0, k
  assert_grads_close([xd.grad] + [got[k].grad for k in ATT_NAMES], [xo.grad] + [sd[k].grad for k in ATT_NAMES], tol,
                     ['x'] + ATT_NAMES)


@pytest.mark.parametrize('attention_type,square_plus,norm_idx,reweight', [
  ('scaled_dot', True, 1, False),      # the reference's tuned Cora configuration (src/best_params.py)
  ('exp_kernel', True, 1, False),      # ... Citeseer
  ('cosine_sim', True, 0, False),      # ... Pubmed's scorer
  ('pearson', False, 1, True),
])
def test_attention_block_backward_general_configurations(attention_type, square_plus, norm_idx, reweight):
  """Gradients through the attention frozen at x(0) for the scorers / normalisers the fused backward kernel does not
  cover: forward native, the once-per-step attention backward through attention_autograd.mean_attention, the
  per-evaluation Laplacian backward native.  Against fp64 autograd through the oracle."""
  from graph_neural_pde_b200.block_transformer_attention import AttODEblock
  n, d, ad, heads = 300, 32, 16, 4
  edge_index = make_graph(n, 6 * n, max_degree=150, seed=21)
  torch.manual_seed(22)
  x = torch.randn(n, d)
  opt = default_opt(function='laplacian', block='attention', hidden_dim=d, attention_dim=ad, heads=heads, method='rk4',
                    step_size=0.5, time=1.0, add_source=True, adjoint=False, attention_type=attention_type,
                    square_plus=square_plus, attention_norm_idx=norm_idx, reweight_attention=reweight)
  data = Data(x=x, edge_index=edge_index, edge_attr=None, num_nodes=n)
  block = AttODEblock(LaplacianODEFunc, (), opt, data, DEV, t=torch.tensor([0, opt['time']])).to(DEV)
  g = torch.Generator().manual_seed(23)
  with torch.no_grad():
    for name, p_ in block.multihead_att_layer.named_parameters():
      if p_.dim() == 2:
        p_.copy_((1.5 * torch.randn(p_.shape, generator=g) / p_.shape[1] ** 0.5).to(DEV))
      elif name == 'output_var':
        p_.fill_(1.2)
      elif name == 'lengthscale':
        p_.fill_(1.7)
      else:
        p_.copy_((0.3 * torch.randn(p_.shape, generator=g)).to(DEV))
    block.odefunc.alpha_train.fill_(0.3)
    block.odefunc.beta_train.fill_(-0.4)
  block.train()
  xd = x.to(DEV).requires_grad_(True)
  block.set_x0(x.to(DEV))
  torch.manual_seed(3)
  w = torch.randn(n, d)
  z = block(xd)
  (z * w.to(DEV)).sum().backward()
  sd = _oracle_params(block)
  xo = x.double().requires_grad_(True)
  ei, ew = O.constant_block_graph(edge_index, None, n, {**opt, 'data_norm': 'rw'}, torch.float64)
  att, _ = O.attention_forward(xo, ei, C.sub_state(sd, 'multihead_att_layer.'), opt, ew)
  wts = O.laplacian_weights(opt, ew, att)
  fo = lambda t, y: O.laplacian_rhs(y, ei, wts, opt, sd['odefunc.alpha_train'], sd['odefunc.beta_train'], x.double())
  zo = O.block_forward(fo, xo, opt['time'], opt)
  (zo * w.double()).sum().backward()
  C.assert_close(z, zo, C.RTOL, 'z')
  names = list(ATT_NAMES)
  if attention_type == 'exp_kernel':
    names += ['multihead_att_layer.output_var', 'multihead_att_layer.lengthscale']
  got = dict(block.named_parameters())
  for k in names:
    assert got[k].grad is not None, k
  assert_grads_close([xd.grad] + [got[k].grad for k in names], [xo.grad] + [sd[k].grad for k in names], GTOL,
                     ['x'] + names)


def test_run_gnn_trains_attention_block(capsys):
  """run_GNN.main with --block attention --function laplacian (the paper's GRAND-l) trains: the start-up check lets
  it through and the loss falls."""
  from graph_neural_pde_b200 import run_GNN
  args = run_GNN.build_parser().parse_args(['--dataset', 'Cora', '--function', 'laplacian', '--block', 'attention',
                                            '--method', 'rk4', '--time', '1.0', '--step_size', '0.5', '--epoch', '6',
                                            '--hidden_dim', '32', '--attention_dim', '16', '--heads', '4', '--lr', '0.02',
                                            '--input_dropout', '0.0', '--quiet'])
  torch.manual_seed(0)
  run_GNN.main(vars(args))
  out = capsys.readouterr().out
  losses = [float(line.split('Loss ')[1].split(',')[0]) for line in out.splitlines() if line.startswith('Epoch:')]
  assert len(losses) == 5 and losses[-1] < losses[0], losses


def test_run_gnn_trains_the_reference_cora_configuration(capsys):
  """The reference's tuned Cora configuration (src/best_params.py through `--best_params reference`; `--block attention`
  given explicitly, as its CLI default would otherwise override the block): GRAND-l with the attention frozen at x(0),
  scaled_dot + squareplus over attention_norm_idx 1, dopri5 to T = 18.3 without the adjoint, early-stopping model.
  Trains on the synthetic Cora stand-in: finite losses that fall."""
  from graph_neural_pde_b200 import run_GNN
  args = run_GNN.build_parser().parse_args(['--dataset', 'Cora', '--best_params', 'reference', '--block', 'attention',
                                            '--epoch', '6', '--quiet'])
  torch.manual_seed(0)
  run_GNN.main(vars(args))
  out = capsys.readouterr().out
  losses = [float(line.split('Loss ')[1].split(',')[0]) for line in out.splitlines() if line.startswith('Epoch:')]
  assert len(losses) == 5 and all(l == l and abs(l) < 1e6 for l in losses), losses
  assert min(losses[1:]) < losses[0], losses


def test_run_gnn_trains_end_to_end(capsys):
  """The reference's entry point (run_GNN.main) on the synthetic Cora stand-in: a few epochs of GRAND-nl with the
  default early-stopping model; the loss must fall and the attribute chain main() reads must be alive."""
  from graph_neural_pde_b200 import run_GNN
  args = run_GNN.build_parser().parse_args(['--dataset', 'Cora', '--function', 'transformer', '--block', 'constant',
                                            '--method', 'rk4', '--time', '1.0', '--step_size', '0.5', '--epoch', '6',
                                            '--hidden_dim', '32', '--attention_dim', '16', '--heads', '4', '--lr', '0.02',
                                            '--input_dropout', '0.0', '--quiet'])
  torch.manual_seed(0)
  train_acc, val_acc, test_acc = run_GNN.main(vars(args))
  out = capsys.readouterr().out
  losses = [float(line.split('Loss ')[1].split(',')[0]) for line in out.splitlines() if line.startswith('Epoch:')]
  assert len(losses) == 5 and losses[-1] < losses[0], losses
  assert 'backward nfe' in out and 0 <= val_acc <= 1 and 0 <= test_acc <= 1
  fwd_nfe = int(out.splitlines()[0].split('forward nfe ')[1].split(',')[0])
  assert fwd_nfe == 8                      # 2 rk4 steps x 4 stages in the training forward


def test_adjoint_gradients_close_to_direct_gradients():
  """odeint_adjoint (reverse-time rk4 of the augmented system) vs backprop through the forward stages: both
  discretise the same continuous sensitivity, so they agree to the step's truncation error, not to rounding."""
  n, d = 300, 32
  grads = {}
  for adjoint in (False, True):
    block, opt, edge_index, x = _make('transformer', n, d, 16, 4, method='rk4', adjoint=adjoint, seed=5)
    xd = x.to(DEV).requires_grad_(True)
    block.set_x0(x.to(DEV))
    torch.manual_seed(3)
    w = torch.randn(n, d).to(DEV)
    z = block(xd)
    (z * w).sum().backward()
    grads[adjoint] = [xd.grad.clone()] + [dict(block.named_parameters())[k].grad.clone() for k in NAMES['transformer']]
    nfe_after_backward = block.odefunc.nfe
    assert nfe_after_backward > 8 or not adjoint      # the adjoint pass evaluates the function again (bm meter)
  assert_grads_close(grads[True], grads[False], 5e-2, ['x'] + NAMES['transformer'])


@pytest.mark.parametrize('adjoint_method', ['adaptive_heun', 'dopri5'])
def test_adaptive_adjoint_gradients_close_to_direct_gradients(adjoint_method):
  """The reference's DEFAULT adjoint method is adaptive_heun (src/run_GNN.py:334; Computers uses dopri5): the augmented
  system (y, a_y, a_theta) integrated backwards by the adaptive solver (autograd_rhs.integrate_system_adaptive, every
  RHS evaluation and its VJP on the native kernels) vs backprop through a fine fixed-grid forward of the same block."""
  n, d = 300, 32
  grads = {}
  for adjoint in (False, True):
    block, opt, edge_index, x = _make('transformer', n, d, 16, 4, method='rk4', adjoint=adjoint, seed=6,
                                      step_size=0.125, adjoint_method=adjoint_method, tol_scale_adjoint=1000.0)
    xd = x.to(DEV).requires_grad_(True)
    block.set_x0(x.to(DEV))
    torch.manual_seed(3)
    w = torch.randn(n, d).to(DEV)
    z = block(xd)
    (z * w).sum().backward()
    grads[adjoint] = [xd.grad.clone()] + [dict(block.named_parameters())[k].grad.clone() for k in NAMES['transformer']]
    for gr in grads[adjoint]:
      assert torch.isfinite(gr).all()
  assert_grads_close(grads[True], grads[False], 2e-2, ['x'] + NAMES['transformer'])
