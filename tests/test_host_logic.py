"""CPU-only tests: the C-ABI library loads and exports every symbol include/grand_b200.h declares (no compute calls),
the host mirror's index plumbing, option handling, state_dict compatibility and loud failure without CUDA."""
import copy
import ctypes
import os
import re

import pytest
import torch

import _cases as C
from __graft_entry__ import ROOT, load_package

load_package()
from graph_neural_pde_b200 import _native as N  # noqa: E402
from graph_neural_pde_b200 import utils as U  # noqa: E402
from graph_neural_pde_b200.block_constant import ConstantODEblock  # noqa: E402
from graph_neural_pde_b200.block_transformer_attention import AttODEblock  # noqa: E402
from graph_neural_pde_b200.engine import fixed_grid_dts  # noqa: E402
from graph_neural_pde_b200.function_laplacian_diffusion import LaplacianODEFunc  # noqa: E402
from graph_neural_pde_b200.function_transformer_attention import (ODEFuncTransformerAtt,  # noqa: E402
                                                                   SpGraphTransAttentionLayer)
from graph_neural_pde_b200.graph_plan import GraphPlan  # noqa: E402
from graph_neural_pde_b200.model_configurations import (BlockNotDefined, FunctionNotDefined, set_block,  # noqa: E402
                                                        set_function)
from graph_neural_pde_b200.synthetic import Data, default_opt, make_graph  # noqa: E402
from oracle import grand_oracle as O


def test_library_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, 'include', 'grand_b200.h')).read()
  declared = set(re.findall(r'^(?:int|int64_t|void|const char\*|const int32_t\*)\s+(grand_[a-z0-9_]+)\s*\(', header, flags=re.M))
  assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
  assert os.path.exists(N.LIB_PATH), 'build the library first: python graph-neural-pde_b200/build_native.py'
  handle = ctypes.CDLL(N.LIB_PATH)
  for name in declared:
    assert hasattr(handle, name), name
  assert N.lib().grand_abi_version() == N.ABI_VERSION     # pure host call, no GPU needed
  assert int(re.search(r'#define GRAND_TILE (\d+)', header).group(1)) == N.GRAND_TILE
  assert int(re.search(r'#define GRAND_MAX_TERMS (\d+)', header).group(1)) == N.GRAND_MAX_TERMS
  assert int(re.search(r'#define GRAND_MAX_HEADS (\d+)', header).group(1)) == N.GRAND_MAX_HEADS


def test_argument_validation_happens_before_any_cuda_call():
  g = N.GrandGraph()
  g.n_rows, g.n_cols, g.n_edges, g.n_tiles = 4, 2, 0, 1     # n_cols < n_rows
  ep = N.GrandEpilogue()
  rc = N.lib().grand_rhs_laplacian(ctypes.byref(g), 8, 1, 1, None, ctypes.byref(ep), None)
  assert rc == -1 and b'inconsistent sizes' in N.lib().grand_last_error()


def test_graph_plan_invariants():
  n = 300
  ei = make_graph(n, 1500, max_degree=200, seed=1)
  ei = torch.cat([ei, torch.stack([torch.arange(n), torch.arange(n)])], dim=1)     # loops appended at the end
  p = GraphPlan(ei, n, device='cpu')
  row, col = ei[0], ei[1]
  eid = p.eid.long()
  # CSR order is a stable sort by row
  assert torch.equal(p.erow.long(), row[eid]) and torch.equal(p.col.long(), col[eid])
  assert bool((p.erow[1:] >= p.erow[:-1]).all())
  same_row = p.erow[1:] == p.erow[:-1]
  assert bool((eid[1:][same_row] > eid[:-1][same_row]).all()), 'stable within a row'
  deg = torch.bincount(row, minlength=n)
  assert torch.equal(p.rowptr.long()[1:] - p.rowptr.long()[:-1], deg)
  # tiles: cover every edge exactly once, <= GRAND_TILE edges, every row in exactly one packed window or in the
  # chunks of its long row, chunk tiles first
  covered = torch.zeros(p.n_edges, dtype=torch.int32)
  rows_seen = torch.zeros(n, dtype=torch.int32)
  for t in range(p.n_tiles):
    r, b, nr, s = int(p.tile_row[t]), int(p.tile_beg[t]), int(p.tile_nrows[t]), int(p.tile_split[t])
    assert (s >= 0) == (t < p.n_split_tiles)
    if s >= 0:
      assert nr == 1 and int(p.split_tile0[s]) <= t < int(p.split_tile0[s + 1])
      e = min(b + N.GRAND_TILE, int(p.rowptr[r + 1]))
      assert int(p.rowptr[r]) <= b < e
      rows_seen[r] = 1
    else:
      assert 1 <= nr <= 8 and b == int(p.rowptr[r])
      e = int(p.rowptr[r + nr])
      assert e - b <= N.GRAND_TILE
      assert bool((rows_seen[r:r + nr] == 0).all())
      rows_seen[r:r + nr] += 1
    covered[b:e] += 1
  assert bool((covered == 1).all()) and bool((rows_seen == 1).all())
  assert int((deg > N.GRAND_TILE).sum()) == p.n_split_rows > 0
  # CSC view
  cpos = p.csc_pos.long()
  assert bool((p.col.long()[cpos][1:] >= p.col.long()[cpos][:-1]).all())
  assert torch.equal(torch.sort(cpos)[0], torch.arange(p.n_edges))
  cdeg = torch.bincount(col, minlength=n)
  assert torch.equal(p.colptr.long()[1:] - p.colptr.long()[:-1], cdeg)


def test_stream_work_list_covers_every_edge_once():
  """Units of the persistent edge kernel (csrc/rhs_stream.cu): long segments, medium rows, packed windows."""
  n = 400
  ei = make_graph(n, 3000, max_degree=350, seed=2)
  ei = torch.cat([ei, torch.stack([torch.arange(n), torch.arange(n)])], dim=1)
  for seg in (32, 64, 256):
    p = GraphPlan(ei, n, device='cpu', seg_edges=seg)
    deg = (p.rowptr[1:] - p.rowptr[:-1]).long()
    covered = torch.zeros(p.n_edges, dtype=torch.int32)
    rows_done = torch.zeros(n, dtype=torch.int32)
    sd = p.seg_desc.long()
    assert sd.shape == (p.n_lunits + p.n_munits, 4)
    for u in range(sd.shape[0]):
      row, beg, cnt, split = (int(v) for v in sd[u])
      assert N.GRAND_TILE < int(deg[row]) and 0 < cnt <= seg
      assert int(p.rowptr[row]) <= beg and beg + cnt <= int(p.rowptr[row + 1])
      covered[beg:beg + cnt] += 1
      if u < p.n_lunits:
        assert int(deg[row]) > seg and int(p.lrow_row[split]) == row
        assert int(p.lrow_slot0[split]) <= u < int(p.lrow_slot0[split + 1])
        assert beg == int(p.rowptr[row]) + (u - int(p.lrow_slot0[split])) * seg
      else:
        assert split == -1 and cnt == int(deg[row]) <= seg and beg == int(p.rowptr[row])
        rows_done[row] += 1
    rows_done[p.lrow_row.long()] += 1
    if p.n_lunits + p.n_munits > 1:   # longest first inside each class
      c = sd[:, 2]
      first = sd[p.lrow_slot0[:-1].long(), 0] if p.n_lrows else sd[:0, 0]
      assert bool((deg[first][1:] <= deg[first][:-1]).all())
      assert bool((c[p.n_lunits:][1:] <= c[p.n_lunits:][:-1]).all())
    for t in range(p.n_split_tiles, p.n_tiles):
      r, b, nr = int(p.tile_row[t]), int(p.tile_beg[t]), int(p.tile_nrows[t])
      covered[b:int(p.rowptr[r + nr])] += 1
      rows_done[r:r + nr] += 1
    assert bool((covered == 1).all()) and bool((rows_done == 1).all())
    assert p.n_lrows == int((deg > seg).sum()) and p.n_munits == int(((deg > N.GRAND_TILE) & (deg <= seg)).sum())


def test_plan_with_isolated_nodes_and_no_edges():
  p = GraphPlan(torch.zeros((2, 0), dtype=torch.long), 5, device='cpu')
  assert int(p.tile_nrows.sum()) == 5 and p.n_edges == 0 and p.n_split_tiles == 0
  p = GraphPlan(torch.tensor([[0, 3], [3, 0]]), 6, device='cpu')
  assert int(p.tile_nrows.sum()) == 6 and p.n_tiles == 1


def test_fixed_grid_matches_oracle_and_survey():
  for T, h, steps in [(1.0, 0.1, 10), (18.294754260552843, 1, 19), (5.08, 1, 6), (1.6, 0.5, 4), (3.0, 1.0, 3)]:
    dts, grid = fixed_grid_dts(0.0, T, h)
    want = O.fixed_grid(0.0, T, h)
    assert len(dts) == steps and torch.equal(grid, want)


def test_graph_normalisation_matches_golden(golden):
  for name, c in golden.items():
    i = c['inputs']
    if c['kind'] == 'get_rw_adj':
      ei, ew = U.get_rw_adj(i['edge_index'], i['edge_weight'], i['norm_dim'], i['fill_value'], i['num_nodes'],
                            dtype=torch.float32)
    elif c['kind'] == 'gcn_norm_fill_val':
      ei, ew = U.gcn_norm_fill_val(i['edge_index'], i['edge_weight'], i['fill_value'], i['num_nodes'],
                                   dtype=torch.float32)
    else:
      continue
    assert torch.equal(ei, c['outputs']['edge_index']), name
    assert torch.equal(ew, c['outputs']['edge_weight']), name


def test_state_dict_keys_match_reference(golden):
  """Modules are constructible without a GPU; their state_dicts must load the reference's checkpoints strictly."""
  dev = torch.device('cpu')
  seen = 0
  for name, c in golden.items():
    if c['kind'] != 'block':
      continue
    opt = copy.deepcopy(c['opt'])
    inp = c['inputs']
    data = Data(x=inp['x'], edge_index=inp['edge_index'], edge_attr=inp['edge_attr'], num_nodes=inp['num_nodes'])
    func_cls = set_function(opt)
    block_cls = set_block(opt)
    block = block_cls(func_cls, (), opt, data, dev, t=torch.tensor([0, opt['time']]))
    block.load_state_dict(c['state_dict'], strict=True)
    if opt['block'] == 'hard_attention' and inp['train']:
      # the golden edge set is the one a training forward sampled; a fresh block holds the full graph
      assert block.odefunc.edge_index.shape[1] == c['outputs']['odefunc_edge_weight'].shape[0]
    else:
      assert torch.equal(block.odefunc.edge_index, c['outputs']['odefunc_edge_index'])
    C.assert_close(block.odefunc.edge_weight, c['outputs']['odefunc_edge_weight'], 1e-7, name)
    seen += 1
  assert seen >= 20


def test_registry():
  assert set_block({'block': 'constant'}) is ConstantODEblock and set_block({'block': 'attention'}) is AttODEblock
  assert set_function({'function': 'laplacian'}) is LaplacianODEFunc
  assert set_function({'function': 'transformer'}) is ODEFuncTransformerAtt
  with pytest.raises(BlockNotDefined):
    set_block({'block': 'nope'})
  with pytest.raises(FunctionNotDefined):
    set_function({'function': 'nope'})
  from graph_neural_pde_b200.block_transformer_hard_attention import HardAttODEblock
  from graph_neural_pde_b200.function_GAT_attention import ODEFuncAtt
  assert set_block({'block': 'hard_attention'}) is HardAttODEblock and set_function({'function': 'GAT'}) is ODEFuncAtt
  with pytest.raises(NotImplementedError):
    set_block({'block': 'mixed'})


def test_heads_must_divide_attention_dim():
  """reference test/test_attention_ode_block.py:51-56 relies on the AssertionError"""
  opt = default_opt(heads=3, attention_dim=16)
  with pytest.raises(AssertionError):
    SpGraphTransAttentionLayer(8, 8, opt, torch.device('cpu'))


def test_constant_init_and_attention_dim_fallback():
  opt = default_opt(heads=2)
  del opt['attention_dim']
  layer = SpGraphTransAttentionLayer(6, 6, opt, torch.device('cpu'))
  assert layer.attention_dim == 6 and layer.d_k == 3
  assert torch.all(layer.Q.weight == 1e-5) and torch.all(layer.K.weight == 1e-5) and torch.all(layer.Wout.weight == 1e-5)


def test_product_path_fails_loudly_without_cuda():
  n, d = 10, 8
  ei = make_graph(n, 20, seed=0)
  opt = default_opt(function='transformer', hidden_dim=d, attention_dim=8, heads=2, method='rk4', step_size=1)
  x = torch.randn(n, d)
  data = Data(x=x, edge_index=ei, edge_attr=None, num_nodes=n)
  block = ConstantODEblock(ODEFuncTransformerAtt, (), opt, data, torch.device('cpu'), t=torch.tensor([0, 1.0]))
  block.eval()
  block.set_x0(x)
  with torch.no_grad():
    with pytest.raises(N.GrandError, match='no CPU fallback'):
      block.odefunc(0.0, x)
    with pytest.raises(N.GrandError, match='no CPU fallback'):
      block(x)


def test_nfe_guard_counts_like_the_reference():
  """`if nfe > max_nfe: raise` before the increment: max_nfe + 1 evaluations pass, the next one raises."""
  opt = default_opt(function='laplacian', hidden_dim=4, max_nfe=5)
  data = Data(x=torch.zeros(3, 4), edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=None, num_nodes=3)
  f = LaplacianODEFunc(4, 4, opt, data, torch.device('cpu'))
  f._bump_nfe(6)
  assert f.nfe == 6
  with pytest.raises(U.MaxNFEException):
    f._bump_nfe()


def test_hard_attention_selection_host_logic(golden, monkeypatch):
  """HardAttODEblock.select_edges (quantile threshold, optional flux weighting, mask, per-node renormalisation) on the
  CPU with the attention values taken from the oracle: the kept edge set must equal the reference's exactly and the
  renormalised weights to rounding (reference src/block_transformer_hard_attention.py:43-68)."""
  from graph_neural_pde_b200.block_transformer_hard_attention import HardAttODEblock
  dev = torch.device('cpu')
  names = [n for n, c in golden.items() if c['kind'] == 'block' and c['opt']['block'] == 'hard_attention']
  assert len(names) == 6
  for name in names:
    c = golden[name]
    opt = copy.deepcopy(c['opt'])
    inp = c['inputs']
    data = Data(x=inp['x'], edge_index=inp['edge_index'], edge_attr=inp['edge_attr'], num_nodes=inp['num_nodes'])
    block = HardAttODEblock(set_function(opt), (), opt, data, dev, t=torch.tensor([0, opt['time']]))
    block.load_state_dict(c['state_dict'], strict=True)
    block.train(bool(inp['train']))
    S = c['state_dict']
    if opt['function'] == 'laplacian':
      att, _ = O.attention_forward(inp['x'], block.data_edge_index, C.sub_state(S, 'multihead_att_layer.'), opt,
                                   block.odefunc.edge_weight)
    else:
      att, _ = O.attention_forward(inp['x'], block.data_edge_index, C.sub_state(S, 'odefunc.multihead_att_layer.'), opt,
                                   None)
    monkeypatch.setattr(block, 'get_attention_weights', lambda x, att=att: att)
    edge_index, weights = block.select_edges(inp['x'])
    assert torch.equal(edge_index, c['outputs']['odefunc_edge_index']), name
    C.assert_close(weights, c['outputs']['attention_weights'], 2e-6, name)
    block._share(edge_index, weights)
    assert block.reg_odefunc.odefunc.attention_weights is weights and block.odefunc.edge_index is edge_index


def _pack_rows_python(deg, tile, max_rows):
  """Straightforward restatement of the packing rule, the checker for the native grand_plan_pack_rows."""
  first, count = [], []
  start, edges, rows = -1, 0, 0
  for r, dr in enumerate(deg):
    if dr > tile:
      if rows:
        first.append(start); count.append(rows); rows = 0
      continue
    if rows and (edges + dr > tile or rows == max_rows):
      first.append(start); count.append(rows); rows = 0
    if rows == 0:
      start, edges = r, 0
    edges += dr
    rows += 1
  if rows:
    first.append(start); count.append(rows)
  return first, count


def test_native_row_packing_matches_python_rule():
  import numpy as np
  from graph_neural_pde_b200.graph_plan import pack_short_rows
  rng = np.random.default_rng(0)
  cases = [np.zeros(0, dtype=np.int64), np.zeros(5, dtype=np.int64), np.full(40, 32), np.full(7, 33),
           rng.integers(0, 50, size=1000), rng.integers(0, 6, size=3000), np.array([0, 0, 40, 0, 31, 1, 1, 100, 2])]
  for deg in cases:
    for tile, max_rows in [(32, 8), (32, 1), (16, 4)]:
      first, count = pack_short_rows(deg, tile, max_rows)
      want_first, want_count = _pack_rows_python(deg.tolist(), tile, max_rows)
      assert first.tolist() == want_first and count.tolist() == want_count
      covered = sum(count.tolist())
      assert covered == int((np.asarray(deg) <= tile).sum())          # every short row in exactly one window
  with pytest.raises(Exception):
    pack_short_rows(np.array([-1, 2]), 32, 8)


def test_model_shell_matches_reference(golden, monkeypatch):
  """Encoder / decoder either side of the ODE block (reference src/GNN.py:17-77, src/base_classes.py:98-131), eval
  mode, every shell option: `embed` must reproduce the tensor the reference hands to its block and `readout` the
  reference's logits from the reference's block output (tests/golden/make_golden_v3.py recorded both).  This is the
  shell's WIRING (dropout / MLP / concatenation / augmentation order) on the CPU: the dense layers themselves are the
  native library's (linear.NativeLinear refuses CPU tensors; tests/test_gpu_models.py runs them on the GPU), so here
  they are stood in for by the plain matmul."""
  from graph_neural_pde_b200 import linear as L
  from graph_neural_pde_b200.models import GNN
  with pytest.raises(Exception, match='CUDA-only'):
    L.NativeLinear(3, 2)(torch.zeros(4, 3))
  monkeypatch.setattr(L.NativeLinear, 'forward', lambda self, x: torch.nn.functional.linear(x, self.weight, self.bias))
  from graph_neural_pde_b200.utils import DummyDataset
  names = sorted(k for k, v in golden.items() if v['kind'] == 'model_shell')
  assert len(names) == 8
  for name in names:
    c = golden[name]
    opt, inp, out = copy.deepcopy(c['opt']), c['inputs'], c['outputs']
    # the reference constructor rewrote hidden_dim in place (beltrami / use_labels); undo it for a fresh construction
    if opt['use_labels']:
      opt['hidden_dim'] -= inp['num_classes']
    data = Data(x=inp['x'][:, :inp['num_features']], edge_index=inp['edge_index'], edge_attr=None,
                num_nodes=inp['num_nodes'])
    model = GNN(opt, DummyDataset(data, inp['num_classes']), torch.device('cpu'))
    model.load_state_dict(c['state_dict'], strict=True)
    model.eval()
    with torch.no_grad():
      x0 = model.embed(inp['x'], inp['pos_encoding'])
      C.assert_close(x0, out['block_in'], 2e-6, name + ' embed')
      logits = model.readout(out['block_out'], x0.shape[1])
      C.assert_close(logits, out['logits'], 2e-6, name + ' readout')


def test_cli_flags_and_defaults_match_reference():
  """Every flag of the reference's run_GNN.py (src/run_GNN.py:283-436, recorded by tests/golden/make_cli_defaults.py)
  parses here with the same default and type, so reference command lines and opt dictionaries carry over."""
  import importlib
  import json
  import os
  parser = importlib.import_module('graph_neural_pde_b200.run_GNN').build_parser()
  mine = {o[2:]: a for a in parser._actions for o in a.option_strings if o.startswith('--')}
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cli_defaults.json')
  ref = json.load(open(path))
  assert len(ref) >= 100
  for name, spec in ref.items():
    assert name in mine, 'flag --%s missing' % name
    if spec['default'] != '<expr>':
      assert mine[name].default == spec['default'], '--%s default %r vs %r' % (name, mine[name].default, spec['default'])
    if 'type' in spec:
      assert mine[name].type.__name__ == spec['type'], '--%s type' % name
    if 'dest' in spec:
      assert mine[name].dest == spec['dest'], '--%s dest' % name
  opt = vars(parser.parse_args(['--block', 'hard_attention', '--att_samp_pct', '0.5', '--use_flux']))
  assert opt['att_samp_pct'] == 0.5 and opt['use_flux'] is True


def test_reference_best_params_overlay():
  """`--best_params reference` = the reference's {**cmd_opt, **best_params_dict[dataset]} + merge_cmd_args
  (src/run_GNN.py:188-221): table values win, except the few command-line values the reference lets through."""
  import importlib
  mod = importlib.import_module('graph_neural_pde_b200.run_GNN')
  parse = lambda argv: vars(mod.build_parser().parse_args(argv))
  opt = mod.reference_best_params(parse(['--dataset', 'ogbn-arxiv', '--best_params', 'reference']))
  got = (opt['function'], opt['block'], opt['method'], opt['adjoint'], opt['adjoint_method'])
  # reference quirk kept: --block / --function have non-None CLI defaults, so merge_cmd_args lets the DEFAULT
  # 'constant' override the table's 'hard_attention' unless --block is given
  assert got == ('laplacian', 'constant', 'dopri5', True, 'rk4')
  assert mod.reference_best_params(parse(['--dataset', 'ogbn-arxiv', '--block', 'hard_attention']))['block'] == \
      'hard_attention'
  assert opt['hidden_dim'] == 162 and opt['heads'] == 2 and abs(opt['att_samp_pct'] - 0.8105268910037231) < 1e-12
  opt = mod.reference_best_params(parse(['--dataset', 'Cora', '--time', '3', '--method', 'rk4', '--step_size', '0.5']))
  assert (opt['block'], opt['time'], opt['method'], opt['step_size']) == ('constant', 3.0, 'rk4', 0.5)
  assert opt['hidden_dim'] == 80 and opt['square_plus'] is True          # table wins over the CLI default
  assert mod.reference_best_params(parse(['--dataset', 'Nowhere']))['hidden_dim'] == 16
  # and against the reference's own merge_cmd_args, exec'd from its source by tests/golden/make_cli_defaults.py
  import json
  import os
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cli_overlay_cases.json')
  for case in json.load(open(path)):
    got = mod.reference_best_params(parse(case['argv']))
    for key, want in case['opt'].items():
      assert got[key] == want, '%s: %s = %r, reference %r' % (case['argv'], key, got[key], want)


@pytest.mark.parametrize('n,m,max_degree,n_cols,seg', [
  (400, 1500, 200, None, 64), (900, 3000, 350, None, 32), (900, 3000, 350, None, 256), (50, 20, 8, None, 64),
  (300, 2500, 290, 450, 64),      # a shard: columns index [own | halo] rows
  (30000, 120000, 3000, None, 512),   # enough edges for the multi-threaded counting sorts (> 2^16)
  (7, 0, 4, None, 64),            # no edges at all: every row still gets a window
  (0, 0, 4, None, 64),
])
def test_native_plan_equals_the_torch_restatement(n, m, max_degree, n_cols, seg):
  """grand_plan_create (csrc/plan_host.cu) against tests/_plan_reference.py (the torch index-op builder it replaced):
  every array of the plan bit for bit -- CSR / eid, windows, per-lane records, CSC view, work list."""
  from _plan_reference import TorchPlan
  if m > 0:
    ei = make_graph(n, m, max_degree=max_degree, seed=n + m)
    if n_cols is not None:        # redirect a third of the columns into the halo range
      g = torch.Generator().manual_seed(5)
      halo = torch.randint(n, n_cols, (ei.shape[1],), generator=g)
      ei = torch.stack([ei[0], torch.where(torch.rand(ei.shape[1], generator=g) < 0.33, halo, ei[1])])
  else:
    ei = torch.zeros(2, 0, dtype=torch.long)
  got = GraphPlan(ei, n, device='cpu', num_cols=n_cols, seg_edges=seg)
  want = TorchPlan(ei, n, device='cpu', num_cols=n_cols, seg_edges=seg)
  for name in ('n_rows', 'n_cols', 'n_edges', 'n_tiles', 'n_split_rows', 'n_split_tiles', 'max_degree', 'seg_edges',
               'n_lunits', 'n_munits', 'n_lrows'):
    assert getattr(got, name) == getattr(want, name), name
  for name in ('rowptr', 'col', 'erow', 'eid', 'tile_row', 'tile_beg', 'tile_split', 'tile_nrows', 'win_hdr', 'win_rec',
               'split_tile0', 'colptr', 'csc_pos', 'seg_desc', 'lrow_slot0', 'lrow_row'):
    a, b = getattr(got, name), getattr(want, name)
    assert a.dtype == torch.int32 and tuple(a.shape) == tuple(b.shape), (name, a.shape, b.shape)
    assert torch.equal(a, b), name
  assert torch.equal(got.eid_long, want.eid_long)


def test_native_plan_rejects_bad_edge_lists():
  with pytest.raises(AssertionError, match='out of range'):
    GraphPlan(torch.tensor([[0, 5], [1, 1]]), 3, device='cpu')
  with pytest.raises(AssertionError, match='out of range'):
    GraphPlan(torch.tensor([[0, 1], [1, -1]]), 3, device='cpu')


def test_function_keeps_the_engines_of_the_last_two_edge_lists(monkeypatch):
  """ODEFunc.engine: reassigning `edge_index` drops the current engine (reference src/run_GNN.py:252-254 reassigns it),
  but going back to a recently used edge-list OBJECT (hard attention: full graph in evaluation, a new sample every
  training forward) reuses its engine instead of rebuilding the plan."""
  from graph_neural_pde_b200 import base_classes as BC
  built = []

  class FakePlan:
    def __init__(self, ei, num_nodes, device):
      built.append(ei)

  class FakeEngine:
    def __init__(self, plan, d):
      self.d, self.device = d, torch.device('cpu')

  monkeypatch.setattr(BC, 'GraphPlan', FakePlan)
  monkeypatch.setattr(BC, 'RhsEngine', FakeEngine)
  f = BC.ODEFunc(default_opt(), Data(x=torch.zeros(4, 8), edge_index=torch.zeros(2, 0, dtype=torch.long), edge_attr=None,
                                     num_nodes=4), torch.device('cpu'))
  x = torch.zeros(4, 8)
  full = torch.tensor([[0, 1], [1, 0]])
  f.edge_index = full
  e_full = f.engine(x)
  assert f.engine(x) is e_full and len(built) == 1
  for k in range(3):                                   # train forward (new sample), then evaluation (full graph)
    sample = torch.tensor([[0], [1]])
    f.edge_index = sample
    e_s = f.engine(x)
    assert e_s is not e_full and len(built) == 2 + k
    f.edge_index = full
    assert f.engine(x) is e_full and len(built) == 2 + k     # no rebuild of the full graph's plan
  other = torch.tensor([[0, 1], [1, 0]])               # equal content, different object: a new plan, as before
  f.edge_index = other
  assert f.engine(x) is not e_full and len(built) == 5
  assert f.engine(torch.zeros(4, 12)) is not f.engine(x)      # another state width: another engine
