"""The oracle (oracle/grand_oracle.py) replayed against golden vectors recorded from the reference's own modules
(tests/golden/make_golden.py).  CPU only; this is what pins the oracle."""
import pytest
import torch

import _cases as C
from oracle import grand_oracle as O

ORACLE_RTOL = 2e-6  # fp32 oracle vs fp32 reference op graph: same ops, same order -> a few ulp


def _names(golden, kind):
  return sorted(k for k, v in golden.items() if v['kind'] == kind)


def test_fixture_inventory(golden):
  kinds = {}
  for v in golden.values():
    kinds[v['kind']] = kinds.get(v['kind'], 0) + 1
  assert kinds['attention_layer'] >= 34 and kinds['block'] >= 20 and kinds['get_rw_adj'] == 16
  assert kinds['gcn_norm_fill_val'] == 8 and kinds['model_early'] == 4


def test_graph_normalisation(golden):
  for name in _names(golden, 'get_rw_adj'):
    c = golden[name]
    i = c['inputs']
    ei, ew = O.get_rw_adj(i['edge_index'], i['edge_weight'], i['norm_dim'], i['fill_value'], i['num_nodes'],
                          dtype=torch.float32)
    assert torch.equal(ei, c['outputs']['edge_index']), name
    assert torch.equal(ew, c['outputs']['edge_weight']), name
  for name in _names(golden, 'gcn_norm_fill_val'):
    c = golden[name]
    i = c['inputs']
    ei, ew = O.gcn_norm_fill_val(i['edge_index'], i['edge_weight'], i['fill_value'], i['num_nodes'],
                                 dtype=torch.float32)
    assert torch.equal(ei, c['outputs']['edge_index']), name
    assert torch.equal(ew, c['outputs']['edge_weight']), name


def test_attention_layer(golden):
  for name in _names(golden, 'attention_layer'):
    c = golden[name]
    att, prods = C.oracle_attention_layer(c)
    C.assert_close(prods, c['outputs']['prods'], ORACLE_RTOL, name + ' prods')
    C.assert_close(att, c['outputs']['attention'], ORACLE_RTOL, name + ' attention')
    # fp64 arbiter agrees with the fp32 reference to fp32 rounding
    att64, _ = C.oracle_attention_layer(c, torch.float64)
    C.assert_close(att64, c['outputs']['attention'], 1e-5, name + ' attention(fp64)')


def test_K3_is_exactly_half(golden):
  c = golden['att_layer/K3_symmetric']
  att, _ = C.oracle_attention_layer(c)
  assert torch.equal(att, 0.5 * torch.ones(6, 2))
  assert torch.equal(c['outputs']['attention'], 0.5 * torch.ones(6, 2))


def test_blocks(golden):
  for name in _names(golden, 'block'):
    c = golden[name]
    rhs, z, nfe, (ei, ew) = C.oracle_block(c)
    assert torch.equal(ei, c['outputs']['odefunc_edge_index']), name
    C.assert_close(ew, c['outputs']['odefunc_edge_weight'], 1e-7, name + ' edge_weight')
    C.assert_close(rhs, c['outputs']['rhs'], ORACLE_RTOL, name + ' rhs')
    assert nfe == c['outputs']['nfe'], '%s nfe %d vs %d' % (name, nfe, c['outputs']['nfe'])
    C.assert_close(z, c['outputs']['z'], 1e-5, name + ' z')


def test_hard_attention_edge_selection(golden):
  """HardAttODEblock (reference src/block_transformer_hard_attention.py:49-68): the kept edge set is reproduced
  exactly, the renormalised weights to rounding; training keeps ~att_samp_pct of the edges, eval keeps all."""
  names = [n for n in _names(golden, 'block') if golden[n]['opt']['block'] == 'hard_attention']
  assert len(names) == 6
  for name in names:
    c = golden[name]
    _, _, _, (ei, _), extra = C.oracle_block(c, want_extra=True)
    assert torch.equal(ei, c['outputs']['odefunc_edge_index']), name
    C.assert_close(extra['attention_weights'], c['outputs']['attention_weights'], ORACLE_RTOL, name + ' weights')
    kept = ei.shape[1] / c['outputs']['odefunc_edge_weight'].shape[0]
    assert (abs(kept - c['opt']['att_samp_pct']) < 0.02) if c['inputs']['train'] else kept == 1.0, name


def test_blocks_fp64_arbiter(golden):
  for name in _names(golden, 'block'):
    c = golden[name]
    if c['opt']['method'] == 'dopri5':
      continue  # accepted-step sequence may legitimately differ at fp64; covered by the fp32 replay
    rhs, z, nfe, _ = C.oracle_block(c, torch.float64)
    C.assert_close(rhs, c['outputs']['rhs'], 1e-5, name + ' rhs(fp64)')
    C.assert_close(z, c['outputs']['z'], 2e-5, name + ' z(fp64)')
