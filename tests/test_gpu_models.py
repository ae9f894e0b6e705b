"""Whole-model golden and a training step of the as-shipped hard-attention + adjoint configuration.  Written at the end
of round 1 without GPU access (xfail-guarded then); both passed on the driver's B200 run and in every round-2 run, so the
guard is gone: they are ordinary GPU tests now."""
import copy

import pytest
import torch

import _cases as C
from __graft_entry__ import load_package

pytestmark = [pytest.mark.gpu]

load_package()
from graph_neural_pde_b200.models import GNN  # noqa: E402
from graph_neural_pde_b200.synthetic import Data  # noqa: E402
from graph_neural_pde_b200.utils import DummyDataset  # noqa: E402

DEV = torch.device('cuda:0')


def test_whole_model_golden(golden):
  """GNN (encoder -> CUDA block -> decoder) against the logits the reference's GNN produced (golden_v3): every shell
  option, state widths 12 / 17 (not a multiple of 4: scalar-lane kernels) / 24, both functions."""
  names = sorted(k for k, v in golden.items() if v['kind'] == 'model_shell')
  assert len(names) == 8
  for name in names:
    c = golden[name]
    opt, inp, out = copy.deepcopy(c['opt']), c['inputs'], c['outputs']
    if opt['use_labels']:
      opt['hidden_dim'] -= inp['num_classes']
    data = Data(x=inp['x'][:, :inp['num_features']], edge_index=inp['edge_index'], edge_attr=None,
                num_nodes=inp['num_nodes'])
    model = GNN(opt, DummyDataset(data, inp['num_classes']), DEV).to(DEV)
    model.load_state_dict(c['state_dict'], strict=True)
    model.eval()
    pos = inp['pos_encoding'].to(DEV) if inp['pos_encoding'] is not None else None
    with torch.no_grad():
      logits = model(inp['x'].to(DEV), pos)
    C.assert_close(logits, out['logits'], 2 * C.RTOL, name + ' logits')


def test_training_step_under_hard_attention_with_adjoint():
  """The configuration the reference ships for Photo / ogbn-arxiv (src/best_params.py): GRAND-l under the hard-attention
  block, dopri5 forward, adjoint backward with rk4 -- two optimiser steps on a small synthetic graph: finite loss that
  moves, finite non-zero gradients on the ODE scalars and both linear layers, fewer edges kept than given."""
  from graph_neural_pde_b200.synthetic import SyntheticDataset, default_opt
  torch.manual_seed(0)
  ds = SyntheticDataset('Cora', seed=0, num_features=32)
  opt = default_opt(dataset='Cora', function='laplacian', block='hard_attention', method='dopri5', adjoint=True,
                    adjoint_method='rk4', adjoint_step_size=1, step_size=1, time=2.0, tol_scale=100.0,
                    tol_scale_adjoint=100.0, hidden_dim=32, attention_dim=16, heads=4, att_samp_pct=0.8,
                    input_dropout=0.0, no_early=True, max_nfe=10000)
  model = GNN(opt, ds, DEV).to(DEV)
  data = ds.data.to(DEV)
  optim = torch.optim.Adam(model.parameters(), lr=0.02)
  losses = []
  for _ in range(2):
    model.train()
    optim.zero_grad()
    out = model(data.x)
    loss = torch.nn.functional.cross_entropy(out[data.train_mask], data.y.squeeze()[data.train_mask])
    loss.backward()
    for name in ('odeblock.odefunc.alpha_train', 'm1.weight', 'm2.weight'):
      g = dict(model.named_parameters())[name].grad
      assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0, name
    optim.step()
    losses.append(float(loss))
  assert all(torch.isfinite(torch.tensor(losses))) and losses[1] != losses[0]
  kept = model.odeblock.odefunc.edge_index.shape[1]
  assert kept < model.odeblock.data_edge_index.shape[1] and model.odeblock.odefunc.attention_weights.shape[0] == kept
