"""Generate golden vectors by running the reference's OWN modules (unmodified, from /root/reference/src).

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py

The reference's in-tree code (function_transformer_attention.py, function_laplacian_diffusion.py,
block_constant.py, block_transformer_attention.py, base_classes.py, utils.py, early_stop_solver.py, GNN*.py)
is imported as-is.  Its absent third-party leaves (torch_geometric / torch_scatter / torch_sparse / torchdiffeq /
ogb) are provided by `oracle/shims` (plain-PyTorch restatements of those packages' published algorithms) and
the dataset module `data` (PyG dataset downloads) by an empty stub, since the generator feeds synthetic graphs.

Output: tests/golden/golden_v1.pt -- {case name: {opt, inputs, state_dict, outputs}}; every tensor fp32/int64,
small enough to commit.  tests/test_golden.py replays every case against the oracle (CPU) and, under -m gpu,
against the CUDA path.
"""
import copy
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
REF_SRC = '/root/reference/src'
REF_TEST = '/root/reference/test'


def install_reference_on_path():
  assert os.path.isdir(REF_SRC), 'the golden generator needs /root/reference (build container only)'
  sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shims'))
  sys.path.insert(1, REF_SRC)
  sys.path.insert(2, REF_TEST)
  stub = types.ModuleType('data')
  stub.get_dataset = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('no datasets in this image'))
  sys.modules['data'] = stub


def make_graph(n, m_pairs, seed, weighted=False, with_existing_loops=0):
  """Random undirected simple graph -> directed COO [2, 2*m] in shuffled order (+ a few self loops)."""
  rng = np.random.default_rng(seed)
  pairs = set()
  while len(pairs) < m_pairs:
    a, b = rng.integers(0, n, size=2)
    if a != b:
      pairs.add((min(a, b), max(a, b)))
  pairs = np.array(sorted(pairs))
  ei = np.concatenate([pairs, pairs[:, ::-1]], axis=0)
  if with_existing_loops:
    loops = rng.choice(n, size=with_existing_loops, replace=False)
    ei = np.concatenate([ei, np.stack([loops, loops], 1)], axis=0)
  ei = ei[rng.permutation(len(ei))]
  # make sure the max node id is present so maybe_num_nodes == n
  edge_index = torch.tensor(ei.T.copy(), dtype=torch.long)
  edge_attr = None
  if weighted:
    edge_attr = torch.tensor(rng.uniform(0.5, 2.0, size=edge_index.shape[1]), dtype=torch.float32)
  return edge_index, edge_attr


class Data:
  def __init__(self, x, edge_index, edge_attr, num_nodes):
    self.x, self.edge_index, self.edge_attr, self.num_nodes = x, edge_index, edge_attr, num_nodes
    self.num_features = x.shape[1]


def randomise_(module, seed, scale=None):
  """Trained-like parameters: N(0, 1/fan_in) weights, U(-.5,.5) biases, non-trivial scalars."""
  g = torch.Generator().manual_seed(seed)
  with torch.no_grad():
    for name, p in module.named_parameters():
      if p.dim() == 2:
        p.copy_(torch.randn(p.shape, generator=g) / np.sqrt(p.shape[1]) * (scale or 1.0))
      elif name.endswith('bias'):
        p.copy_(torch.rand(p.shape, generator=g) - 0.5)
      elif name.endswith('alpha_train'):
        p.fill_(0.3)
      elif name.endswith('beta_train'):
        p.fill_(-0.7)
      elif 'lengthscale' in name:
        p.fill_(1.7)
      elif 'output_var' in name:
        p.fill_(0.8)


def sd(module):
  return {k: v.detach().clone() for k, v in module.state_dict().items()}


def main():
  install_reference_on_path()
  warnings.simplefilter('ignore')
  from test_params import OPT
  from function_transformer_attention import SpGraphTransAttentionLayer, ODEFuncTransformerAtt
  from function_laplacian_diffusion import LaplacianODEFunc
  from block_constant import ConstantODEblock
  from block_transformer_attention import AttODEblock
  from utils import get_rw_adj, gcn_norm_fill_val
  from GNN import GNN
  from GNN_early import GNNEarly
  from utils import DummyDataset

  base = {**OPT, 'mix_features': False, 'attention_dim': 16, 'hidden_dim': 12, 'heads': 4, 'max_nfe': 100000,
          'block': 'constant', 'function': 'transformer', 'earlystopxT': 3}
  device = torch.device('cpu')
  cases = {}

  n = 48
  edge_index, _ = make_graph(n, 110, seed=1)
  edge_index_w, edge_attr_w = make_graph(n, 110, seed=2, weighted=True, with_existing_loops=5)
  gx = torch.Generator().manual_seed(7)
  x12 = torch.randn(n, 12, generator=gx)

  # --- 1. graph normalisation ---------------------------------------------------------------------------------
  for name, (ei, ew) in {'plain': (edge_index, None), 'weighted_loops': (edge_index_w, edge_attr_w)}.items():
    for fill in [0, 0.3, 1, 3.2]:
      for norm_dim in [0, 1]:
        oi, ow = get_rw_adj(ei, edge_weight=ew, norm_dim=norm_dim, fill_value=fill, num_nodes=n, dtype=torch.float32)
        cases['rw_adj/%s/fill%s/dim%d' % (name, fill, norm_dim)] = dict(
          kind='get_rw_adj', inputs=dict(edge_index=ei, edge_weight=ew, norm_dim=norm_dim, fill_value=fill, num_nodes=n),
          outputs=dict(edge_index=oi, edge_weight=ow))
      oi, ow = gcn_norm_fill_val(ei, edge_weight=ew, fill_value=fill, num_nodes=n, dtype=torch.float32)
      cases['gcn_norm/%s/fill%s' % (name, fill)] = dict(
        kind='gcn_norm_fill_val', inputs=dict(edge_index=ei, edge_weight=ew, fill_value=fill, num_nodes=n),
        outputs=dict(edge_index=oi, edge_weight=ow))

  # --- 2. attention layer: every score type x normaliser x norm index x reweight ------------------------------
  rw_index, rw_weight = get_rw_adj(edge_index, edge_weight=None, norm_dim=1, fill_value=1, num_nodes=n,
                                   dtype=torch.float32)
  for att_type in ['scaled_dot', 'cosine_sim', 'pearson', 'exp_kernel']:
    for square_plus in [False, True]:
      for idx in [0, 1]:
        for reweight in [False, True]:
          opt = {**base, 'attention_type': att_type, 'square_plus': square_plus, 'attention_norm_idx': idx,
                 'reweight_attention': reweight}
          layer = SpGraphTransAttentionLayer(12, 12, opt, device, edge_weights=rw_weight)
          randomise_(layer, seed=11, scale=2.0)
          with torch.no_grad():
            att, (v, prods) = layer(x12, rw_index)
          cases['att_layer/%s/sp%d/idx%d/rw%d' % (att_type, square_plus, idx, reweight)] = dict(
            kind='attention_layer', opt=copy.deepcopy(opt),
            inputs=dict(x=x12, edge=rw_index, edge_weights=rw_weight), state_dict=sd(layer),
            outputs=dict(attention=att, prods=prods))

  # reference's own known answer: K3, x = ones, constant init -> attention == 0.5 (test_transformer_attention.py:92-99)
  edge1 = torch.tensor([[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]])
  x1 = torch.ones((3, 2))
  opt = {**OPT, 'attention_dim': 32, 'mix_features': False, 'heads': 2}
  layer = SpGraphTransAttentionLayer(2, 2, opt, device, concat=True)
  with torch.no_grad():
    att, (v, prods) = layer(x1, edge1)
  cases['att_layer/K3_symmetric'] = dict(kind='attention_layer', opt=copy.deepcopy(opt),
                                         inputs=dict(x=x1, edge=edge1, edge_weights=None), state_dict=sd(layer),
                                         outputs=dict(attention=att, prods=prods))

  # beltrami + exp_kernel (split feature / positional kernels)
  opt = {**base, 'beltrami': True, 'attention_type': 'exp_kernel', 'feat_hidden_dim': 8, 'pos_enc_hidden_dim': 4,
         'hidden_dim': 12}
  layer = SpGraphTransAttentionLayer(12, 12, opt, device, edge_weights=rw_weight)
  randomise_(layer, seed=12)
  with torch.no_grad():
    att, (v, prods) = layer(x12, rw_index)
  cases['att_layer/beltrami_exp_kernel'] = dict(kind='attention_layer', opt=copy.deepcopy(opt),
                                                inputs=dict(x=x12, edge=rw_index, edge_weights=rw_weight),
                                                state_dict=sd(layer), outputs=dict(attention=att, prods=prods))

  # --- 3. ODE functions inside their blocks: one RHS eval and the whole block forward -------------------------
  def run_block(tag, opt, ei, ew, x, train_mode, seed):
    data = Data(x, ei, ew, n)
    func_cls = ODEFuncTransformerAtt if opt['function'] == 'transformer' else LaplacianODEFunc
    block_cls = AttODEblock if opt['block'] == 'attention' else ConstantODEblock
    block = block_cls(func_cls, (), opt, data, device, t=torch.tensor([0, opt['time']]))
    randomise_(block, seed=seed)
    block.train(train_mode)
    x0 = x.clone()
    block.set_x0(x0)
    with torch.no_grad():
      if opt['block'] == 'attention':
        block.odefunc.attention_weights = block.get_attention_weights(x)
      f = block.odefunc(torch.tensor(0.0), x)
      nfe0 = block.odefunc.nfe
      z = block(x)
      nfe = block.odefunc.nfe - nfe0
    cases[tag] = dict(kind='block', opt=copy.deepcopy(opt),
                      inputs=dict(x=x, edge_index=ei, edge_attr=ew, num_nodes=n, train=train_mode),
                      state_dict=sd(block),
                      outputs=dict(rhs=f, z=z, nfe=nfe, odefunc_edge_index=block.odefunc.edge_index,
                                   odefunc_edge_weight=block.odefunc.edge_weight))

  seed = 100
  for method, step in [('euler', 0.5), ('midpoint', 0.5), ('rk4', 0.25), ('rk4', 1.0), ('dopri5', 1.0)]:
    for function, block in [('transformer', 'constant'), ('laplacian', 'constant'), ('laplacian', 'attention')]:
      for variant in range(3):
        opt = {**base, 'function': function, 'block': block, 'method': method, 'step_size': step, 'time': 1.6,
               'tol_scale': 100.0}
        if variant == 1:
          opt.update(add_source=True, square_plus=True, attention_norm_idx=1, data_norm='gcn')
        if variant == 2:
          opt.update(add_source=True, no_alpha_sigmoid=True, attention_type='cosine_sim', self_loop_weight=0.5)
        if variant > 0 and method in ('euler', 'midpoint'):
          continue
        ei, ew = (edge_index, None) if variant != 2 else (edge_index_w, edge_attr_w)
        seed += 1
        run_block('block/%s_%s/%s_h%s/v%d' % (function, block, method, step, variant), opt, ei, ew, x12,
                  train_mode=(variant == 0), seed=seed)

  # arxiv-shaped head layout (d_k = 4, H = 4, ad = 16, d = 32) and a wider one (H = 8, ad = 64)
  gx = torch.Generator().manual_seed(9)
  x32 = torch.randn(n, 32, generator=gx)
  for tag, extra in {'arxiv_like': dict(hidden_dim=32, attention_dim=16, heads=4, add_source=True, time=5.08,
                                         step_size=1, method='rk4'),
                     'photo_like': dict(hidden_dim=32, attention_dim=64, heads=8, time=3.5824, method='dopri5',
                                        tol_scale=2086.5, step_size=1)}.items():
    seed += 1
    run_block('block/transformer_constant/%s' % tag, {**base, **extra}, edge_index, None, x32, False, seed)

  # --- 4. whole model, eval mode, early-stop integrator (GNNEarly + EarlyStopInt/EarlyStopRK4/EarlyStopDopri5) ---
  class DataWithMasks(Data):
    def __call__(self, *keys):
      for k in keys:
        yield k, getattr(self, k)

  g = torch.Generator().manual_seed(21)
  feats = torch.randn(n, 20, generator=g)
  y = torch.randint(0, 5, (n,), generator=g)
  perm = torch.randperm(n, generator=g)
  masks = {}
  for name, sl in [('train_mask', perm[:16]), ('val_mask', perm[16:32]), ('test_mask', perm[32:])]:
    mk = torch.zeros(n, dtype=torch.bool)
    mk[sl] = True
    masks[name] = mk
  for method, step, T in [('rk4', 0.5, 1.0), ('dopri5', 1.0, 1.3)]:
    for function, block in [('transformer', 'constant'), ('laplacian', 'attention')]:
      opt = {**base, 'function': function, 'block': block, 'method': method, 'step_size': step, 'time': T,
             'tol_scale': 50.0, 'no_early': False, 'max_test_steps': 100, 'dataset': 'Cora', 'dropout': 0.0,
             'input_dropout': 0.0}
      data = DataWithMasks(feats, edge_index, None, n)
      data.y = y
      for k, v in masks.items():
        setattr(data, k, v)
      dataset = DummyDataset(data, 5)
      seed += 1
      torch.manual_seed(seed)
      model = GNNEarly(opt, dataset, device)
      randomise_(model, seed=seed)
      model.eval()
      with torch.no_grad():
        logits = model(feats)
      solver = model.odeblock.test_integrator.solver
      cases['model_early/%s_%s/%s' % (function, block, method)] = dict(
        kind='model_early', opt=copy.deepcopy(opt),
        inputs=dict(x=feats, edge_index=edge_index, y=y, num_nodes=n, num_classes=5, **masks),
        state_dict=sd(model),
        outputs=dict(logits=logits, best_val=float(solver.best_val), best_test=float(solver.best_test),
                     best_train=float(solver.best_train), best_time=float(solver.best_time),
                     nfe=int(model.odeblock.odefunc.nfe)))

  out = os.path.join(HERE, 'golden_v1.pt')
  torch.save(cases, out)
  print('wrote %d cases -> %s (%.1f KiB)' % (len(cases), out, os.path.getsize(out) / 1024))


if __name__ == '__main__':
  main()
