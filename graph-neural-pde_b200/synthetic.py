"""Synthetic stand-ins for the reference's datasets (there is no network and no dataset on disk):
undirected simple Chung-Lu graphs with power-law expected degrees (exponent 2.5) and the node / edge / max-degree
counts of the real graphs, symmetrised to directed COO; self loops are NOT included here -- the blocks add them
(get_rw_adj appends them at the end of the list, PyG order).  Deterministic in `seed`.

Also the reference's option dictionary with CLI defaults (reference src/run_GNN.py:282-436) for building `opt`
directly, which is the only way to reach the BASELINE configs.
"""
import os

import numpy as np
import torch

# name -> (nodes, undirected pairs, max degree, features, classes); SURVEY.md Appendix B
DATASETS = {
  'Cora': (2485, 5069, 168, 1433, 7),
  'Photo': (7487, 119043, 1434, 745, 8),
  'ogbn-arxiv': (169343, 1157799, 13161, 128, 40),
}


def make_graph(n, m_pairs, max_degree=None, seed=0, exponent=2.5):
  """-> edge_index int64 [2, 2 * m_pairs], both directions of every pair, in shuffled order."""
  rng = np.random.default_rng(seed)
  max_degree = max_degree or max(4, int(4 * m_pairs / n))
  ranks = np.arange(n, dtype=np.float64)
  power = -1.0 / (exponent - 1.0)

  def weights(i0):
    w = (ranks + i0) ** power
    return w / w.sum()

  lo, hi = 1e-3, float(n)
  for _ in range(60):   # expected max degree 2 m w_0 == max_degree
    mid = np.sqrt(lo * hi)
    if 2.0 * m_pairs * weights(mid)[0] > max_degree:
      lo = mid
    else:
      hi = mid
  w = weights(hi)
  w = w[rng.permutation(n)]      # hubs get arbitrary ids, like a real dataset
  cdf = np.cumsum(w)
  cdf[-1] = 1.0
  keys = np.empty(0, dtype=np.int64)
  want = m_pairs
  while len(keys) < want:
    draw = int((want - len(keys)) * 1.3) + 1024
    a = np.searchsorted(cdf, rng.random(draw), side='right')
    b = np.searchsorted(cdf, rng.random(draw), side='right')
    ok = a != b
    lo_, hi_ = np.minimum(a[ok], b[ok]), np.maximum(a[ok], b[ok])
    keys = np.unique(np.concatenate([keys, lo_.astype(np.int64) * n + hi_]))
  keys = keys[rng.permutation(len(keys))[:want]]
  a, b = keys // n, keys % n
  src = np.concatenate([a, b])
  dst = np.concatenate([b, a])
  order = rng.permutation(len(src))
  return torch.from_numpy(np.stack([src[order], dst[order]])).long()


class Data:
  """Minimal stand-in for torch_geometric.data.Data as the blocks use it."""

  def __init__(self, x=None, edge_index=None, edge_attr=None, num_nodes=None, y=None, **masks):
    self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
    self.num_nodes = num_nodes if num_nodes is not None else (x.shape[0] if x is not None else None)
    for k, v in masks.items():
      setattr(self, k, v)

  @property
  def num_features(self):
    return self.x.shape[1]

  def to(self, device):
    for k, v in list(self.__dict__.items()):
      if torch.is_tensor(v):
        setattr(self, k, v.to(device))
    return self

  def __call__(self, *keys):
    for k in keys:
      yield k, getattr(self, k)


class SyntheticDataset:
  """`dataset.data`, `dataset.num_classes`, `dataset.num_features` like a PyG dataset."""

  def __init__(self, name, seed=0, num_features=None, root=None):
    n, m, maxdeg, feats, classes = DATASETS[name]
    feats = num_features or feats
    g = torch.Generator().manual_seed(seed)
    self.source = 'synthetic'
    edge_index = None
    # real graph, if somebody put one next to the code: <root>/<name>/edge_index.npy, int64 [2, E0],
    # the dataset's directed edge list before self loops (what PyG's loader hands to the reference); nothing is downloaded
    root = root or os.environ.get('GRAND_DATA_ROOT') or os.path.join(os.path.dirname(os.path.dirname(
      os.path.abspath(__file__))), 'data')
    path = os.path.join(root, name, 'edge_index.npy')
    if os.path.exists(path):
      ei = np.load(path)
      if ei.ndim == 2 and ei.shape[0] == 2 and ei.size > 0:
        edge_index = torch.from_numpy(np.ascontiguousarray(ei)).long()
        n = int(edge_index.max()) + 1 if int(edge_index.max()) + 1 > n else n
        self.source = path
    if edge_index is None:
      edge_index = make_graph(n, m, maxdeg, seed)
    x = torch.randn(n, feats, generator=g)
    y = torch.randint(0, classes, (n,), generator=g)
    perm = torch.randperm(n, generator=g)
    cut1, cut2 = int(0.1 * n), int(0.2 * n)
    masks = {}
    for key, idx in (('train_mask', perm[:cut1]), ('val_mask', perm[cut1:cut2]), ('test_mask', perm[cut2:])):
      m_ = torch.zeros(n, dtype=torch.bool)
      m_[idx] = True
      masks[key] = m_
    if name == 'ogbn-arxiv':
      y = y.unsqueeze(1)
    self.name = name
    self.data = Data(x=x, edge_index=edge_index, edge_attr=None, num_nodes=n, y=y, **masks)
    self.num_classes = classes
    self.num_features = feats


def default_opt(**overrides):
  """CLI defaults of reference src/run_GNN.py:282-436 (only the keys the hot path and its shells read)."""
  opt = {
    'use_cora_defaults': False, 'dataset': 'Cora', 'data_norm': 'rw', 'self_loop_weight': 1.0, 'use_labels': False,
    'geom_gcn_splits': False, 'num_splits': 1, 'label_rate': 0.5, 'planetoid_split': False,
    'hidden_dim': 16, 'fc_out': False, 'input_dropout': 0.5, 'dropout': 0.0, 'batch_norm': False,
    'optimizer': 'adam', 'lr': 0.01, 'decay': 5e-4, 'epoch': 100, 'alpha': 1.0, 'alpha_dim': 'sc',
    'no_alpha_sigmoid': False, 'beta_dim': 'sc', 'block': 'constant', 'function': 'laplacian', 'use_mlp': False,
    'add_source': False, 'cgnn': False,
    'time': 1.0, 'augment': False, 'method': None, 'step_size': 1, 'max_iters': 100,
    'adjoint_method': 'adaptive_heun', 'adjoint': False, 'adjoint_step_size': 1, 'tol_scale': 1.,
    'tol_scale_adjoint': 1.0, 'ode_blocks': 1, 'max_nfe': 1000, 'no_early': False, 'earlystopxT': 3,
    'max_test_steps': 100,
    'leaky_relu_slope': 0.2, 'attention_dropout': 0., 'heads': 4, 'attention_norm_idx': 0, 'attention_dim': 64,
    'mix_features': False, 'reweight_attention': False, 'attention_type': 'scaled_dot', 'square_plus': False,
    'jacobian_norm2': None, 'total_deriv': None, 'kinetic_energy': None, 'directional_penalty': None,
    'not_lcc': True, 'rewiring': None, 'rewire_KNN': False, 'fa_layer': False, 'beltrami': False,
    'feat_hidden_dim': 64, 'pos_enc_hidden_dim': 32, 'pos_enc_type': 'DW64', 'edge_sampling': False, 'gpu': 0,
    'att_samp_pct': 1, 'use_flux': False,
  }
  opt.update(overrides)
  return opt


# BASELINE.json configs as opt dictionaries
def baseline_opt(name):
  if name == 'C1':   # Cora GRAND-l, rk4, T=1, h=0.1
    return default_opt(dataset='Cora', function='laplacian', block='constant', hidden_dim=80, method='rk4',
                       time=1.0, step_size=0.1)
  if name == 'C2':   # Cora GRAND-nl, 4 heads, rk4
    return default_opt(dataset='Cora', function='transformer', block='constant', hidden_dim=80, attention_dim=64,
                       heads=4, method='rk4', time=1.0, step_size=0.1)
  if name == 'C3':   # Photo GRAND-nl, 8 heads, dopri5
    return default_opt(dataset='Photo', function='transformer', block='constant', hidden_dim=64, attention_dim=64,
                       heads=8, method='dopri5', time=3.5824, tol_scale=2086.5, step_size=1)
  if name == 'C4':   # Cora BLEND: joint feature + positional state, 64 + 16
    return default_opt(dataset='Cora', function='transformer', block='constant', beltrami=True, feat_hidden_dim=64,
                       pos_enc_hidden_dim=16, hidden_dim=80, attention_dim=64, heads=4, method='rk4', time=1.0,
                       step_size=0.1)
  if name == 'C5':   # ogbn-arxiv GRAND-nl (reference src/ray_tune.py:19-22)
    return default_opt(dataset='ogbn-arxiv', function='transformer', block='constant', hidden_dim=128,
                       attention_dim=16, heads=4, method='rk4', time=5.08, step_size=1, add_source=True,
                       adjoint=False, max_nfe=100000)
  raise KeyError(name)
