"""Third batch of golden vectors: the model shell either side of the ODE block -- encoder
(m1 | mx + mp, use_mlp residual layers, labels, batch norm, augmentation) and decoder (relu, fc_out, m2) of the
reference's GNN (src/GNN.py:17-77, src/base_classes.py:98-131), eval mode, recorded from the reference's own module.
For every case the block's input and output are captured as well, so the shell can be checked without running a block.

Run in the build container only:   python tests/golden/make_golden_v3.py   ->  tests/golden/golden_v3.pt
"""
import copy
import os
import warnings

import torch

import make_golden as G

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
  G.install_reference_on_path()
  warnings.simplefilter('ignore')
  from test_params import OPT
  from GNN import GNN
  from utils import DummyDataset

  base = {**OPT, 'mix_features': False, 'attention_dim': 16, 'hidden_dim': 12, 'heads': 4, 'max_nfe': 100000,
          'block': 'constant', 'function': 'transformer', 'time': 1.0, 'step_size': 0.5, 'method': 'rk4',
          'no_early': True, 'dataset': 'Cora', 'dropout': 0.3, 'input_dropout': 0.4, 'use_flux': False,
          'att_samp_pct': 1, 'pos_enc_dim': 7, 'feat_hidden_dim': 8, 'pos_enc_hidden_dim': 4}
  device = torch.device('cpu')
  n, classes = 48, 5
  edge_index, _ = G.make_graph(n, 110, seed=1)
  g = torch.Generator().manual_seed(31)
  feats = torch.randn(n, 20, generator=g)
  pos = torch.randn(n, 7, generator=g)
  onehot = torch.nn.functional.one_hot(torch.randint(0, classes, (n,), generator=g), classes).float()

  class DataWithMasks(G.Data):
    def __call__(self, *keys):
      for k in keys:
        yield k, getattr(self, k)

  variants = {
    'plain': {},
    'use_mlp': {'use_mlp': True},
    'fc_out': {'fc_out': True},
    'batch_norm': {'batch_norm': True},
    'augment': {'augment': True},
    'use_labels': {'use_labels': True},
    'beltrami': {'beltrami': True},
    'everything': {'use_mlp': True, 'fc_out': True, 'batch_norm': True, 'augment': True, 'function': 'laplacian'},
  }
  cases = {}
  seed = 900
  for tag, extra in variants.items():
    opt = {**base, **extra}
    x_in = torch.cat([feats, onehot], dim=1) if opt['use_labels'] else feats
    data = DataWithMasks(x_in[:, :20], edge_index, None, n)
    dataset = DummyDataset(data, classes)
    seed += 1
    torch.manual_seed(seed)
    model = GNN(opt, dataset, device)      # rewrites opt['hidden_dim'] in place exactly like the reference
    G.randomise_(model, seed=seed)
    if opt['batch_norm']:                  # non-trivial running statistics, as after training
      gg = torch.Generator().manual_seed(seed)
      for bn in (model.bn_in, model.bn_out):
        bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=gg) * 0.1)
        bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=gg) + 0.5)
    model.eval()
    captured = {}
    inner = model.odeblock.forward

    def spy(x, inner=inner, captured=captured):
      captured['block_in'] = x.detach().clone()
      z = inner(x)
      captured['block_out'] = z.detach().clone()
      return z

    model.odeblock.forward = spy
    with torch.no_grad():
      logits = model(x_in, pos if opt['beltrami'] else None)
    cases['model_shell/%s' % tag] = dict(
      kind='model_shell', opt=copy.deepcopy(opt),
      inputs=dict(x=x_in, pos_encoding=pos if opt['beltrami'] else None, edge_index=edge_index, num_nodes=n,
                  num_classes=classes, num_features=20),
      state_dict=G.sd(model),
      outputs=dict(logits=logits, block_in=captured['block_in'], block_out=captured['block_out']))
    print('%-26s block width %3d  logits %s' % (tag, captured['block_in'].shape[1], tuple(logits.shape)))
  out = os.path.join(HERE, 'golden_v3.pt')
  torch.save(cases, out)
  print('wrote %d cases -> %s (%.1f KiB)' % (len(cases), out, os.path.getsize(out) / 1024))


if __name__ == '__main__':
  main()
