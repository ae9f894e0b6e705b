"""Fourth batch of golden vectors, produced by the reference's OWN modules (see make_golden.py for the shim arrangement):
`function=GAT` with `mix_features=True` -- aggregate x W with every head's attention, average, then multiply by Wout
(src/function_GAT_attention.py:31-37), the one mix_features path of the reference that runs (the transformer function's
raises, see make_golden_v2.py).

Run in the build container only:   python tests/golden/make_golden_v4.py   ->  tests/golden/golden_v4.pt
"""
import contextlib
import copy
import io
import os
import warnings

import torch

import make_golden as G

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
  G.install_reference_on_path()
  warnings.simplefilter('ignore')
  from test_params import OPT
  from function_GAT_attention import ODEFuncAtt
  from block_constant import ConstantODEblock

  base = {**OPT, 'mix_features': True, 'attention_dim': 16, 'hidden_dim': 12, 'heads': 4, 'max_nfe': 100000,
          'block': 'constant', 'function': 'GAT', 'earlystopxT': 3, 'time': 1.6, 'tol_scale': 100.0}
  device = torch.device('cpu')
  cases = {}
  n = 48
  edge_index, _ = G.make_graph(n, 110, seed=1)
  edge_index_w, edge_attr_w = G.make_graph(n, 110, seed=2, weighted=True, with_existing_loops=5)
  x12 = torch.randn(n, 12, generator=torch.Generator().manual_seed(7))
  x32 = torch.randn(n, 32, generator=torch.Generator().manual_seed(9))
  seed = 900
  for variant in range(3):
    opt = {**base, 'method': 'rk4', 'step_size': 0.5}
    x, ei, ew = x12, edge_index, None
    if variant == 1:
      opt.update(add_source=True, attention_norm_idx=1, leaky_relu_slope=0.05, method='euler', step_size=0.4)
    if variant == 2:
      # (fixed grid on purpose: for an adaptive case tests/test_gpu_parity.py needs a tight fp64 oracle solution of this
      # 8-head per-head-spmm function, a minute of CPU per run)
      opt.update(hidden_dim=32, attention_dim=32, heads=8, add_source=True, no_alpha_sigmoid=True, method='midpoint',
                 step_size=0.4)
      x, ei, ew = x32, edge_index_w, edge_attr_w
    seed += 1
    data = G.Data(x, ei, ew, n)
    block = ConstantODEblock(ODEFuncAtt, (), opt, data, device, t=torch.tensor([0, opt['time']]))
    G.randomise_(block, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
      layer = block.odefunc.multihead_att_layer
      layer.a.copy_(torch.randn(layer.a.shape, generator=g))
      layer.Wout.copy_(torch.randn(layer.Wout.shape, generator=g) / layer.Wout.shape[0] ** 0.5)
    block.eval()
    block.set_x0(x.clone())
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
      z = block(x)
      nfe = block.odefunc.nfe
      f = block.odefunc(torch.tensor(0.0), x)
    cases['block/GAT_mix_features/v%d' % variant] = dict(
      kind='block', opt=copy.deepcopy(opt), inputs=dict(x=x, edge_index=ei, edge_attr=ew, num_nodes=n, train=False),
      state_dict=G.sd(block),
      outputs=dict(rhs=f, z=z, nfe=nfe, odefunc_edge_index=block.odefunc.edge_index.clone(),
                   odefunc_edge_weight=None if block.odefunc.edge_weight is None else block.odefunc.edge_weight.clone()))
  out = os.path.join(HERE, 'golden_v4.pt')
  torch.save(cases, out)
  print('wrote %d cases -> %s (%.1f KiB)' % (len(cases), out, os.path.getsize(out) / 1024))
  for k, c in cases.items():
    print('  %-40s nfe %3d  |z| %.4f  E %d' % (k, c['outputs']['nfe'], float(c['outputs']['z'].abs().max()),
                                             c['outputs']['odefunc_edge_index'].shape[1]))


if __name__ == '__main__':
  main()
