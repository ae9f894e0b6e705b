"""Second batch of golden vectors, again produced by the reference's OWN modules (see make_golden.py for the shim
arrangement): the rows SURVEY.md 8(a5) / 8(f).4 name next to the default path --

  * `function=GAT`                additive scorer, leaky ReLU     (src/function_GAT_attention.py:11-116)
  * `block=hard_attention`        quantile-thresholded edge set   (src/block_transformer_hard_attention.py:7-104)

Run in the build container only:   python tests/golden/make_golden_v2.py   ->  tests/golden/golden_v2.pt
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
  from function_transformer_attention import ODEFuncTransformerAtt
  from function_GAT_attention import ODEFuncAtt
  from function_laplacian_diffusion import LaplacianODEFunc
  from block_constant import ConstantODEblock
  from block_transformer_hard_attention import HardAttODEblock

  base = {**OPT, 'mix_features': False, 'attention_dim': 16, 'hidden_dim': 12, 'heads': 4, 'max_nfe': 100000,
          'block': 'constant', 'function': 'transformer', 'earlystopxT': 3, 'time': 1.6, 'tol_scale': 100.0,
          'use_flux': False, 'att_samp_pct': 1}
  device = torch.device('cpu')
  cases = {}
  n = 48
  edge_index, _ = G.make_graph(n, 110, seed=1)
  edge_index_w, edge_attr_w = G.make_graph(n, 110, seed=2, weighted=True, with_existing_loops=5)
  x12 = torch.randn(n, 12, generator=torch.Generator().manual_seed(7))
  x32 = torch.randn(n, 32, generator=torch.Generator().manual_seed(9))
  funcs = {'transformer': ODEFuncTransformerAtt, 'GAT': ODEFuncAtt, 'laplacian': LaplacianODEFunc}

  def run_block(tag, opt, ei, ew, x, train_mode, seed):
    data = G.Data(x, ei, ew, n)
    block_cls = HardAttODEblock if opt['block'] == 'hard_attention' else ConstantODEblock
    block = block_cls(funcs[opt['function']], (), opt, data, device, t=torch.tensor([0, opt['time']]))
    G.randomise_(block, seed=seed)
    if opt['function'] == 'GAT':   # `a` is [2 d_k, 1, 1]: give it a non-trivial value (randomise_ skips 3-D tensors)
      g = torch.Generator().manual_seed(seed + 1)
      with torch.no_grad():
        block.odefunc.multihead_att_layer.a.copy_(torch.randn(block.odefunc.multihead_att_layer.a.shape, generator=g))
    block.train(train_mode)
    block.set_x0(x.clone())
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
      z = block(x)                                     # hard attention: sets odefunc.edge_index / attention_weights
      nfe = block.odefunc.nfe
      f = block.odefunc(torch.tensor(0.0), x)          # one RHS evaluation on the graph the forward left behind
    out = dict(rhs=f, z=z, nfe=nfe, odefunc_edge_index=block.odefunc.edge_index.clone(),
               odefunc_edge_weight=None if block.odefunc.edge_weight is None else block.odefunc.edge_weight.clone())
    if opt['block'] == 'hard_attention':
      out['attention_weights'] = block.odefunc.attention_weights.clone()
    cases[tag] = dict(kind='block', opt=copy.deepcopy(opt),
                      inputs=dict(x=x, edge_index=ei, edge_attr=ew, num_nodes=n, train=train_mode),
                      state_dict=G.sd(block), outputs=out)

  seed = 500
  # mix_features=True is NOT generated: the reference itself raises on that path (the layer returns
  # `attention, (v, prods)` and ODEFuncTransformerAtt.forward hands the whole tuple to multiply_attention, whose
  # `v.shape` fails -- src/function_transformer_attention.py:43-44, :29, :214), so there is no reference behaviour to pin.
  # --- GAT scorer ---------------------------------------------------------------------------------------------------
  for variant in range(3):
    opt = {**base, 'function': 'GAT', 'method': 'rk4', 'step_size': 0.5}
    x, ei, ew = x12, edge_index, None
    if variant == 1:
      opt.update(add_source=True, attention_norm_idx=1, leaky_relu_slope=0.05)
    if variant == 2:
      opt.update(hidden_dim=32, attention_dim=32, heads=8, add_source=True, no_alpha_sigmoid=True, method='dopri5',
                 step_size=1)
      x, ei, ew = x32, edge_index_w, edge_attr_w
    seed += 1
    run_block('block/GAT_constant/v%d' % variant, opt, ei, ew, x, False, seed)

  # --- hard attention block ------------------------------------------------------------------------------------------
  for function in ['laplacian', 'transformer']:
    for train_mode in [True, False]:
      for flux in [False, True]:
        if not train_mode and flux:
          continue
        opt = {**base, 'function': function, 'block': 'hard_attention', 'method': 'rk4', 'step_size': 0.5,
               'att_samp_pct': 0.6, 'use_flux': flux, 'add_source': function == 'laplacian'}
        seed += 1
        run_block('block/%s_hard_attention/train%d_flux%d' % (function, train_mode, flux), opt, edge_index, None, x12,
                  train_mode, seed)

  out = os.path.join(HERE, 'golden_v2.pt')
  torch.save(cases, out)
  print('wrote %d cases -> %s (%.1f KiB)' % (len(cases), out, os.path.getsize(out) / 1024))
  for k, c in cases.items():
    print('  %-55s nfe %3d  |z| %.4f  E %d' % (k, c['outputs']['nfe'], float(c['outputs']['z'].abs().max()),
                                              c['outputs']['odefunc_edge_index'].shape[1]))


if __name__ == '__main__':
  main()
