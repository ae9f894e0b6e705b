"""Where does the edge kernel's time go?  Times the edge phase (grand_rhs_attention_edges) and the pure weighted
aggregation (grand_rhs_laplacian) from CUDA graphs on Chung-Lu graphs of the ogbn-arxiv shape while varying ONE thing
at a time:

  scale    N and E divided by 1 / 2 / 4 / 8 (average degree kept): once X and the streams fit the 126 MB L2 the
           gathers stop missing to DRAM -- the change in edges/s is the cost of those misses
  width    d = 128 vs 64 at full scale: bytes per edge halve, instructions per edge stay -- bandwidth- vs
           latency/issue-bound
  epilogue f_out only vs the fused rk4 stage combine (the average of the four stage epilogues)
  mode     attention (scores + softmax + aggregation) vs given weights (aggregation only)

    python tools/edge_probe.py            (GPU box; prints one line per configuration, JSON at the end)
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from graph_neural_pde_b200 import _native as N  # noqa: E402
from graph_neural_pde_b200.block_constant import ConstantODEblock  # noqa: E402
from graph_neural_pde_b200.function_laplacian_diffusion import LaplacianODEFunc  # noqa: E402
from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt  # noqa: E402
from graph_neural_pde_b200.odeint import NativeProblem  # noqa: E402
from graph_neural_pde_b200.synthetic import Data, baseline_opt, make_graph  # noqa: E402

DEV = torch.device('cuda:0')


def timed_graph(body, launches):
  body(None)
  torch.cuda.synchronize()
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    body(N.current_stream(DEV))
  graph.replay()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  graph.replay()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) * 1e3 / launches


def probe(scale, d, seed=0):
  n, pairs, max_deg = 169343 // scale, 1157799 // scale, max(64, 13161 // scale)
  edge_index = make_graph(n, pairs, max_degree=max_deg, seed=seed)
  opt = dict(baseline_opt('C5'), hidden_dim=d)
  g = torch.Generator().manual_seed(seed)
  x_host = torch.randn(n, d, generator=g)
  data = Data(x=x_host, edge_index=edge_index, edge_attr=None, num_nodes=n)
  out = {}
  for mode, func_cls in (('attention', ODEFuncTransformerAtt), ('weights', LaplacianODEFunc)):
    o = dict(opt, function='transformer' if mode == 'attention' else 'laplacian')
    block = ConstantODEblock(func_cls, (), o, data, DEV, t=torch.tensor([0, o['time']])).to(DEV)
    with torch.no_grad():
      for p in block.parameters():
        if p.dim() == 2:
          p.copy_((torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5).to(p.device))
    block.eval()
    func = block.odefunc
    x = x_host.to(DEV)
    block.set_x0(x)
    prob = NativeProblem(func, x)
    eng = prob.engine
    e = int(func.edge_index.shape[1])
    ws = eng.workspace(prob.spec)
    bufs = [eng.buffer(k) for k in ('y', 'k1', 'k2', 'k3', 'ya', 'yb')]
    y, k1, k2, k3, ya, yb = bufs
    y.copy_(x)
    for b in (k1, k2, k3, ya, yb):
      b.copy_(x)
    h, third = 1.0, 1.0 / 3.0
    stage_eps = [
      (y, dict(f_out=k1, y_out=ya, coef_f=h * third, terms=[(y, 1.0)])),
      (ya, dict(f_out=k2, y_out=yb, coef_f=h, terms=[(y, 1.0), (k1, -h * third)])),
      (yb, dict(f_out=k3, y_out=ya, coef_f=h, terms=[(y, 1.0), (k1, h), (k2, -h)])),
      (ya, dict(y_out=yb, coef_f=h * 0.125, terms=[(y, 1.0), (k1, h * 0.125), (k2, 3 * h * 0.125),
                                                   (k3, 3 * h * 0.125)])),
    ]
    variants = {
      'rk4': [(src, prob.epilogue(**kw)) for src, kw in stage_eps],
      'f_only': [(y, prob.epilogue(f_out=k1))],
    }
    lib = N.lib()
    gs = C.byref(eng.plan.struct())
    if mode == 'attention':
      as_ = C.byref(prob.spec.struct)
      N.check(lib.grand_project_qk(y.data_ptr(), n, n, as_, ws.q, ws.k, N.current_stream(DEV)))
    for vname, eps in variants.items():
      reps = 8 // len(eps) * 2

      def body(st, eps=eps, reps=reps):
        st = N.current_stream(DEV) if st is None else st
        for _ in range(reps):
          for src, ep in eps:
            if mode == 'attention':
              N.check(lib.grand_rhs_attention_edges(gs, as_, src.data_ptr(), C.byref(ws), C.byref(ep), st))
            else:
              N.check(lib.grand_rhs_laplacian(gs, d, src.data_ptr(), prob.w_csr.data_ptr(), C.byref(ws), C.byref(ep),
                                              st))

      us = timed_graph(body, reps * len(eps))
      out['%s/%s' % (mode, vname)] = us
      print('scale 1/%d  N %7d  E %8d  d %3d  %-9s %-6s  %8.1f us   %6.2f Gedge/s   gather %.2f TB/s' % (
        scale, n, e, d, mode, vname, us, e / us / 1e3, e * d * 4 / us / 1e6), flush=True)
    out['n'], out['e'] = n, e
    del block, prob, eng
    torch.cuda.empty_cache()
  return out


def main():
  results = {}
  for scale, d in ((1, 128), (2, 128), (4, 128), (8, 128), (1, 64), (1, 256)):
    results['scale%d_d%d' % (scale, d)] = probe(scale, d)
  print(json.dumps(results))


if __name__ == '__main__':
  main()
