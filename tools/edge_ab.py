"""A/B of edge-kernel builds inside ONE process: the ogbn-arxiv-shaped problem (graph, plan, buffers) is built once, then
every variant library is loaded next to the in-tree one and its `grand_rhs_attention_edges` / `grand_rhs_laplacian`
are timed from a CUDA Graph on the same buffers (the struct layouts are shared).  Each variant's outputs are compared
with the first variant's.

    python tools/edge_ab.py name=path/to/lib.so[,ENV=VALUE...] ...        (GPU box)
    python tools/edge_ab.py --workload Cora base=graph-neural-pde_b200/libgrand_b200.so

ENV settings (e.g. GRAND_EDGE=window, GRAND_L2_HINTS=0) are applied before the variant's first call; the libraries read
them once per loaded image, so a library that appears twice is copied to a scratch name first.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import sys
import tempfile

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
from graph_neural_pde_b200.synthetic import Data  # noqa: E402
import bench  # noqa: E402

DEV = torch.device('cuda:0')


def load_variant(path, seen):
  path = os.path.abspath(path)
  if path in seen:
    tmp = tempfile.NamedTemporaryFile(suffix='.so', delete=False)
    tmp.close()
    shutil.copy(path, tmp.name)
    path = tmp.name
  seen.add(path)
  handle = C.CDLL(path)
  for name, (restype, argtypes) in N.SIGNATURES.items():
    fn = getattr(handle, name)
    fn.restype, fn.argtypes = restype, argtypes
  assert handle.grand_abi_version() == N.ABI_VERSION, 'ABI mismatch in %s' % path
  return handle


def timed_graph(body, launches, repeats=5):
  body(None)
  torch.cuda.synchronize()
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    body(N.current_stream(DEV))
  for _ in range(2):
    graph.replay()
  torch.cuda.synchronize()
  times = []
  for _ in range(repeats):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e3 / launches)
  times.sort()
  return times[len(times) // 2]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--workload', default='ogbn-arxiv')
  ap.add_argument('--max-degree', type=int, default=0, help='regenerate the graph with this maximum degree (hub test)')
  ap.add_argument('variants', nargs='+')
  args = ap.parse_args()
  ds, opt, x_host = bench.build_workload(args.workload, DEV)
  n, d = x_host.shape
  if args.max_degree:
    from graph_neural_pde_b200.synthetic import make_graph
    ds.data.edge_index = make_graph(n, ds.data.edge_index.shape[1] // 2, max_degree=args.max_degree, seed=0)
  data = Data(x=x_host, edge_index=ds.data.edge_index, edge_attr=None, num_nodes=n)
  g = torch.Generator().manual_seed(0)
  problems = {}
  for mode, cls in (('attention', ODEFuncTransformerAtt), ('weights', LaplacianODEFunc)):
    o = dict(opt, function='transformer' if mode == 'attention' else 'laplacian')
    block = ConstantODEblock(cls, (), o, data, DEV, t=torch.tensor([0, o['time']])).to(DEV)
    bench.init_trained_like(block)
    block.eval()
    x = x_host.to(DEV)
    block.set_x0(x)
    prob = NativeProblem(block.odefunc, x)
    eng = prob.engine
    ws = eng.workspace(prob.spec)
    y, k1, k2, k3, ya, yb = (eng.buffer(k) for k in ('y', 'k1', 'k2', 'k3', 'ya', 'yb'))
    y.copy_(x)
    h, third = 1.0, 1.0 / 3.0
    stage_eps = [
      (y, dict(f_out=k1, y_out=ya, coef_f=h * third, terms=[(y, 1.0)])),
      (ya, dict(f_out=k2, y_out=yb, coef_f=h, terms=[(y, 1.0), (k1, -h * third)])),
      (yb, dict(f_out=k3, y_out=ya, coef_f=h, terms=[(y, 1.0), (k1, h), (k2, -h)])),
      (ya, dict(y_out=yb, coef_f=h * 0.125, terms=[(y, 1.0), (k1, h * 0.125), (k2, 3 * h * 0.125),
                                                   (k3, 3 * h * 0.125)])),
    ]
    problems[mode] = dict(block=block, prob=prob, eng=eng, ws=ws, y=y, k1=k1, yb=yb,
                          rk4=[(src, prob.epilogue(**kw)) for src, kw in stage_eps],
                          f_only=[(y, prob.epilogue(f_out=k1))])
    if mode == 'attention':
      N.check(N.lib().grand_project_qk(y.data_ptr(), n, n, C.byref(prob.spec.struct), ws.q, ws.k,
                                       N.current_stream(DEV)))
  e = int(problems['attention']['block'].odefunc.edge_index.shape[1])
  plan = problems['attention']['eng'].plan
  print('N %d E %d d %d  units: long %d (rows %d) medium %d windows %d  seg_edges %d' % (
    n, e, d, plan.n_lunits, plan.n_lrows, plan.n_munits, plan.n_tiles - plan.n_split_tiles, plan.seg_edges), flush=True)

  seen, ref_out, results = set(), {}, {}
  for spec in args.variants:
    name, rest = spec.split('=', 1)
    parts = rest.split(',')
    saved = {}
    for kv in parts[1:]:
      k, v = kv.split('=', 1)
      saved[k] = os.environ.get(k)
      os.environ[k] = v
    lib = load_variant(parts[0], seen)
    row = {}
    for mode, P in problems.items():
      prob, eng, ws = P['prob'], P['eng'], P['ws']
      gs = C.byref(eng.plan.struct())
      for vname in ('rk4', 'f_only'):
        if mode == 'weights' and vname == 'rk4':
          continue
        eps = P[vname]
        reps = 8 // len(eps) * 2

        def body(st, eps=eps, reps=reps, mode=mode, prob=prob, ws=ws, gs=gs):
          st = N.current_stream(DEV) if st is None else st
          for _ in range(reps):
            for src, ep in eps:
              if mode == 'attention':
                rc = lib.grand_rhs_attention_edges(gs, C.byref(prob.spec.struct), src.data_ptr(), C.byref(ws),
                                                   C.byref(ep), st)
              else:
                rc = lib.grand_rhs_laplacian(gs, d, src.data_ptr(), prob.w_csr.data_ptr(), C.byref(ws), C.byref(ep), st)
              if rc != 0:
                raise RuntimeError('%s: rc %d %s' % (name, rc, lib.grand_last_error()))

        P['y'].copy_(x_host.to(DEV))
        us = timed_graph(body, reps * len(eps))
        row['%s/%s' % (mode, vname)] = us
        torch.cuda.synchronize()
        out = (P['k1'] if vname == 'f_only' else P['yb']).clone()
        key = (mode, vname)
        if key not in ref_out:
          ref_out[key] = out
          row['%s/%s/err' % (mode, vname)] = 0.0
        else:
          row['%s/%s/err' % (mode, vname)] = float((out - ref_out[key]).abs().max() / ref_out[key].abs().max())
    for k, v in saved.items():
      if v is None:
        os.environ.pop(k, None)
      else:
        os.environ[k] = v
    results[name] = row
    print('%-18s att/rk4 %7.1f us  att/f_only %7.1f us  weights/f_only %7.1f us   max rel diff vs first %.1e' % (
      name, row['attention/rk4'], row['attention/f_only'], row['weights/f_only'],
      max(v for k, v in row.items() if k.endswith('/err'))), flush=True)
  print(json.dumps(results))


if __name__ == '__main__':
  main()
