import os, sys, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench
from __graft_entry__ import load_package
load_package()
import graph_neural_pde_b200.graph_plan as GP
dev = torch.device('cuda:0')
def rhs(seg):
  GP.GraphPlan.__init__.__defaults__ = GP.GraphPlan.__init__.__defaults__[:-1] + (seg,)
  ds, opt, x, block = bench.make_block('ogbn-arxiv', dev)
  xd = x.to(dev)
  block.set_x0(xd)
  with torch.no_grad():
    f = block.odefunc(0.0, xd)
    f2 = block.odefunc(0.0, xd)
  plan = block.odefunc.engine(xd).plan
  return f, f2, plan
f256, f256b, p256 = rhs(256)
print('seg256 run-to-run', float((f256 - f256b).abs().max()))
for seg in (128, 64, 32):
  f, fb, p = rhs(seg)
  err = (f - f256).abs().max(dim=1).values
  bad = torch.nonzero(err > 1e-5 * float(f256.abs().max())).flatten()
  deg = (p.rowptr[1:] - p.rowptr[:-1]).long()
  print('seg', seg, 'units L/M', p.n_lunits, p.n_munits, 'lrows', p.n_lrows, 'max err', float(err.max()), 'bad rows', bad.numel(), 'run-to-run', float((f - fb).abs().max()))
  if bad.numel():
    d = deg[bad]
    print('   degrees of bad rows: min %d max %d; sample %s' % (int(d.min()), int(d.max()), d[:20].tolist()))
    print('   n segments of bad rows:', ((d + seg - 1) // seg)[:20].tolist())
