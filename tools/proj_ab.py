"""Projection / dense-layer kernels against an fp64 reference, and their timings, in one process:

    python tools/proj_ab.py            (GPU box)

For every shape: grand_project_qk with the mma.sync kernel (impl 1) and the tcgen05 kernel (impl 2), max error relative to
max |reference| (bar 1e-5, BASELINE.json) and microseconds per launch (CUDA events, 20 launches after warm-up); then
grand_linear (encoder / decoder shapes incl. padded outputs and the exact-fp32 fallback)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from graph_neural_pde_b200 import _native as N  # noqa: E402

DEV = torch.device('cuda:0')


def timed(fn, reps=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) * 1e3 / reps


def project(n_rows, n_cols, d, ad, seed=0):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(n_cols, d, generator=g).to(DEV)
  wq, wk = (torch.randn(ad, d, generator=g) / d ** 0.5).to(DEV), (torch.randn(ad, d, generator=g) / d ** 0.5).to(DEV)
  bq, bk = torch.randn(ad, generator=g).to(DEV), torch.randn(ad, generator=g).to(DEV)
  a = N.GrandAttention()
  a.d, a.ad, a.heads, a.score_type, a.norm_type, a.norm_idx = d, ad, 1, 0, 0, 0
  a.wq, a.bq, a.wk, a.bk = wq.data_ptr(), bq.data_ptr(), wk.data_ptr(), bk.data_ptr()
  want_q = (x[:n_rows].double() @ wq.double().t() + bq.double())
  want_k = (x.double() @ wk.double().t() + bk.double())
  st = N.current_stream(DEV)
  line = 'project n_rows %7d n_cols %7d d %4d ad %3d:' % (n_rows, n_cols, d, ad)
  for impl, name in ((1, 'mma'), (2, 'tc'), (3, 'tm')):
    N.check(N.lib().grand_set_projection_impl(impl))
    q = torch.full((n_rows, ad), float('nan'), device=DEV)
    k = torch.full((n_cols, ad), float('nan'), device=DEV)

    def run():
      N.check(N.lib().grand_project_qk(x.data_ptr(), n_rows, n_cols, C.byref(a), q.data_ptr(), k.data_ptr(), st))

    us = timed(run)
    eq = float((q.double() - want_q).abs().max() / want_q.abs().max()) if n_rows else 0.0
    ek = float((k.double() - want_k).abs().max() / want_k.abs().max())
    gbs = (n_cols * d + (n_rows + n_cols) * ad) * 4 / us / 1e3
    line += '  %s %7.1f us %6.0f GB/s err q %.1e k %.1e' % (name, us, gbs, eq, ek)
  print(line, flush=True)


def linear(rows, d_in, d_out, seed=0, bias=True):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(rows, d_in, generator=g).to(DEV)
  w = (torch.randn(d_out, d_in, generator=g) / d_in ** 0.5).to(DEV)
  b = torch.randn(d_out, generator=g).to(DEV) if bias else None
  want = x.double() @ w.double().t() + (b.double() if bias else 0)
  st = N.current_stream(DEV)
  line = 'linear  rows %7d d_in %5d d_out %4d:' % (rows, d_in, d_out)
  for impl, name in ((1, 'fp32'), (2, 'tc'), (3, 'tm')):
    N.check(N.lib().grand_set_projection_impl(impl))
    y = torch.full((rows, d_out), float('nan'), device=DEV)

    def run():
      N.check(N.lib().grand_linear(x.data_ptr(), rows, d_in, w.data_ptr(), b.data_ptr() if bias else None, d_out,
                                   y.data_ptr(), d_out, st))

    us = timed(run)
    err = float((y.double() - want).abs().max() / want.abs().max())
    line += '  %s %7.1f us err %.1e' % (name, us, err)
  tus = timed(lambda: torch.nn.functional.linear(x, w, b))
  print(line + '  cublas %7.1f us' % tus, flush=True)


if __name__ == '__main__':
  torch.backends.cuda.matmul.allow_tf32 = False
  if '--arxiv' in sys.argv:
    project(169343, 169343, 128, 32)
    linear(169343, 128, 128)
    sys.exit(0)
  project(1000, 1000, 128, 32)
  project(300, 517, 80, 16)
  project(169343, 169343, 128, 32)
  project(169343, 169343, 128, 32, seed=1)
  project(2485, 2485, 80, 16)
  project(7650, 7650, 64, 32)
  project(5000, 5000, 100, 24)      # d not a multiple of 32 (zero-filled chunk tail); N = 48
  project(4096, 4096, 16, 8)
  project(9000, 9000, 128, 128)     # N = 256
  project(3000, 3000, 36, 10)       # N = 20: not a multiple of 16 -> mma.sync / ffma on both
  linear(169343, 128, 128)
  linear(169343, 128, 40)
  linear(169343, 40, 128, bias=False)
  linear(2485, 1433, 80)
  linear(2485, 80, 7)
  N.check(N.lib().grand_set_projection_impl(-1))
