"""Tight fp64 solutions z(T) of every adaptive-solver (dopri5) golden block case, computed once with the oracle
(oracle/grand_oracle.py odeint_dopri5 at rtol = atol = 1e-13) and cached in tests/golden/exact_dopri5.pt.

They are the arbiter of the adaptive-solver parity criterion (tests/test_gpu_parity.py::test_blocks_golden): both the
reference's z and ours must be equally close to the exact solution.  Computing them takes ~10 s per case on CPU, so the
tests read this cache and fall back to computing when a case is missing.   python tests/golden/make_exact.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..')))

import _cases as C  # noqa: E402


def main():
  cases = {}
  for name in ('golden_v1.pt', 'golden_v2.pt'):
    cases.update(torch.load(os.path.join(HERE, name), map_location='cpu', weights_only=False))
  out = {}
  for name, c in sorted(cases.items()):
    if c['kind'] == 'block' and c['opt']['method'] == 'dopri5':
      out[name] = C.compute_exact_block_solution(c)
      print('%-55s |z| %.6f' % (name, float(out[name].abs().max())), flush=True)
  torch.save(out, os.path.join(HERE, 'exact_dopri5.pt'))
  print('wrote %d solutions' % len(out))


if __name__ == '__main__':
  main()
