"""Build libgrand_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and runnable directly:

    python graph-neural-pde_b200/build_native.py [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgrand_b200.so')
SOURCES = ['project.cu', 'project_mma.cu', 'project_tc.cu', 'linear.cu', 'rhs_window.cu', 'rhs_stream.cu', 'rhs_ring.cu', 'rhs_coop.cu', 'rhs_backward.cu', 'edge_kernels.cu', 'capi.cu', 'dopri5_device.cu',
           'plan_host.cu']
HEADERS = ['common.cuh', 'attention.cuh', 'edge_common.cuh', 'rhs_stream.cu', os.path.join('..', '..', 'include', 'grand_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']


def _nvcc():
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return 'nvcc'


def _stale(target, deps):
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
  os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
  hdrs = [os.path.join(CSRC, h) for h in HEADERS]
  objs, jobs = [], []
  for src in SOURCES:
    s = os.path.join(CSRC, src)
    o = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
    objs.append(o)
    if force or _stale(o, [s] + hdrs):
      jobs.append([_nvcc()] + NVCC_FLAGS + ['-c', s, '-o', o])

  def run(cmd):
    if verbose:
      print(' '.join(cmd), flush=True)
    subprocess.run(cmd, check=True)

  with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
    list(ex.map(run, jobs))
  if force or jobs or _stale(LIB, objs):
    run([_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs)  # static cudart
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv))
