#!/usr/bin/env python
"""bench.py -- ODE RHS throughput of the GRAND-nl hot path (BASELINE.json metric) on synthetic ogbn-arxiv-shaped input.

    python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU path (oracle restatement)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # node-partitioned, N > 1

A "step" is one forward pass of the ODE block over the whole graph: rk4 (3/8 rule), T = 5.08, h = 1 -> 6 solver steps
= 24 RHS evaluations (BASELINE.json configs[4]).  metric = edges*heads/s = E * H * evals / time.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'ODE RHS evals/sec (edges*heads/s)'
UNIT = 'edges*heads/s'
L2_MB = 126.0


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    try:
      return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
      pass
  return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def algorithmic_bytes(n, e, d, ad, add_source):
  """compulsory traffic of one GRAND-nl RHS evaluation (int32 CSR, Q/K materialised once), split into
  the projection kernel's share and the edge (dominant) kernel's share."""
  proj = 4 * (n * d + 2 * n * ad)
  edge = 4 * (2 * n * ad + n * d + n * d + (n * d if add_source else 0)) + 4 * e + 4 * (n + 1)
  return proj, edge


class ClockSampler:
  """nvidia-smi clock / throttle-reason samples while the timed region runs (B200_PROFILING.md clocks line)."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index=0):
    self.index, self.lines, self.proc = index, [], None

  def __enter__(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None
    return self

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def __exit__(self, *exc):
    if self.proc is not None:
      self.proc.terminate()
      try:
        self.proc.wait(timeout=2)
      except Exception:
        self.proc.kill()

  def summary(self):
    sm, mx, reasons = [], 0, set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for line in self.lines:
      parts = [p.strip() for p in line.split(',')]
      if len(parts) < 7:
        continue
      try:
        sm.append(float(parts[0]))
        mx = max(mx, float(parts[1]))
      except ValueError:
        continue
      for name, flag in zip(names, parts[3:7]):
        if flag.lower().startswith('active'):
          reasons.add(name)
    sm.sort()
    med = sm[len(sm) // 2] if sm else None
    return {'sm_mhz': med, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons), 'samples': len(sm)}


WORKLOADS = {'ogbn-arxiv': ('ogbn-arxiv', 'C5'), 'Cora': ('Cora', 'C2'), 'Cora-GRAND-l': ('Cora', 'C1'),
             'Photo': ('Photo', 'C3'), 'Cora-BLEND': ('Cora', 'C4')}


def build_workload(name, device=None, seed=0):
  from __graft_entry__ import load_package
  load_package()
  from graph_neural_pde_b200.synthetic import SyntheticDataset, baseline_opt
  dataset, key = WORKLOADS[name]
  ds = SyntheticDataset(dataset, seed=seed, num_features=8)   # features: not on the path
  opt = baseline_opt(key)
  n = ds.data.num_nodes
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(n, opt['hidden_dim'], generator=g)
  return ds, opt, x


def init_trained_like(block, seed=0):
  """'trained-like' parameters (SURVEY.md 8(d) regime ii): W ~ N(0, 1/d), alpha_train = beta_train = 0."""
  g = torch.Generator().manual_seed(seed)
  with torch.no_grad():
    for name, p in block.named_parameters():
      if p.dim() == 2:
        p.copy_((torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5).to(p.device))
    # the Linear biases keep PyTorch's default law U(+-1/sqrt(fan_in)) (reference :122-126 leaves them alone) but are
    # drawn from the seeded generator: every rank of a multi-GPU run, and every block of an A/B, gets the same values
    for name, p in block.named_parameters():
      if name.endswith('.bias') and p.dim() == 1:
        fan_in = dict(block.named_parameters())[name[:-5] + '.weight'].shape[1]
        p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) / fan_in ** 0.5).to(p.device))


def make_block(name, device):
  """The reference's ConstantODEblock over the workload's ODE function (this repo's mirror of both)."""
  from __graft_entry__ import load_package
  load_package()
  from graph_neural_pde_b200.block_constant import ConstantODEblock
  from graph_neural_pde_b200.function_laplacian_diffusion import LaplacianODEFunc
  from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt
  from graph_neural_pde_b200.synthetic import Data
  ds, opt, x_host = build_workload(name)
  n = ds.data.num_nodes
  data = Data(x=x_host, edge_index=ds.data.edge_index, edge_attr=None, num_nodes=n)
  cls = ODEFuncTransformerAtt if opt['function'] == 'transformer' else LaplacianODEFunc
  block = ConstantODEblock(cls, (), opt, data, device, t=torch.tensor([0, opt['time']])).to(device)
  init_trained_like(block)
  block.eval()
  return ds, opt, x_host, block


def oracle_func(block_state, opt, edge_index, n, x0):
  from oracle import grand_oracle as O
  ei, ew = O.constant_block_graph(edge_index, None, n, opt, torch.float32)
  alpha, beta = block_state['odefunc.alpha_train'], block_state['odefunc.beta_train']
  if opt['function'] != 'transformer':
    return lambda t, y: O.laplacian_rhs(y, ei, ew, opt, alpha, beta, x0)
  prefix = 'odefunc.multihead_att_layer.'
  P = {k[len(prefix):]: v for k, v in block_state.items() if k.startswith(prefix)}
  return lambda t, y: O.transformer_rhs(y, ei, P, opt, alpha, beta, x0)


def cpu_thread_sweep(func, x, budget_s=60.0):
  """Best torch thread count for the CPU oracle on this box: one warm-up + two timed RHS evaluations at every count in
  {8, 16, 32, 64, 128, all} that the box has (the 128-thread default of round 1 was 2-3x slower than 8-16 threads for
  these memory-bound scatter / gather ops) -> (threads, seconds per eval, {threads: seconds})."""
  total = os.cpu_count() or 1
  cands = sorted({t for t in (8, 16, 32, 64, 128, total) if t <= total} or {total})
  seen, t_start = {}, time.perf_counter()
  with torch.no_grad():
    for t in cands:
      torch.set_num_threads(t)
      func(0.0, x)
      t0 = time.perf_counter()
      func(0.0, x)
      func(0.0, x)
      seen[t] = (time.perf_counter() - t0) / 2
      if time.perf_counter() - t_start > budget_s:
        break
  best = min(seen, key=seen.get)
  torch.set_num_threads(best)
  return best, seen[best], seen


def time_cpu_reference(func, x, h, warmup, steps, budget_s=200.0, t_eval=None):
  """Times `steps` samples of the CPU oracle after `warmup` untimed ones -> (list of seconds per sample, RHS evals per
  sample).  A sample is one RK4 (3/8 rule) step = 4 RHS evals; when (warmup + steps) of those would not fit `budget_s`
  of wall clock a sample is bounded to ONE RHS evaluation, so that the run ends within minutes whatever --steps says."""
  from oracle import grand_oracle as O

  with torch.no_grad():
    if t_eval is None:
      t0 = time.perf_counter()
      func(0.0, x)
      t_eval = time.perf_counter() - t0
    evals = 4 if (warmup + steps) * 4 * t_eval <= budget_s else 1

    def sample():
      return x + O.rk4_alt_step(func, 0.0, h, h, x) if evals == 4 else func(0.0, x)

    for _ in range(warmup):
      sample()
    times = []
    for _ in range(steps):
      t0 = time.perf_counter()
      sample()
      times.append(time.perf_counter() - t0)
  return times, evals


def median(v):
  v = sorted(v)
  return v[len(v) // 2]


def workload_config(name, opt, n, e, world=1):
  """The same dict in the native and the reference arm (the driver compares them key by key)."""
  fn = 'GRAND-nl (transformer attention, %d heads' % opt['heads'] if opt['function'] == 'transformer' else 'GRAND-l (laplacian'
  return {'workload': '%s %s, %s, T=%s, h=%s%s), synthetic Chung-Lu graph with the dataset\'s N/E/max-degree' % (
            name, fn, opt['method'], opt['time'], opt['step_size'], ', add_source' if opt['add_source'] else ''),
          'nodes': n, 'edges_incl_self_loops': e, 'hidden_dim': opt['hidden_dim'],
          'attention_dim': opt['attention_dim'] if opt['function'] == 'transformer' else 0,
          'heads': opt['heads'] if opt['function'] == 'transformer' else 1,
          'parallelism': 'single GPU' if world == 1 else 'np%d: 1-D node partition, halo rows pushed over NVLink' % world,
          'l2_policy': 'no flush: the per-step working set (state, x0, 5 stage buffers, Q/K) of the arxiv workload is '
                       '715 MB > 126 MB L2; the Cora / Photo sub-workloads are L2-resident by nature and say so',
          'weights': 'trained-like N(0,1/d), seed 0'}


def run_reference(args, rank, world):
  if rank != 0:
    return
  name = args.workload
  ds, opt, x, block = make_block(name, torch.device('cpu'))
  n = ds.data.num_nodes
  state = {k: v.detach() for k, v in block.state_dict().items()}
  func = oracle_func(state, opt, ds.data.edge_index, n, x)
  e = int(block.odefunc.edge_index.shape[1])
  heads = opt['heads'] if opt['function'] == 'transformer' else 1
  threads, t_eval, sweep = cpu_thread_sweep(func, x)
  times, evals = time_cpu_reference(func, x, float(opt['step_size']), args.warmup, args.steps, t_eval=t_eval)
  sec = median(times)
  value = e * heads * evals / sec
  sample = ('1 rk4 (3/8) step = 4 RHS evals' if evals == 4 else
            '1 RHS eval (bounded: an rk4 step per sample would not fit the time budget)')
  line = {
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
    'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong',
    'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': workload_config(name, opt, n, e, world=max(world, args.gpus)),
    'sample': sample + ' per step; value = median over the timed steps',
    'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                     'host_cores': os.cpu_count(), 'thread_sweep_s_per_eval': {str(k): round(v, 4) for k, v in sweep.items()},
                     'sample': 'each step = ' + sample + ' of the oracle restatement (torch CPU ops) at the best thread '
                               'count of a sweep; the reference itself cannot be imported: its third-party wheels are absent'},
    'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  emit(line)


# ---- GPU timing helpers -----------------------------------------------------------------------------------------------
def _ev():
  return torch.cuda.Event(enable_timing=True)


def warm_up(fn, min_s=1.0, max_s=5.0, tol=0.02):
  """Run `fn` until at least `min_s` of wall clock has passed AND three consecutive steps agree within `tol` (clocks,
  caches and the CUDA-Graph capture have settled), at most `max_s`."""
  last, t_start = [], time.perf_counter()
  while True:
    a, b = _ev(), _ev()
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    last = (last + [a.elapsed_time(b)])[-3:]
    el = time.perf_counter() - t_start
    if el >= max_s or (el >= min_s and len(last) == 3 and max(last) - min(last) <= tol * min(last)):
      return el


def timed_repeats(fn, steps, repeats, sync=None):
  """`repeats` timed regions of exactly `steps` calls each, bracketed by a synchronise (and `sync`, the multi-rank
  barrier) on both sides -> list of ms per step."""
  out = []
  for _ in range(repeats):
    torch.cuda.synchronize()
    if sync is not None:
      sync()
      torch.cuda.synchronize()
    a, b = _ev(), _ev()
    a.record()
    for _ in range(steps):
      fn()
    b.record()
    torch.cuda.synchronize()
    if sync is not None:
      sync()
    out.append(a.elapsed_time(b) / steps)
  return out


def timed_graph(body, count, stream_of, repeats=5):
  """Replay `body` from a CUDA Graph (host launch latency is not what is measured), CUDA events on the replay stream;
  returns the median ms per launch group."""
  body(None)
  torch.cuda.synchronize()
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    body(stream_of())
  for _ in range(2):
    graph.replay()
  torch.cuda.synchronize()
  times = []
  for _ in range(repeats):
    e0, e1 = _ev(), _ev()
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / count)
  return median(times)


def algorithmic_bytes_l(n, e, d, add_source):
  """SURVEY.md 8(d), GRAND-l: B_l = 4 [2 N d (+ N d)] + 8 E + 4 (N + 1)."""
  return 4 * (2 * n * d + (n * d if add_source else 0)) + 8 * e + 4 * (n + 1)


def edge_kernel_name(attention, d, n_edges=1 << 20):
  """Which edge kernel the library launches for this workload (csrc/rhs_ring.cu ring_kernel_ok, csrc/rhs_stream.cu)."""
  sel = os.environ.get('GRAND_EDGE', '')
  ring_ok = d <= 128 and d % 4 == 0
  if sel == 'window':
    base = 'rhs_window_kernel'
  elif ring_ok and sel != 'stream' and ((attention and n_edges >= (1 << 20)) or sel == 'ring'):
    base = 'rhs_ring_kernel (neighbour rows staged through shared memory by cp.async.bulk)'
  else:
    base = 'rhs_stream_kernel'
  return base + ('<MODE_FUSED> (+ lseg_stats pre-pass of the long rows)' if attention else '<MODE_WEIGHT>')


def dominant_traffic(name):
  """ncu dram__bytes_read + write per launch of the dominant kernel, per workload (profiles/dominant_kernel_traffic.json);
  None where no capture of that workload exists."""
  path = os.path.join(ROOT, 'profiles', 'dominant_kernel_traffic.json')
  try:
    return json.load(open(path)).get('per_workload', {}).get(name)
  except Exception:
    return None


def measure_workload(name, args, dev, full):
  """One BASELINE.json configuration on one GPU.  full: everything the main line needs (e2e legs, dominant-kernel
  roofline, CPU baseline with >= 5 samples); otherwise the resident timing, the whole-evaluation roofline fraction and
  a short CPU baseline (the `other_workloads` entries)."""
  import ctypes as C
  from graph_neural_pde_b200 import _native as N
  from graph_neural_pde_b200.engine import fixed_grid_dts
  from graph_neural_pde_b200.odeint import NativeProblem

  ds, opt, x_host, block = make_block(name, dev)
  n, d = x_host.shape
  func = block.odefunc
  attention = opt['function'] == 'transformer'
  e = int(func.edge_index.shape[1])
  heads = opt['heads'] if attention else 1
  ad = opt['attention_dim'] if attention else 0
  steps = args.steps if full else max(3, min(args.steps, 10))

  x_pinned = x_host.pin_memory()
  z_pinned = torch.empty_like(x_pinned).pin_memory()
  x_dev = x_pinned.to(dev)

  def step_resident():
    func.nfe = 0   # run_GNN.py resets the NFE meters every epoch (reference src/run_GNN.py:90-95)
    block.set_x0(x_dev)
    with torch.no_grad():
      return block(x_dev)

  def step_e2e():
    func.nfe = 0
    x_dev.copy_(x_pinned, non_blocking=True)
    block.set_x0(x_dev)
    with torch.no_grad():
      z = block(x_dev)
    z_pinned.copy_(z, non_blocking=True)
    return z

  # launches per step and RHS evaluations per step, counted on the first (eager, not yet graph-captured) forward
  before = N.lib().grand_launch_count()
  step_resident()
  launches_per_step = int(N.lib().grand_launch_count() - before)
  evals = int(func.nfe)
  torch.cuda.synchronize()
  warm_s = warm_up(step_resident, min_s=1.0 if full else 0.3, max_s=5.0 if full else 1.5)

  res = {'name': name, 'data_source': getattr(ds, 'source', 'synthetic'),
         'evals': evals, 'launches_per_step': launches_per_step, 'n': n, 'e': e, 'd': d, 'ad': ad,
         'heads': heads, 'opt': opt, 'warm_up_seconds': warm_s}
  with ClockSampler(dev.index or 0) as clocks:
    # ---- timed region 1: inputs resident in HBM ----
    rep = timed_repeats(step_resident, steps, args.repeats)
    res['ms_repeats'] = rep
    res['ms_step'] = median(rep)
    if full:
      # ---- timed region 2: end to end through the block API from pinned host buffers ----
      # (a) serial: H2D -> block(x) -> D2H one after the other on one stream
      for _ in range(2):
        step_e2e()
      res['ms_e2e_serial'] = median(timed_repeats(step_e2e, steps, args.repeats))
      # (b) pipelined: the same per-step work (H2D of x, block(x), D2H of z, all inside the timed region), double
      #     buffered so that the copies of neighbouring steps overlap the compute on the two copy engines
      cur = torch.cuda.current_stream()
      h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
      xin = [x_dev, torch.empty_like(x_dev)]
      zout = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
      zpin = [z_pinned, torch.empty_like(z_pinned).pin_memory()]
      in_ready, in_free = [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]
      out_ready, out_free = [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]
      it_no = [0]

      def pipelined_step():
        b_ = it_no[0] & 1
        it_no[0] += 1
        with torch.cuda.stream(h2d):
          h2d.wait_event(in_free[b_])
          xin[b_].copy_(x_pinned, non_blocking=True)
          in_ready[b_].record(h2d)
        cur.wait_event(in_ready[b_])
        cur.wait_event(out_free[b_])
        func.nfe = 0
        block.set_x0(xin[b_])
        with torch.no_grad():
          zout[b_].copy_(block(xin[b_]))
        in_free[b_].record(cur)
        out_ready[b_].record(cur)
        with torch.cuda.stream(d2h):
          d2h.wait_event(out_ready[b_])
          zpin[b_].copy_(zout[b_], non_blocking=True)
          out_free[b_].record(d2h)

      def pipelined(count):
        for b_ in (0, 1):
          in_free[b_].record(cur)
          out_free[b_].record(d2h)
        it_no[0] = 0
        for _ in range(count):
          pipelined_step()
        cur.wait_stream(d2h)

      pipelined(2)
      reps = []
      for _ in range(args.repeats):
        torch.cuda.synchronize()
        a5, b5 = _ev(), _ev()
        a5.record()
        pipelined(steps)
        b5.record()
        torch.cuda.synchronize()
        reps.append(a5.elapsed_time(b5) / steps)
      res['ms_e2e'] = median(reps)
      res['h2d_bytes'], res['d2h_bytes'] = int(x_pinned.numel() * 4), int(z_pinned.numel() * 4)

      # ---- dominant kernel on its own: the fused edge kernel with the four rk4 stage epilogues ----
      block.set_x0(x_dev)
      prob = NativeProblem(func, x_dev)
      eng = prob.engine
      ybuf = eng.buffer('y')
      ybuf.copy_(x_dev)
      ws = eng.workspace(prob.spec)
      stage = eng.buffer('stages', rows=5 * n)
      yb, yc, k1, k2, k3 = (stage[i * n:(i + 1) * n] for i in range(5))
      dts, _ = fixed_grid_dts(0.0, opt['time'], opt['step_size'])
      h = dts[0]
      third = 1.0 / 3.0
      stages = [
        (ybuf, dict(f_out=k1, y_out=yb, coef_f=h * third, terms=[(ybuf, 1.0)])),
        (yb, dict(f_out=k2, y_out=yc, coef_f=h, terms=[(ybuf, 1.0), (k1, -h * third)])),
        (yc, dict(f_out=k3, y_out=yb, coef_f=h, terms=[(ybuf, 1.0), (k1, h), (k2, -h)])),
        (yb, dict(y_out=yc, coef_f=h * 0.125, terms=[(ybuf, 1.0), (k1, h * 0.125), (k2, 3 * h * 0.125),
                                                       (k3, 3 * h * 0.125)])),
      ]
      eps = [(src, prob.epilogue(**kw)) for src, kw in stages]
      lib = N.lib()
      gs = C.byref(eng.plan.struct())
      stream_of = lambda: N.current_stream(dev)
      reps_k = 6
      if attention:
        as_ = C.byref(prob.spec.struct)
        N.check(lib.grand_project_qk(ybuf.data_ptr(), n, n, as_, ws.q, ws.k, stream_of()))

        def edge_body(st):
          st = stream_of() if st is None else st
          for _ in range(reps_k):
            for src, ep in eps:
              N.check(lib.grand_rhs_attention_edges(gs, as_, src.data_ptr(), C.byref(ws), C.byref(ep), st))

        def proj_body(st):
          st = stream_of() if st is None else st
          for _ in range(reps_k * 4):
            N.check(lib.grand_project_qk(ybuf.data_ptr(), n, n, as_, ws.q, ws.k, st))

        res['ms_edge'] = timed_graph(edge_body, reps_k * len(eps), stream_of)   # edge kernel + its long-row pre-pass
        res['ms_proj'] = timed_graph(proj_body, reps_k * 4, stream_of)
      else:
        def edge_body(st):
          st = stream_of() if st is None else st
          for _ in range(reps_k):
            for src, ep in eps:
              N.check(lib.grand_rhs_laplacian(gs, d, src.data_ptr(), prob.w_csr.data_ptr(), C.byref(ws), C.byref(ep), st))

        res['ms_edge'] = timed_graph(edge_body, reps_k * len(eps), stream_of)
        res['ms_proj'] = 0.0
  res['clocks'] = clocks.summary()

  # ---- CPU baseline: the oracle on this box's host cores, bounded sample, best thread count of a sweep ----
  res['cpu'] = None
  if not args.skip_cpu:
    try:
      state = {k: v.detach().cpu() for k, v in block.state_dict().items()}
      cpu_func = oracle_func(state, opt, ds.data.edge_index, n, x_host)
      threads, t_eval, sweep = cpu_thread_sweep(cpu_func, x_host, budget_s=60.0 if full else 10.0)
      times, cpu_evals = time_cpu_reference(cpu_func, x_host, float(opt['step_size']), 1, 5,
                                            budget_s=30.0 if full else 5.0, t_eval=t_eval)
      sec = median(times)
      res['cpu'] = {'value': e * heads * cpu_evals / sec, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                    'host_cores': os.cpu_count(),
                    'thread_sweep_s_per_eval': {str(k): round(v, 5) for k, v in sweep.items()},
                    'sample': 'median of 5 timed samples of %s of the oracle (torch CPU ops) after 1 warm-up, at the '
                              'best thread count of the sweep' % ('1 rk4 step = 4 RHS evals' if cpu_evals == 4 else '1 RHS eval'),
                    'ms_per_rhs_eval': sec * 1e3 / cpu_evals}
    except Exception as exc:   # a sub-workload's baseline must not take the main line down
      res['cpu'] = {'value': None, 'error': repr(exc)[:200]}
  del block
  torch.cuda.empty_cache()
  return res


def eval_bytes(res):
  """Algorithmic bytes of one RHS evaluation -> (projection share, edge share)."""
  n, e, d, ad, opt = res['n'], res['e'], res['d'], res['ad'], res['opt']
  if opt['function'] == 'transformer':
    return algorithmic_bytes(n, e, d, ad, opt['add_source'])
  return 0, algorithmic_bytes_l(n, e, d, opt['add_source'])


def sub_result(res, peak):
  """`other_workloads` entry: value, us per evaluation, whole-evaluation roofline fraction, CPU baseline."""
  proj_b, edge_b = eval_bytes(res)
  us_eval = res['ms_step'] * 1e3 / res['evals']
  achieved = (proj_b + edge_b) / (us_eval * 1e-6) / 1e9
  return {'workload': workload_config(res['name'], res['opt'], res['n'], res['e'])['workload'],
          'value': res['e'] * res['heads'] * res['evals'] / (res['ms_step'] * 1e-3), 'unit': UNIT,
          'ms_per_step': res['ms_step'], 'rhs_evals_per_step': res['evals'], 'us_per_rhs_eval': us_eval,
          'gpu_launches_per_step': res['launches_per_step'],
          'roofline': {'bound': 'hbm', 'kernel': 'whole RHS evaluation', 'achieved': achieved, 'peak': peak,
                       'unit': 'GB/s', 'frac': achieved / peak, 'traffic': dominant_traffic(res['name']),
                       'algorithmic_bytes_per_eval': proj_b + edge_b,
                       'note': 'working set << 126 MB L2 and bytes / peak << launch latency: this fraction measures '
                               'launch / dependency latency, not bandwidth (SURVEY.md 8(d) caveat)'},
          'cpu_baseline': res['cpu'], 'clocks': res['clocks']}


def run_native_train(args):
  """--mode train: one training step of the block = forward + backward.  ogbn-arxiv: the reference's
  ray_tune configuration (src/ray_tune.py:19-22): adjoint rk4, adjoint_step_size 1 -> forward 24 RHS evaluations on the
  fused engine, backward 24 stages of the augmented system, each = one forward RHS + its vector-Jacobian product."""
  from __graft_entry__ import load_package
  load_package()
  from graph_neural_pde_b200 import _native as N
  dev = torch.device('cuda:0')
  torch.cuda.set_device(dev)
  name = args.workload
  ds, opt, x_host = build_workload(name)
  if opt['method'] != 'rk4' or opt['function'] != 'transformer':
    raise SystemExit('--mode train is wired for the rk4 attention workloads (ogbn-arxiv, Cora, Cora-BLEND)')
  from graph_neural_pde_b200.block_constant import ConstantODEblock
  from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt
  from graph_neural_pde_b200.synthetic import Data
  opt = dict(opt, adjoint=True, adjoint_method='rk4', adjoint_step_size=opt['step_size'])
  n, d = x_host.shape
  data = Data(x=x_host, edge_index=ds.data.edge_index, edge_attr=None, num_nodes=n)
  block = ConstantODEblock(ODEFuncTransformerAtt, (), opt, data, dev, t=torch.tensor([0, opt['time']])).to(dev)
  init_trained_like(block)
  block.train()
  func = block.odefunc
  e = int(func.edge_index.shape[1])
  x_dev = x_host.to(dev)
  w = torch.randn(n, d, generator=torch.Generator().manual_seed(1)).to(dev) / n

  def step():
    func.nfe = 0
    for p in block.parameters():
      p.grad = None
    xin = x_dev.clone().requires_grad_(True)
    block.set_x0(xin.detach())
    z = block(xin)
    (z * w).sum().backward()
    return xin.grad

  before = N.lib().grand_launch_count()
  step()
  launches = int(N.lib().grand_launch_count() - before)
  torch.cuda.synchronize()
  warm_up(step, min_s=1.0, max_s=5.0)
  steps = max(2, min(args.steps, 5))
  with ClockSampler(0) as clocks:
    rep = timed_repeats(step, steps, args.repeats)
  ms = median(rep)
  fwd = 4 * len(__import__('graph_neural_pde_b200.engine', fromlist=['fixed_grid_dts']).fixed_grid_dts(0.0, opt['time'], opt['step_size'])[0])
  line = {'metric': METRIC, 'mode': 'train (forward + adjoint backward)', 'value': e * opt['heads'] * 2 * fwd / (ms * 1e-3),
          'unit': UNIT, 'n_gpus': 1, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
          'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': workload_config(name, opt, n, e),
          'timing': {'ms_per_step_repeats': rep, 'repeats': args.repeats, 'statistic': 'median'},
          'rhs_evals_forward': fwd, 'adjoint_stages_backward': fwd,
          'value_note': 'edges*heads * (forward RHS evaluations + backward adjoint stages) / time; a backward stage is one '
                        'RHS evaluation plus its vector-Jacobian product (edge-phase VJP kernel, transposed aggregation, '
                        'cuBLAS tails)',
          'gpu_launches_per_step': launches, 'clocks': clocks.summary()}
  emit(line)


def run_native_single(args):
  from __graft_entry__ import load_package
  load_package()
  dev = torch.device('cuda:0')
  torch.cuda.set_device(dev)
  name = args.workload
  res = measure_workload(name, args, dev, full=True)
  peak, peak_src = measured_peaks()
  n, e, d, ad, heads, evals, opt = res['n'], res['e'], res['d'], res['ad'], res['heads'], res['evals'], res['opt']
  ms_step, ms_edge, ms_proj = res['ms_step'], res['ms_edge'], res['ms_proj']
  proj_bytes, edge_bytes = eval_bytes(res)
  achieved = edge_bytes / (ms_edge * 1e-3) / 1e9
  others = {}
  if not args.no_others:
    for other in WORKLOADS:
      if other != name:
        try:
          others[other] = sub_result(measure_workload(other, args, dev, full=False), peak)
        except Exception as exc:
          others[other] = {'error': repr(exc)[:300]}
  streams = 4.25 * n * d * 4
  line = {
    'metric': METRIC, 'value': e * heads * evals / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps,
    'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
    'dtype': 'f32',
    'data': 'synthetic' if res['data_source'] == 'synthetic' else 'real edge list from %s, synthetic state' % res['data_source'],
    'config': workload_config(name, opt, n, e),
    'timing': {'ms_per_step_repeats': res['ms_repeats'], 'repeats': args.repeats, 'statistic': 'median',
               'warm_up_seconds': res['warm_up_seconds'],
               'warm_up_rule': '--warmup steps, then until >= 1 s has passed and 3 consecutive steps agree within 2 %'},
    'rhs_evals_per_step': evals, 'us_per_rhs_eval': ms_step * 1e3 / evals,
    'e2e': {'value': e * heads * evals / (res['ms_e2e'] * 1e-3), 'unit': UNIT, 'ms_per_step': res['ms_e2e'],
            'h2d_bytes_per_step': res['h2d_bytes'], 'd2h_bytes_per_step': res['d2h_bytes'],
            'mode': 'double-buffered: every step copies x from pinned host memory, runs block(x) and copies z back; '
                    'the copies of step i+-1 overlap the compute of step i on the copy engines',
            'serial_ms_per_step': res['ms_e2e_serial'],
            'serial_value': e * heads * evals / (res['ms_e2e_serial'] * 1e-3)},
    'gpu_launches': res['launches_per_step'] * args.steps,
    'gpu_launches_per_step': res['launches_per_step'],
    'clocks': res['clocks'],
    'roofline': {'bound': 'hbm',
                 'kernel': edge_kernel_name(bool(ad), d, e),
                 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                 'traffic': dominant_traffic(name), 'peak_source': peak_src,
                 'algorithmic_bytes_per_launch': edge_bytes,
                 'note': 'algorithmic bytes follow SURVEY.md 8(d) (RHS only); the kernel additionally streams the rk4 '
                         'stage-combine operands the reference moves in separate axpy kernels (avg 2.5 reads + 1.75 '
                         'writes of [N,d] per launch), see fused_stream_bytes.  What bounds it is in profiles/README.md',
                 'fused_stream_bytes': int(streams),
                 'frac_incl_fused_streams': (edge_bytes + streams) / (ms_edge * 1e-3) / 1e9 / peak,
                 'us_per_launch': ms_edge * 1e3, 'projection_us_per_launch': ms_proj * 1e3,
                 'projection_algorithmic_bytes': proj_bytes,
                 'projection_achieved_GBs': proj_bytes / (ms_proj * 1e-3) / 1e9 if ms_proj else None,
                 'rhs_eval_algorithmic_bytes': proj_bytes + edge_bytes,
                 'rhs_eval_achieved_GBs': (proj_bytes + edge_bytes) / (ms_step * 1e-3 / evals) / 1e9,
                 'rhs_eval_frac': (proj_bytes + edge_bytes) / (ms_step * 1e-3 / evals) / 1e9 / peak},
    'cpu_baseline': res['cpu'],
    'other_workloads': others,
  }
  emit(line)


def run_native_multi(args, rank, world):
  """N > 1: 1-D node partition of the same workload over the ranks (strong scaling).  Halo rows are stored straight into
  the peers' buffers over NVLink by the kernels that produce them; one small barrier kernel per RHS evaluation; the
  whole forward is one CUDA Graph per rank.  Timing: barrier + synchronize on both sides of every repeat, CUDA events,
  MAX over ranks of the per-rank median."""
  import ctypes as C
  import torch.distributed as dist
  from __graft_entry__ import load_package
  load_package()
  from graph_neural_pde_b200 import _native as N
  from graph_neural_pde_b200.distributed import ShardedTransformerRHS
  from graph_neural_pde_b200.engine import fixed_grid_dts

  local_rank = int(os.environ.get('LOCAL_RANK', rank))
  dev = torch.device('cuda', local_rank)
  torch.cuda.set_device(dev)
  if not dist.is_initialized():
    dist.init_process_group('nccl', device_id=dev)
  name = args.workload
  ds, opt, x_host, block = make_block(name, dev)      # same seed on every rank: replicated graph and parameters
  n, d = x_host.shape
  func = block.odefunc
  e = int(func.edge_index.shape[1])
  heads, ad = opt['heads'], opt['attention_dim']
  dts, _ = fixed_grid_dts(0.0, opt['time'], opt['step_size'])
  evals = 4 * len(dts)
  shard = ShardedTransformerRHS(func, n, dev, transport=args.transport)
  part = shard.part
  x_pinned = x_host[part.r0:part.r1].contiguous().pin_memory()
  z_pinned = torch.empty_like(x_pinned).pin_memory()
  x_dev = x_pinned.to(dev)

  def step_resident():
    func.nfe = 0
    with torch.no_grad():
      return shard.forward_block(x_dev, x_dev, opt)

  def step_e2e():
    func.nfe = 0
    x_dev.copy_(x_pinned, non_blocking=True)
    with torch.no_grad():
      z = shard.forward_block(x_dev, x_dev, opt)
    z_pinned.copy_(z, non_blocking=True)

  before = N.lib().grand_launch_count()
  z_first = step_resident().clone()
  launches_per_step = int(N.lib().grand_launch_count() - before)
  torch.cuda.synchronize()
  dist.barrier()

  # ---- parity, outside the timed region: rank 0 runs the whole graph on its GPU through the single-GPU engine ----
  parity = torch.zeros(1, device=dev)
  if rank == 0:
    xd = x_host.to(dev)
    block.set_x0(xd)
    func.nfe = 0
    with torch.no_grad():
      z_full = block(xd)
    parity[0] = (z_first - z_full[part.r0:part.r1]).abs().max() / z_full.abs().max()
    del z_full, xd
  dist.all_reduce(parity, op=dist.ReduceOp.MAX)

  sync = dist.barrier
  for _ in range(max(args.warmup, 3)):     # in lock-step on every rank: first replay = capture
    step_resident()
  t_end = time.perf_counter() + 1.0        # then >= 1 s more (same count everywhere: the decision is rank 0's)
  while True:
    for _ in range(5):
      step_resident()
    torch.cuda.synchronize()
    go = torch.tensor([1.0 if time.perf_counter() < t_end else 0.0], device=dev)
    dist.broadcast(go, 0)
    if float(go.item()) == 0.0:
      break

  with ClockSampler(local_rank) as clocks:
    rep = timed_repeats(step_resident, args.steps, args.repeats, sync=sync)
    step_e2e()
    rep_e2e = timed_repeats(step_e2e, args.steps, args.repeats, sync=sync)
  clock_summary = clocks.summary()
  shard.check_barriers()
  both = torch.tensor([median(rep), median(rep_e2e)], device=dev)
  dist.all_reduce(both, op=dist.ReduceOp.MAX)
  ms_step, ms_e2e = float(both[0]), float(both[1])

  # ---- per-phase split of one RHS evaluation, each phase timed on its own in lock-step (not overlapped) ----
  phases = {}
  if shard.transport == 'peer' and not args.no_others:
    lib, st = N.lib(), N.current_stream(dev)
    g_, a_ = C.byref(shard.plan.struct()), C.byref(shard.spec.struct)
    y, yb = shard.buffer('y'), shard.buffer('yb')
    ws = shard.ws
    kb = shard.k_bufs[0]
    ep = shard.epilogue(shard.buffer('x0'), f_out=shard.buffer('k1'), y_out=yb, coef_f=0.3, terms=[(y, 1.0)])
    ep.peers = C.pointer(shard.peers)
    saved_k = ws.k
    ws.k = kb.data_ptr()

    def timed_phase(fn, reps=8):
      fn()
      torch.cuda.synchronize()
      dist.barrier()
      a, b = _ev(), _ev()
      a.record()
      for _ in range(reps):
        fn()
      b.record()
      torch.cuda.synchronize()
      t = torch.tensor([a.elapsed_time(b) * 1e3 / reps], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      return float(t.item())

    phases['us_project'] = timed_phase(lambda: N.check(lib.grand_project_qk(y.data_ptr(), part.n_own, part.n_own, a_, ws.q, kb.data_ptr(), st)))
    phases['us_push_k'] = timed_phase(lambda: N.check(lib.grand_peer_push_rows(C.byref(shard.peers), kb.data_ptr(), part.n_own, shard.spec.ad_eff, st)))
    phases['us_barrier'] = timed_phase(lambda: N.check(lib.grand_peer_barrier(C.byref(shard.peers), st)))
    phases['us_edge_with_x_push'] = timed_phase(lambda: N.check(lib.grand_rhs_attention_edges(g_, a_, y.data_ptr(), C.byref(ws), C.byref(ep), st)))
    ep.peers = None
    phases['us_edge_without_push'] = timed_phase(lambda: N.check(lib.grand_rhs_attention_edges(g_, a_, y.data_ptr(), C.byref(ws), C.byref(ep), st)))
    ws.k = saved_k
    shard.check_barriers()

  summaries = [None] * world
  dist.all_gather_object(summaries, part.summary())
  launches = torch.tensor([launches_per_step], device=dev)
  dist.all_reduce(launches, op=dist.ReduceOp.SUM)
  if rank == 0:
    peak, peak_src = measured_peaks()
    proj_bytes, edge_bytes = algorithmic_bytes(n, e, d, ad, opt['add_source'])
    row_bytes = (d + ad) * 4
    halo_bytes = sum(s_['halo_rows'] for s_ in summaries) * row_bytes
    per_eval_s = ms_step * 1e-3 / evals
    line = {
      'metric': METRIC, 'value': e * heads * evals / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
      'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': workload_config(name, opt, n, e, world=world),
      'timing': {'repeats': args.repeats, 'statistic': 'max over ranks of the per-rank median of the repeats',
                 'warm_up_rule': '--warmup steps, then >= 1 s more in lock-step'},
      'rhs_evals_per_step': evals, 'us_per_rhs_eval': ms_step * 1e3 / evals,
      'parity_max_err': float(parity.item()),
      'parity_note': "rank 0's rows against the single-GPU engine on the whole graph, max-normalised; outside the timed region",
      'transport': shard.transport,
      'phases_per_rhs_eval': phases,
      'partition': {'halo_rows_per_rank': [s_['halo_rows'] for s_ in summaries],
                    'halo_fraction_of_graph_per_rank': [round(s_['halo_fraction_of_graph'], 4) for s_ in summaries],
                    'edges_per_rank': [s_['edges'] for s_ in summaries],
                    'rows_pushed_per_rank': [s_['send_rows'] for s_ in summaries],
                    'nvlink_bytes_per_rhs_eval_all_ranks': halo_bytes,
                    'note': 'the synthetic Chung-Lu graph has no community structure: every rank references most of the '
                            'graph; the exchange volume is the worst case for a node partition'},
      'e2e': {'value': e * heads * evals / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
              'h2d_bytes_per_step': int(n * d * 4), 'd2h_bytes_per_step': int(n * d * 4)},
      'gpu_launches': int(launches.item()) * args.steps, 'gpu_launches_per_step_all_ranks': int(launches.item()),
      'clocks': clock_summary,
      'roofline': {'bound': 'hbm', 'kernel': 'whole RHS evaluation across %d GPUs (projection + K push + barrier + fused '
                                             'edge kernel with the X push in its epilogue)' % world,
                   'achieved': (proj_bytes + edge_bytes) / per_eval_s / 1e9, 'peak': peak * world, 'unit': 'GB/s',
                   'frac': (proj_bytes + edge_bytes) / per_eval_s / 1e9 / (peak * world), 'traffic': None,
                   'peak_source': peak_src + ' x n_gpus',
                   'algorithmic_bytes_per_launch': proj_bytes + edge_bytes,
                   'nvlink_bytes_per_rhs_eval': halo_bytes,
                   'nvlink_time_floor_us_per_eval': max(s_['halo_rows'] for s_ in summaries) * row_bytes / 770e9 * 1e6,
                   'nvlink_note': 'floor = the busiest rank\'s inbound halo bytes / 770 GB/s (measured peer copy rate, '
                                  'B200_PROFILING.md); it bounds the evaluation from below whatever the compute does'},
    }
    emit(line)
  dist.barrier()
  shard.close()
  dist.destroy_process_group()


_RESULT_FD = None


def quiet_stdout():
  """Everything libraries print on stdout (NCCL prints its version line there under NCCL_DEBUG=VERSION) is sent to
  stderr, so that the ONE JSON line emit() writes is the only thing on the real stdout."""
  global _RESULT_FD
  sys.stdout.flush()
  _RESULT_FD = os.dup(1)
  os.dup2(2, 1)


def emit(line):
  text = json.dumps(line) + '\n'
  if _RESULT_FD is None:
    sys.stdout.write(text)
    sys.stdout.flush()
  else:
    os.write(_RESULT_FD, text.encode())


def main():
  quiet_stdout()
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='native', choices=['native', 'reference'])
  ap.add_argument('--skip-cpu', action='store_true', help='skip the CPU baseline leg (kernel A/B runs)')
  ap.add_argument('--workload', default='ogbn-arxiv', choices=sorted(WORKLOADS))
  ap.add_argument('--mode', default='forward', choices=['forward', 'train'], help='train: forward + adjoint backward of the block')
  ap.add_argument('--transport', default=None, choices=['peer', 'nccl'], help='multi-GPU halo transport (default peer)')
  ap.add_argument('--repeats', type=int, default=5, help='timed repeats of --steps steps each; the median is reported')
  ap.add_argument('--no-others', action='store_true', help='skip the other BASELINE configurations (kernel A/B runs)')
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if args.impl == 'reference':
    return run_reference(args, rank, world)
  if world > 1 or args.gpus > 1:
    return run_native_multi(args, rank, world)
  if args.mode == 'train':
    return run_native_train(args)
  return run_native_single(args)


if __name__ == '__main__':
  main()
