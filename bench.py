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


def build_workload(name, device, seed=0):
  from __graft_entry__ import load_package
  load_package()
  from graph_neural_pde_b200.synthetic import SyntheticDataset, baseline_opt
  ds = SyntheticDataset({'Cora-BLEND': 'Cora'}.get(name, name), seed=seed, num_features=8)   # features: not on the path
  key = {'ogbn-arxiv': 'C5', 'Cora': 'C2', 'Photo': 'C3', 'Cora-BLEND': 'C4'}[name]
  opt = baseline_opt(key)
  n = ds.data.num_nodes
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(n, opt['hidden_dim'], generator=g)
  return ds, opt, x


def init_trained_like(block, seed=0):
  """'trained-like' parameters (SURVEY.md 8(d) regime ii): W ~ N(0, 1/d), alpha_train = beta_train = 0."""
  g = torch.Generator().manual_seed(seed)
  with torch.no_grad():
    for p in block.parameters():
      if p.dim() == 2:
        p.copy_((torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5).to(p.device))


def oracle_func(block_state, opt, edge_index, n, x0):
  from oracle import grand_oracle as O
  ei, ew = O.constant_block_graph(edge_index, None, n, opt, torch.float32)
  alpha, beta = block_state['odefunc.alpha_train'], block_state['odefunc.beta_train']
  prefix = 'odefunc.multihead_att_layer.'
  P = {k[len(prefix):]: v for k, v in block_state.items() if k.startswith(prefix)}
  return lambda t, y: O.transformer_rhs(y, ei, P, opt, alpha, beta, x0)


def time_cpu_reference(func, x, h, warmup, steps, budget_s=200.0):
  """Times `steps` samples of the CPU oracle after `warmup` untimed ones -> (seconds per sample, RHS evals per sample).
  A sample is one RK4 (3/8 rule) step = 4 RHS evals; when (warmup + steps) of those would not fit `budget_s` of wall
  clock (one probe evaluation decides) a sample is bounded to ONE RHS evaluation, so that the run ends within minutes
  whatever --steps the caller asks for."""
  from oracle import grand_oracle as O

  with torch.no_grad():
    t0 = time.perf_counter()
    func(0.0, x)                                   # probe (also first-touch warm-up of the thread pool / allocator)
    t_eval = time.perf_counter() - t0
    evals = 4 if (warmup + steps) * 4 * t_eval <= budget_s else 1

    def sample():
      return x + O.rk4_alt_step(func, 0.0, h, h, x) if evals == 4 else func(0.0, x)

    for _ in range(warmup):
      sample()
    t0 = time.perf_counter()
    for _ in range(steps):
      sample()
    dt = time.perf_counter() - t0
  return dt / steps, evals


def run_reference(args, rank, world):
  if rank != 0:
    return
  name = args.workload
  ds, opt, x = build_workload(name, 'cpu')
  from __graft_entry__ import load_package
  load_package()
  from graph_neural_pde_b200.block_constant import ConstantODEblock
  from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt
  from graph_neural_pde_b200.synthetic import Data
  n = ds.data.num_nodes
  data = Data(x=x, edge_index=ds.data.edge_index, edge_attr=None, num_nodes=n)
  block = ConstantODEblock(ODEFuncTransformerAtt, (), opt, data, torch.device('cpu'), t=torch.tensor([0, opt['time']]))
  init_trained_like(block)
  state = {k: v.detach() for k, v in block.state_dict().items()}
  func = oracle_func(state, opt, ds.data.edge_index, n, x)
  threads = os.cpu_count() or 1
  torch.set_num_threads(threads)
  e = int(block.odefunc.edge_index.shape[1])
  heads = opt['heads']
  sec, evals = time_cpu_reference(func, x, float(opt['step_size']), args.warmup, args.steps)
  value = e * heads * evals / sec
  sample = ('1 rk4 (3/8) step = 4 RHS evals' if evals == 4 else
            '1 RHS eval (bounded: an rk4 step per sample would not fit the time budget)')
  line = {
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
    'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong',
    'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': workload_config(name, opt, n, e, extra={'sample': sample + ' per step'}),
    'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                     'sample': 'each step = ' + sample + ' of the oracle restatement (torch CPU ops, all host '
                               'threads); the reference itself cannot be imported: its third-party wheels are absent'},
    'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  emit(line)


def workload_config(name, opt, n, e, extra=None):
  cfg = {'workload': '%s GRAND-nl (transformer attention, %d heads, %s, T=%s, h=%s%s), synthetic Chung-Lu graph with '
                     'the dataset\'s N/E/max-degree' % (name, opt['heads'], opt['method'], opt['time'],
                                                        opt['step_size'], ', add_source' if opt['add_source'] else ''),
         'nodes': n, 'edges_incl_self_loops': e, 'hidden_dim': opt['hidden_dim'], 'attention_dim': opt['attention_dim'],
         'heads': opt['heads']}
  cfg.update(extra or {})
  return cfg


def run_native_single(args):
  from __graft_entry__ import load_package
  load_package()
  import ctypes as C
  from graph_neural_pde_b200 import _native as N
  from graph_neural_pde_b200.block_constant import ConstantODEblock
  from graph_neural_pde_b200.engine import fixed_grid_dts
  from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt
  from graph_neural_pde_b200.odeint import NativeProblem
  from graph_neural_pde_b200.synthetic import Data

  dev = torch.device('cuda:0')
  torch.cuda.set_device(dev)
  name = args.workload
  ds, opt, x_host = build_workload(name, dev)
  n, d = x_host.shape
  data = Data(x=x_host, edge_index=ds.data.edge_index, edge_attr=None, num_nodes=n)
  block = ConstantODEblock(ODEFuncTransformerAtt, (), opt, data, dev, t=torch.tensor([0, opt['time']])).to(dev)
  init_trained_like(block)
  block.eval()
  func = block.odefunc
  e = int(func.edge_index.shape[1])
  heads, ad = opt['heads'], opt['attention_dim']
  dts, _ = fixed_grid_dts(0.0, opt['time'], opt['step_size'])

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
  block.set_x0(x_dev)
  prob = NativeProblem(func, x_dev)
  eng = prob.engine
  ybuf = eng.buffer('y')
  ybuf.copy_(x_dev)

  for _ in range(max(args.warmup, 3)):
    step_resident()
  torch.cuda.synchronize()

  ev = lambda: torch.cuda.Event(enable_timing=True)
  with ClockSampler(0) as clocks:
    # ---- timed region 1: inputs resident in HBM ----
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(args.steps):
      z = step_resident()
    b.record()
    torch.cuda.synchronize()
    ms_step = a.elapsed_time(b) / args.steps
    # ---- timed region 2: end to end through the block API from pinned host buffers ----
    # (a) serial: H2D -> block(x) -> D2H one after the other on one stream
    for _ in range(2):
      step_e2e()
    torch.cuda.synchronize()
    a2, b2 = ev(), ev()
    a2.record()
    for _ in range(args.steps):
      step_e2e()
    b2.record()
    torch.cuda.synchronize()
    ms_e2e_serial = a2.elapsed_time(b2) / args.steps
    # (b) pipelined: the same per-step work (H2D of x, block(x), D2H of z, all inside the timed region), double
    #     buffered so that the copies of neighbouring steps overlap the compute on the two copy engines
    cur = torch.cuda.current_stream()
    h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
    xin = [x_dev, torch.empty_like(x_dev)]
    zout = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    zpin = [z_pinned, torch.empty_like(z_pinned).pin_memory()]
    in_ready = [torch.cuda.Event(), torch.cuda.Event()]
    in_free = [torch.cuda.Event(), torch.cuda.Event()]
    out_ready = [torch.cuda.Event(), torch.cuda.Event()]
    out_free = [torch.cuda.Event(), torch.cuda.Event()]

    def pipelined(count):
      for b_ in (0, 1):
        in_free[b_].record(cur)
        out_free[b_].record(d2h)
      for it in range(count):
        b_ = it & 1
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
      cur.wait_stream(d2h)

    pipelined(2)
    torch.cuda.synchronize()
    a5, b5 = ev(), ev()
    a5.record()
    pipelined(args.steps)
    b5.record()
    torch.cuda.synchronize()
    ms_e2e = a5.elapsed_time(b5) / args.steps
    # ---- dominant kernel on its own: the fused edge kernel with the four rk4 stage epilogues ----
    ws = eng.workspace(prob.spec)
    n_planes = 5
    stage = eng.buffer('stages', rows=n_planes * n)
    yb, yc, k1, k2, k3 = (stage[i * n:(i + 1) * n] for i in range(5))
    h = dts[0]
    third = 1.0 / 3.0
    stages = [
      (ybuf, dict(f_out=k1, y_out=yb, coef_f=h * third, terms=[(ybuf, 1.0)])),
      (yb, dict(f_out=k2, y_out=yc, coef_f=h, terms=[(ybuf, 1.0), (k1, -h * third)])),
      (yc, dict(f_out=k3, y_out=yb, coef_f=h, terms=[(ybuf, 1.0), (k1, h), (k2, -h)])),
      (yb, dict(y_out=yc, coef_f=h * 0.125, terms=[(ybuf, 1.0), (k1, h * 0.125), (k2, 3 * h * 0.125),
                                                     (k3, 3 * h * 0.125)])),
    ]
    ybuf.copy_(x_dev)
    eps = [(src, prob.epilogue(**kw)) for src, kw in stages]
    lib = N.lib()
    gs, as_ = C.byref(eng.plan.struct()), C.byref(prob.spec.struct)
    stream = N.current_stream(dev)
    N.check(lib.grand_project_qk(ybuf.data_ptr(), n, n, as_, ws.q, ws.k, stream))

    def timed_graph(body, count):
      """Replay `body` from a CUDA Graph (so host launch latency is not what is measured) and time it with events on
      the replay stream; returns ms per launch group."""
      body()
      torch.cuda.synchronize()
      graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(graph):
        st = N.current_stream(dev)
        body(st)
      graph.replay()
      torch.cuda.synchronize()
      e0, e1 = ev(), ev()
      e0.record()
      graph.replay()
      e1.record()
      torch.cuda.synchronize()
      return e0.elapsed_time(e1) / count

    reps = 6

    def edge_body(st=None):
      st = stream if st is None else st
      for _ in range(reps):
        for src, ep in eps:
          N.check(lib.grand_rhs_attention_edges(gs, as_, src.data_ptr(), C.byref(ws), C.byref(ep), st))

    def proj_body(st=None):
      st = stream if st is None else st
      for _ in range(reps * 4):
        N.check(lib.grand_project_qk(ybuf.data_ptr(), n, n, as_, ws.q, ws.k, st))

    ms_edge_group = timed_graph(edge_body, reps * len(eps))   # fused kernel + its split-stats pre-pass
    ms_proj = timed_graph(proj_body, reps * 4)
  clock_summary = clocks.summary()

  peak, peak_src = measured_peaks()
  proj_bytes, edge_bytes = algorithmic_bytes(n, e, d, ad, opt['add_source'])
  achieved = edge_bytes / (ms_edge_group * 1e-3) / 1e9
  traffic = None
  tpath = os.path.join(ROOT, 'profiles', 'dominant_kernel_traffic.json')
  if os.path.exists(tpath):
    try:
      traffic = json.load(open(tpath)).get('dram_bytes_per_launch')
    except Exception:
      traffic = None
  working_set_mb = (4.0 * n * d * (1 + 1 + 5 + 1) + 8.0 * n * ad) / 1e6

  # ---- CPU baseline: the oracle on this box's host cores, bounded sample ----
  state = {k: v.detach().cpu() for k, v in block.state_dict().items()}
  cpu_func = oracle_func(state, opt, ds.data.edge_index, n, x_host)
  threads = os.cpu_count() or 1
  torch.set_num_threads(threads)
  if args.skip_cpu:
    sec = float('nan')
  else:
    sec, cpu_evals = time_cpu_reference(cpu_func, x_host, float(opt['step_size']), 1, 2, budget_s=60.0)
  cpu_value = None if args.skip_cpu else e * heads * cpu_evals / sec

  value = e * heads * evals / (ms_step * 1e-3)
  line = {
    'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': max(args.warmup, 3),
    'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
    'data': 'synthetic',
    'config': workload_config(name, opt, n, e, extra={
      'rhs_evals_per_step': evals, 'us_per_rhs_eval': ms_step * 1e3 / evals,
      'l2_policy': 'no flush: per-step working set %.0f MB (state, x0, 5 stage buffers, Q/K) > %.0f MB L2'
                   % (working_set_mb, L2_MB),
      'parallelism': 'single GPU', 'weights': 'trained-like N(0,1/d), seed 0'}),
    'e2e': {'value': e * heads * evals / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
            'h2d_bytes_per_step': int(x_pinned.numel() * 4), 'd2h_bytes_per_step': int(z_pinned.numel() * 4),
            'mode': 'double-buffered: every step copies x from pinned host memory, runs block(x) and copies z back; '
                    'the copies of step i+-1 overlap the compute of step i on the copy engines',
            'serial_ms_per_step': ms_e2e_serial, 'serial_value': e * heads * evals / (ms_e2e_serial * 1e-3)},
    'gpu_launches': launches_per_step * args.steps,
    'gpu_launches_per_step': launches_per_step,
    'clocks': clock_summary,
    'roofline': {'bound': 'hbm', 'kernel': 'rhs_kernel<MODE_FUSED,4,1> (+ split_tile_stats pre-pass)',
                 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                 'peak_source': peak_src, 'algorithmic_bytes_per_launch': edge_bytes,
                 'note': 'algorithmic bytes follow SURVEY.md 8(d) (RHS only); the kernel additionally streams the rk4 '
                         'stage-combine operands the reference moves in separate axpy kernels (avg 2.5 reads + 1.75 '
                         'writes of [N,d] per launch), see fused_stream_bytes',
                 'fused_stream_bytes': int(4.25 * n * d * 4),
                 'frac_incl_fused_streams': (edge_bytes + 4.25 * n * d * 4) / (ms_edge_group * 1e-3) / 1e9 / peak,
                 'us_per_launch': ms_edge_group * 1e3, 'projection_us_per_launch': ms_proj * 1e3,
                 'projection_algorithmic_bytes': proj_bytes,
                 'projection_achieved_GBs': proj_bytes / (ms_proj * 1e-3) / 1e9,
                 'rhs_eval_algorithmic_bytes': proj_bytes + edge_bytes,
                 'rhs_eval_achieved_GBs': (proj_bytes + edge_bytes) / (ms_step * 1e-3 / evals) / 1e9,
                 'stage_mode': os.environ.get('GRAND_STAGE', 'default')},
    'cpu_baseline': {'value': cpu_value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                     'sample': None if args.skip_cpu else
                     '2 timed samples of %s of the oracle (torch CPU ops) after 1 warm-up'
                     % ('1 rk4 step = 4 RHS evals' if cpu_evals == 4 else '1 RHS eval'),
                     'ms_per_rhs_eval': None if args.skip_cpu else sec * 1e3 / cpu_evals},
  }
  emit(line)


def run_native_multi(args, rank, world):
  """N > 1: 1-D node partition of the same workload over the ranks (strong scaling), halo exchange of boundary rows
  every RHS evaluation over NCCL.  Timing: barrier + synchronize on both sides, CUDA events, MAX over ranks."""
  import torch.distributed as dist
  from __graft_entry__ import load_package
  load_package()
  from graph_neural_pde_b200 import _native as N
  from graph_neural_pde_b200.block_constant import ConstantODEblock
  from graph_neural_pde_b200.distributed import ShardedTransformerRHS
  from graph_neural_pde_b200.engine import fixed_grid_dts
  from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt
  from graph_neural_pde_b200.synthetic import Data

  local_rank = int(os.environ.get('LOCAL_RANK', rank))
  dev = torch.device('cuda', local_rank)
  torch.cuda.set_device(dev)
  if not dist.is_initialized():
    dist.init_process_group('nccl', device_id=dev)
  name = args.workload
  ds, opt, x_host = build_workload(name, dev)      # same seed on every rank: replicated graph and parameters
  n, d = x_host.shape
  data = Data(x=x_host, edge_index=ds.data.edge_index, edge_attr=None, num_nodes=n)
  block = ConstantODEblock(ODEFuncTransformerAtt, (), opt, data, dev, t=torch.tensor([0, opt['time']])).to(dev)
  init_trained_like(block)
  block.eval()
  func = block.odefunc
  e = int(func.edge_index.shape[1])
  heads, ad = opt['heads'], opt['attention_dim']
  dts, _ = fixed_grid_dts(0.0, opt['time'], opt['step_size'])
  evals = 4 * len(dts)
  shard = ShardedTransformerRHS(func, n, dev)
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
  step_resident()
  launches_per_step = int(N.lib().grand_launch_count() - before)
  for _ in range(max(args.warmup, 3)):
    step_resident()

  def timed(fn):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
      fn()
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

  with ClockSampler(local_rank) as clocks:
    ms_step = timed(step_resident)
    step_e2e()
    ms_e2e = timed(step_e2e)
  clock_summary = clocks.summary()

  summaries = [None] * world
  dist.all_gather_object(summaries, part.summary())
  launches = torch.tensor([launches_per_step], device=dev)
  dist.all_reduce(launches, op=dist.ReduceOp.SUM)
  if rank == 0:
    peak, peak_src = measured_peaks()
    proj_bytes, edge_bytes = algorithmic_bytes(n, e, d, ad, opt['add_source'])
    halo_bytes = sum(s['halo_rows'] for s in summaries) * d * 4
    per_eval_s = ms_step * 1e-3 / evals
    line = {
      'metric': METRIC, 'value': e * heads * evals / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': world,
      'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
      'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': workload_config(name, opt, n, e, extra={
        'rhs_evals_per_step': evals, 'us_per_rhs_eval': ms_step * 1e3 / evals,
        'parallelism': 'np%d: 1-D node partition balanced by edges, halo rows exchanged per RHS eval '
                       '(NCCL all-to-all-v)' % world,
        'halo_rows_per_rank': [s['halo_rows'] for s in summaries],
        'halo_fraction_of_graph_per_rank': [round(s['halo_fraction_of_graph'], 4) for s in summaries],
        'edges_per_rank': [s['edges'] for s in summaries],
        'halo_bytes_per_rhs_eval_all_ranks': halo_bytes,
        'l2_policy': 'no flush: per-rank working set >= %.0f MB > %.0f MB L2'
                     % (4.0 * (n / world) * d * 8 / 1e6 + halo_bytes / world / 1e6, L2_MB),
        'weights': 'trained-like N(0,1/d), seed 0'}),
      'e2e': {'value': e * heads * evals / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
              'h2d_bytes_per_step': int(n * d * 4), 'd2h_bytes_per_step': int(n * d * 4)},
      'gpu_launches': int(launches.item()) * args.steps, 'gpu_launches_per_step_all_ranks': int(launches.item()),
      'clocks': clock_summary,
      'roofline': {'bound': 'hbm', 'kernel': 'whole RHS evaluation across %d GPUs (projection + fused edge kernel + '
                                             'halo pack/exchange)' % world,
                   'achieved': (proj_bytes + edge_bytes) / per_eval_s / 1e9, 'peak': peak * world, 'unit': 'GB/s',
                   'frac': (proj_bytes + edge_bytes) / per_eval_s / 1e9 / (peak * world), 'traffic': None,
                   'peak_source': peak_src + ' x n_gpus',
                   'algorithmic_bytes_per_launch': proj_bytes + edge_bytes,
                   'nvlink_bytes_per_rhs_eval': halo_bytes,
                   'nvlink_time_floor_us_per_eval': max(s['halo_rows'] for s in summaries) * d * 4 / 770e9 * 1e6},
    }
    emit(line)
  dist.barrier()
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
  ap.add_argument('--workload', default='ogbn-arxiv', choices=['ogbn-arxiv', 'Cora', 'Photo', 'Cora-BLEND'])
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if args.impl == 'reference':
    return run_reference(args, rank, world)
  if world > 1 or args.gpus > 1:
    return run_native_multi(args, rank, world)
  return run_native_single(args)


if __name__ == '__main__':
  main()
