"""World-size-2 (and 3) gloo tests of the node-partition host logic (graph-neural-pde_b200/distributed.py): ownership
ranges, halo lists, send/receive counts, local graphs and the halo exchange itself.  The per-rank arithmetic is done
with the CPU oracle on the rank's local graph; the assembled result must equal the oracle on the whole graph."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from __graft_entry__ import load_package
from oracle import grand_oracle as O

load_package()
from graph_neural_pde_b200.distributed import Partition, balanced_row_ranges, exchange_halo  # noqa: E402
from graph_neural_pde_b200.synthetic import default_opt, make_graph  # noqa: E402


def _graph(n=400, m=1800):
  ei = make_graph(n, m, max_degree=120, seed=4)
  return torch.cat([ei, torch.stack([torch.arange(n), torch.arange(n)])], dim=1), n


def test_partition_invariants():
  ei, n = _graph()
  for world in (1, 2, 3, 5, 8):     # 8 = the node the scaling run uses
    parts = [Partition(ei, n, world, r) for r in range(world)]
    assert parts[0].bounds[0] == 0 and parts[0].bounds[-1] == n
    assert sum(p.n_own for p in parts) == n and sum(p.n_local_edges for p in parts) == ei.shape[1]
    edges = [p.n_local_edges for p in parts]
    assert max(edges) <= 1.35 * (sum(edges) / world) + 130, 'edge balance %s' % edges
    for p in parts:
      # every local column resolves to the right global id
      gl = torch.cat([torch.arange(p.r0, p.r1), torch.from_numpy(p.halo_global)])
      orig = ei[:, p.local_edge_ids]
      assert torch.equal(gl[p.local_edge_index[1]], orig[1]) and torch.equal(p.local_edge_index[0] + p.r0, orig[0])
      assert sum(p.recv_counts) == p.n_halo and p.recv_counts[p.rank] == 0 and p.send_counts[p.rank] == 0
      for q in parts:   # what p receives from q is what q sends to p
        assert p.recv_counts[q.rank] == q.send_counts[p.rank]


def test_push_lists_fill_every_halo():
  """Transport 'peer': every rank stores its boundary rows straight into the peers' halo sections.  Replayed on the host:
  after all ranks have pushed, every rank's [own | halo] buffer equals the global rows it references."""
  import numpy as np
  n = 500
  ei = make_graph(n, 3000, max_degree=80, seed=3)
  ei = torch.cat([ei, torch.stack([torch.arange(n), torch.arange(n)])], dim=1)
  for world in (2, 3, 4, 8):
    parts = [Partition(ei, n, world, r) for r in range(world)]
    x = np.arange(n, dtype=np.float64)
    bufs = [np.full(p.n_cols, -1.0) for p in parts]
    for p in parts:
      bufs[p.rank][:p.n_own] = x[p.r0:p.r1]
      assert p.push_ptr[0] == 0 and p.push_ptr[-1] == len(p.push_dst) == sum(p.send_counts)
      assert p.n_cols_all[p.rank] == p.n_cols and len(p.n_cols_all) == world
    for p in parts:
      for row in range(p.n_own):
        for t in range(int(p.push_ptr[row]), int(p.push_ptr[row + 1])):
          peer, dst = (int(v) for v in p.push_dst[t])
          assert peer != p.rank and parts[peer].n_own <= dst < parts[peer].n_cols
          assert bufs[peer][dst] == -1.0, 'a halo row is delivered exactly once'
          bufs[peer][dst] = x[p.r0 + row]
    for p in parts:
      assert np.array_equal(bufs[p.rank], np.concatenate([x[p.r0:p.r1], x[p.halo_global]]))


def test_balanced_ranges_handle_degenerate_inputs():
  assert balanced_row_ranges([0, 0, 0, 0], 2) == [0, 2, 4]
  b = balanced_row_ranges([1000, 1, 1, 1], 3)
  assert b[0] == 0 and b[-1] == 4 and all(b[i] <= b[i + 1] for i in range(3))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, out_queue):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    ei, n = _graph()
    d = 12
    torch.manual_seed(0)
    x = torch.randn(n, d)
    opt = default_opt(function='transformer', hidden_dim=d, attention_dim=8, heads=2, add_source=True)
    g = torch.Generator().manual_seed(1)
    P = {'Q.weight': torch.randn(8, d, generator=g) / d ** 0.5, 'Q.bias': torch.randn(8, generator=g) * 0.1,
         'K.weight': torch.randn(8, d, generator=g) / d ** 0.5, 'K.bias': torch.randn(8, generator=g) * 0.1}
    alpha, beta = torch.tensor(0.3), torch.tensor(-0.2)
    part = Partition(ei, n, world, rank)
    # local state: [own | halo], halo filled by the exchange
    xl = torch.zeros(part.n_cols, d)
    xl[:part.n_own] = x[part.r0:part.r1]
    send = xl[torch.from_numpy(part.send_index)] if len(part.send_index) else torch.zeros(0, d)
    exchange_halo(send.contiguous(), xl[part.n_own:], part.send_counts, part.recv_counts)
    assert torch.equal(xl[part.n_own:], x[torch.from_numpy(part.halo_global)]), 'halo rows'
    # the RHS on the local graph needs Q for own rows and K for own + halo rows; transformer_rhs computes both for
    # every local row, which is a superset -- rows >= n_own of the result are not part of the shard's output
    x0l = torch.zeros_like(xl)
    x0l[:part.n_own] = x[part.r0:part.r1]
    lei = part.local_edge_index
    att, _ = O.attention_forward(xl, lei, P, opt)
    ax = O.spmm(lei, att.mean(dim=1), part.n_cols, part.n_cols, xl)
    f_local = O.residual(ax, xl, alpha, beta, x0l, opt)[:part.n_own]
    want = O.transformer_rhs(x, ei, P, opt, alpha, beta, x)[part.r0:part.r1]
    err = float((f_local - want).abs().max() / want.abs().max())
    out_queue.put((rank, err))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_rhs_equals_global_rhs_gloo(world):
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  results = [q.get(timeout=120) for _ in range(world)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, err in results:
    assert err < 2e-6, 'rank %d: %.3e' % (rank, err)


def test_locality_order_shrinks_the_halo_on_a_graph_with_communities():
  """distributed.locality_order (label propagation) + relabel: on a planted-partition graph whose node ids were
  shuffled, the contiguous shards of `Partition` cut most edges in id order and few after the renumbering; on the
  community-free Chung-Lu graph of the bench there is nothing to find.  The renumbered problem is the same problem:
  the oracle's RHS commutes with the permutation."""
  import numpy as np
  from graph_neural_pde_b200.distributed import Partition, locality_order, relabel
  from oracle import grand_oracle as O
  rng = np.random.default_rng(0)
  n, k = 4000, 8
  comm = rng.integers(0, k, n)                                  # shuffled ids: community membership is random in id order
  src = rng.integers(0, n, 60000)
  same = rng.random(60000) < 0.995                              # 99.5 % of the edges stay inside a community
  members = [np.nonzero(comm == c)[0] for c in range(k)]
  dst = np.where(same, np.array([members[comm[s]][rng.integers(0, len(members[comm[s]]))] for s in src]),
                 rng.integers(0, n, 60000))
  ei = torch.from_numpy(np.stack([np.concatenate([src, dst, np.arange(n)]), np.concatenate([dst, src, np.arange(n)])]))
  perm, inv = locality_order(ei, n)
  assert sorted(perm.tolist()) == list(range(n)) and torch.equal(inv[perm], torch.arange(n))
  ei2 = relabel(ei, inv)
  for world in (2, 4, 8):
    before = max(Partition(ei, n, world, r).summary()['halo_fraction_of_graph'] for r in range(world))
    after = max(Partition(ei2, n, world, r).summary()['halo_fraction_of_graph'] for r in range(world))
    assert after < 0.5 * before, (world, before, after)
  # same problem: f(x)[perm] == f'(x[perm]) with the relabelled edge list (edge order preserved)
  torch.manual_seed(0)
  d, ad, h = 8, 8, 2
  x = torch.randn(n, d, dtype=torch.float64)
  opt = default_opt(hidden_dim=d, attention_dim=ad, heads=h)
  P = {'Q.weight': torch.randn(ad, d, dtype=torch.float64), 'Q.bias': torch.randn(ad, dtype=torch.float64),
       'K.weight': torch.randn(ad, d, dtype=torch.float64), 'K.bias': torch.randn(ad, dtype=torch.float64)}
  a, b = torch.tensor(0.3, dtype=torch.float64), torch.tensor(0.1, dtype=torch.float64)
  f = O.transformer_rhs(x, ei, P, opt, a, b, x)
  f2 = O.transformer_rhs(x[perm], ei2, P, opt, a, b, x[perm])
  assert float((f[perm] - f2).abs().max()) < 1e-12
