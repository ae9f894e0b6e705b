"""Multi-GPU parity check, run under torchrun on N >= 2 GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank builds the same seeded graph and parameters, integrates its shard with the node-partitioned path
(graph-neural-pde_b200/distributed.py) and compares its rows (a) with the single-GPU engine run on the whole graph on
the same device and (b) with the CPU oracle.  Exit code 0 and a 'MULTI_GPU_CHECK OK' line on success."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from graph_neural_pde_b200.block_constant import ConstantODEblock  # noqa: E402
from graph_neural_pde_b200.distributed import ShardedTransformerRHS  # noqa: E402
from graph_neural_pde_b200.function_transformer_attention import ODEFuncTransformerAtt  # noqa: E402
from graph_neural_pde_b200.synthetic import Data, default_opt, make_graph  # noqa: E402
from oracle import grand_oracle as O  # noqa: E402


def main():
  rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
  dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', rank)))
  torch.cuda.set_device(dev)
  dist.init_process_group('nccl', device_id=dev)
  n, d = 3000, 128
  edge_index = make_graph(n, 20000, max_degree=400, seed=11)
  torch.manual_seed(3)
  x = torch.randn(n, d)
  opt = default_opt(function='transformer', block='constant', hidden_dim=d, attention_dim=16, heads=4, method='rk4',
                    step_size=0.5, time=1.5, add_source=True, max_nfe=100000)
  data = Data(x=x, edge_index=edge_index, edge_attr=None, num_nodes=n)
  block = ConstantODEblock(ODEFuncTransformerAtt, (), opt, data, dev, t=torch.tensor([0, opt['time']])).to(dev)
  g = torch.Generator().manual_seed(5)
  with torch.no_grad():
    for p in block.parameters():
      if p.dim() == 2:
        p.copy_((2.0 * torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5).to(dev))
    block.odefunc.beta_train.fill_(0.25)
  block.eval()
  xd = x.to(dev)
  with torch.no_grad():
    block.set_x0(xd)
    z_full = block(xd)
  err_single = 0.0
  for transport in ('peer', 'nccl'):
    shard = ShardedTransformerRHS(block.odefunc, n, dev, transport=transport)
    part = shard.part
    with torch.no_grad():
      for it in range(4):   # eager, capture + replay, replay, replay with another input: buffers and the graph are reusable
        xin = xd if it < 3 else 0.5 * xd
        z_shard = shard.forward_block(xin[part.r0:part.r1], xd[part.r0:part.r1], opt).clone()
        if it == 2:
          z_keep = z_shard
    torch.cuda.synchronize()
    shard.check_barriers()
    err_t = float((z_keep - z_full[part.r0:part.r1]).abs().max() / z_full.abs().max())
    print('rank %d transport %s: err vs single-GPU %.2e' % (rank, transport, err_t), flush=True)
    err_single = max(err_single, err_t)
    z_shard = z_keep
    if transport == 'peer':
      shard.close()
  # CPU oracle on rank 0 only (it is the slow leg)
  err_oracle = 0.0
  if rank == 0:
    sd = {k: v.detach().cpu() for k, v in block.state_dict().items()}
    ei, ew = O.constant_block_graph(edge_index, None, n, opt, torch.float32)
    P = {k[len('odefunc.multihead_att_layer.'):]: v for k, v in sd.items()
         if k.startswith('odefunc.multihead_att_layer.')}
    f = lambda t, y: O.transformer_rhs(y, ei, P, opt, sd['odefunc.alpha_train'], sd['odefunc.beta_train'], x)
    want = O.block_forward(f, x, opt['time'], opt)
    err_oracle = float((z_shard.cpu() - want[part.r0:part.r1]).abs().max() / want.abs().max())
  errs = torch.tensor([err_single, err_oracle], device=dev)
  dist.all_reduce(errs, op=dist.ReduceOp.MAX)
  info = part.summary()
  print('rank %d: rows %d edges %d halo %d (%.1f%% of graph) err vs single-GPU %.2e' %
        (rank, info['rows'], info['edges'], info['halo_rows'], 100 * info['halo_fraction_of_graph'], err_single),
        flush=True)
  ok = float(errs[0]) < 1e-5 and float(errs[1]) < 1e-5
  dist.barrier()
  if rank == 0:
    print('MULTI_GPU_CHECK %s  max err vs single-GPU %.2e, vs oracle %.2e (world %d)' %
          ('OK' if ok else 'FAILED', float(errs[0]), float(errs[1]), world), flush=True)
  dist.destroy_process_group()
  sys.exit(0 if ok else 1)


if __name__ == '__main__':
  main()
