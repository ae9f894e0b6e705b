# 2 GPUs: multi-GPU parity (both transports, eager + graph), bench N=2 (peer transport), plus the new single-GPU GAT cases
set -u
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "blocks_golden" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -12 | tee $O/r02m2_check.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null > $O/r02m2_bench_n2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m2_bench_n2.json'))
print('N=2 ms/step %.3f value %.3e parity %s'%(d['ms_per_step'], d['value'], d.get('parity_max_err')))
print({k:v for k,v in d.items() if k.startswith('us_') or 'phase' in k})
PY
