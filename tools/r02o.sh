set -u
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_backward.py -x -q -k "reference_cora" 2>&1 | tail -6
timeout 300 python bench.py --mode train --steps 3 --warmup 1 2>/dev/null > $O/r02o_bench_train.json; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_train.json')); print({k:d[k] for k in ('mode','value','ms_per_step') if k in d}, d.get('config',{}).get('workload'))"
