set -u
cd /root/repo
O=gpurun_out
TAG=${1:-r02p}
NG=${2:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -E "MULTI_GPU_CHECK|Error|error" | head -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 10 --warmup 3 > $O/${TAG}_bench_n${NG}.json 2> $O/${TAG}_bench_n${NG}.err || tail -5 $O/${TAG}_bench_n${NG}.err
python - "$O/${TAG}_bench_n${NG}.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print('N=%d %s ms/step %.3f us/eval %.1f e2e %.3f parity %.2e phases %s halo %s nvlink floor %.1f' % (d['n_gpus'], d['transport'], d['ms_per_step'], d['us_per_rhs_eval'], d['e2e']['ms_per_step'], d['parity_max_err'], {k: round(v, 1) for k, v in d['phases_per_rhs_eval'].items()}, d['partition']['halo_fraction_of_graph_per_rank'], d['roofline']['nvlink_time_floor_us_per_eval']))
except Exception as e:
    print('FAILED', e)
PY
