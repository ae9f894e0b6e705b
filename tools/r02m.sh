set -u
cd /root/repo
O=gpurun_out
TAG=${1:-r02m}
NG=${2:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $O/${TAG}_check.txt 2>&1; tail -8 $O/${TAG}_check.txt
for tr in peer nccl; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 10 --warmup 3 --transport $tr > $O/${TAG}_bench_n${NG}_$tr.json 2> $O/${TAG}_bench_n${NG}_$tr.err || tail -5 $O/${TAG}_bench_n${NG}_$tr.err
python - "$O/${TAG}_bench_n${NG}_$tr.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print('N=%d %s ms/step %.3f us/eval %.1f e2e %.3f parity %.2e phases %s halo %s' % (d['n_gpus'], d['transport'], d['ms_per_step'], d['us_per_rhs_eval'], d['e2e']['ms_per_step'], d['parity_max_err'], {k: round(v, 1) for k, v in d['phases_per_rhs_eval'].items()}, d['partition']['halo_fraction_of_graph_per_rank']))
except Exception as e:
    print('FAILED', e)
PY
done
