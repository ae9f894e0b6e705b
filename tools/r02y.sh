set -u
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python tools/proj_ab.py 2>&1 | tail -20 | tee $O/r02y_proj_ab.txt
timeout 600 python -m pytest tests/test_gpu_linear.py -x -q 2>&1 | tail -4
for p in tm mma; do
GRAND_PROJ=$p timeout 200 python bench.py --skip-cpu --no-others --repeats 3 2>/dev/null > $O/r02y_bench_short_$p.json
python - $p <<'PY'
import json, sys
d = json.load(open('gpurun_out/r02y_bench_short_%s.json' % sys.argv[1])); r = d['roofline']
print('GRAND_PROJ=%s' % sys.argv[1], 'ms/step %.3f e2e %.3f edge_us %.1f proj_us %.1f frac %.3f eval_frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['us_per_launch'], r['projection_us_per_launch'], r['frac'], r['rhs_eval_frac']))
PY
done | tee $O/r02y_proj_bench.txt
