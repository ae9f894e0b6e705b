set -u
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do
GRAND_PDL=$v timeout 200 python bench.py --skip-cpu --no-others --repeats 3 2>/dev/null > $O/r02p_bench_short_pdl$v.json
python - $v <<'PY'
import json, sys
d = json.load(open('gpurun_out/r02p_bench_short_pdl%s.json' % sys.argv[1])); r = d['roofline']
print('GRAND_PDL=%s' % sys.argv[1], 'ms/step %.3f (%s) e2e %.3f edge_us %.1f proj_us %.1f' % (d['ms_per_step'], [round(v,3) for v in d['timing']['ms_per_step_repeats']], d['e2e']['ms_per_step'], r['us_per_launch'], r['projection_us_per_launch']))
PY
done | tee $O/r02p_pdl.txt
for w in Cora Photo; do for v in 1 0; do
GRAND_PDL=$v timeout 100 python bench.py --workload $w --steps 10 --skip-cpu --no-others --repeats 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w GRAND_PDL=$v: ms/step %.3f us/eval %.1f'%(d['ms_per_step'],d['us_per_rhs_eval']))"
done; done | tee -a $O/r02p_pdl.txt
