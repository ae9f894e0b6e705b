# tcgen05 projection default + FFMA2: tests, prefetch distance, FFMA2 / coefficient A/B of both edge kernels, small graphs
# ring vs stream, full bench line.
set -u
cd /root/repo
O=gpurun_out; mkdir -p $O
L=graph-neural-pde_b200/libgrand_b200.so
V=graph-neural-pde_b200/build/variants
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r02v_pytest.txt
for pf in 0 2 4; do echo "GRAND_TC_PREFETCH=$pf"; GRAND_TC_PREFETCH=$pf timeout 200 python tools/proj_ab.py --arxiv 2>&1 | tail -2; done | tee $O/r02v_prefetch.txt
timeout 400 python tools/edge_ab.py ring=$L ring_nof2=$V/lib_ring_nof2.so ring_coef=$V/lib_ring_coef.so stream=$L,GRAND_EDGE=stream stream_nof2=$V/lib_stream_nof2.so,GRAND_EDGE=stream stream_coef=$V/lib_stream_coef.so,GRAND_EDGE=stream 2>&1 | grep -v "^{" | tee $O/r02v_ab.txt
for w in Photo Cora; do for e in default stream; do
GRAND_EDGE=$e timeout 100 python bench.py --workload $w --steps 10 --skip-cpu --no-others --repeats 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w GRAND_EDGE=$e: ms/step %.3f evals %d us/eval %.1f launches/step %d'%(d['ms_per_step'],d['rhs_evals_per_step'],d['us_per_rhs_eval'],d['gpu_launches_per_step']))"
done; done | tee $O/r02v_small.txt
for p in tc mma; do
GRAND_PROJ=$p timeout 200 python bench.py --skip-cpu --no-others --repeats 3 2>/dev/null > $O/r02v_bench_short_$p.json
python - $p <<'PY'
import json, sys
d = json.load(open('gpurun_out/r02v_bench_short_%s.json' % sys.argv[1])); r = d['roofline']
print('GRAND_PROJ=%s' % sys.argv[1], 'ms/step %.3f e2e %.3f edge_us %.1f proj_us %.1f frac %.3f eval_frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['us_per_launch'], r['projection_us_per_launch'], r['frac'], r['rhs_eval_frac']))
PY
done | tee $O/r02v_proj_bench.txt
