# Re-entry validation of HEAD (+ the stream kernel's spill fix): GPU tests, ring vs stream A/B at both segment sizes,
# short bench lines for both, the full default bench line + reference arm, ncu launch list + one full capture of the
# default edge kernel and of the projection.
set -u
cd /root/repo
O=gpurun_out; mkdir -p $O
L=graph-neural-pde_b200/libgrand_b200.so
V=graph-neural-pde_b200/build/variants
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/r02t_pytest.txt
for seg in 512 256; do
GRAND_SEG_EDGES=$seg timeout 300 python tools/edge_ab.py ring=$L stream=$L,GRAND_EDGE=stream ring12=$V/lib_ring12.so ringeb1=$V/lib_ringeb1.so ringall=$L,GRAND_EDGE=ring 2>&1 | grep -v "^{" | tee -a $O/r02t_ab.txt
done
for e in default stream; do
GRAND_EDGE=$e timeout 200 python bench.py --skip-cpu --no-others --repeats 3 2>/dev/null > $O/r02t_bench_short_$e.json
python - $e <<'PY'
import json, sys
d = json.load(open('gpurun_out/r02t_bench_short_%s.json' % sys.argv[1])); r = d['roofline']
print(sys.argv[1], 'ms/step %.3f e2e %.3f edge_us %.1f proj_us %.1f frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['us_per_launch'], r['projection_us_per_launch'], r['frac']))
PY
done
timeout 600 python bench.py > $O/r02t_bench_arxiv_n1_full.json 2> $O/r02t_bench.err || tail -3 $O/r02t_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02t_bench_reference.json 2>> $O/r02t_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'grand' -c 120 --csv --log-file $O/r02t_launches_arxiv.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rhs_ring_kernel -s 6 -c 1 -f -o $O/r02t_rhs_ring \
  python bench.py --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
GRAND_EDGE=stream timeout 300 ncu --set full --clock-control none --import-source on -k regex:rhs_stream_kernel -s 6 -c 1 -f -o $O/r02t_rhs_stream \
  python bench.py --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
ls -la $O/r02t_* | awk '{print $5, $9}'
