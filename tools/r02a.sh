# round 2, first GPU call: parity suite with the persistent edge kernel as default, in-process A/B of its build variants
# against the round-1 window kernel, a short bench, and ncu (launch list + one full capture of the new kernel).
set -u
cd /root/repo
O=gpurun_out
TAG=${1:-r02a}
V=graph-neural-pde_b200/build/variants
L=graph-neural-pde_b200/libgrand_b200.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/${TAG}_pytest.txt; cat $O/${TAG}_pytest.txt
args="base=$L window=$L,GRAND_EDGE=window h0=$L,GRAND_L2_HINTS=0 h1=$L,GRAND_L2_HINTS=1 h2=$L,GRAND_L2_HINTS=2"
for lib in $V/lib_*.so; do name=$(basename $lib .so); args="$args ${name#lib_}=$lib"; done
timeout 600 python tools/edge_ab.py $args > $O/${TAG}_edge_ab.txt 2> $O/${TAG}_edge_ab.err; cat $O/${TAG}_edge_ab.txt | grep -v '^{'; tail -3 $O/${TAG}_edge_ab.err
GRAND_SEG_EDGES=128 timeout 300 python tools/edge_ab.py seg128=$L > $O/${TAG}_edge_ab_seg128.txt 2>> $O/${TAG}_edge_ab.err; grep -v '^{' $O/${TAG}_edge_ab_seg128.txt
GRAND_SEG_EDGES=512 timeout 300 python tools/edge_ab.py seg512=$L > $O/${TAG}_edge_ab_seg512.txt 2>> $O/${TAG}_edge_ab.err; grep -v '^{' $O/${TAG}_edge_ab_seg512.txt
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || tail -5 $O/${TAG}_bench.err
python - "$TAG" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/%s_bench.json' % sys.argv[1])); r = d['roofline']
    print('bench ms/step %.3f us/eval %.1f edge us %.1f proj us %.1f e2e %.3f frac %.3f' % (d['ms_per_step'], d['config']['us_per_rhs_eval'], r['us_per_launch'], r['projection_us_per_launch'], d['e2e']['ms_per_step'], r['frac']))
except Exception as e:
    print('bench FAILED', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_stream_kernel -s 9 -c 1 -f -o $O/${TAG}_rhs_stream \
  python bench.py --steps 1 --warmup 1 --skip-cpu > /dev/null 2>&1
ls -la $O/${TAG}_* | awk '{print $5, $9}'
