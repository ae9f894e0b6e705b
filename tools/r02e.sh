set -u
cd /root/repo
O=gpurun_out
TAG=${1:-r02e}
V=graph-neural-pde_b200/build/variants
L=graph-neural-pde_b200/libgrand_b200.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/${TAG}_pytest.txt; cat $O/${TAG}_pytest.txt
timeout 120 ./tools/l2_probe > $O/${TAG}_l2_probe.txt 2>&1; cat $O/${TAG}_l2_probe.txt
args="base=$L window=$L,GRAND_EDGE=window"
for lib in $V/lib_*.so; do name=$(basename $lib .so); args="$args ${name#lib_}=$lib"; done
timeout 600 python tools/edge_ab.py $args > $O/${TAG}_edge_ab.txt 2> $O/${TAG}_edge_ab.err; cat $O/${TAG}_edge_ab.txt | grep -v '^{'; tail -3 $O/${TAG}_edge_ab.err
b() { # name env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
  python - "$TAG" "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/%s_bench_%s.json' % (sys.argv[1], sys.argv[2]))); r = d['roofline']
    print('%-10s ms/step %.3f us/eval %.1f edge us %.1f proj us %.1f e2e %.3f launches/step %d' % (sys.argv[2], d['ms_per_step'], d['config']['us_per_rhs_eval'], r['us_per_launch'], r['projection_us_per_launch'], d['e2e']['ms_per_step'], d['gpu_launches_per_step']))
except Exception as e:
    print(sys.argv[2], 'bench FAILED', e)
PY
}
b fused GRAND_FUSE_PROJ=1
b nofuse GRAND_FUSE_PROJ=0
b w320 GRAND_LIB=/root/repo/$V/lib_w320.so
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_stream_kernel -s 12 -c 2 -f -o $O/${TAG}_rhs_stream \
  python bench.py --steps 1 --warmup 1 --skip-cpu > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu > /dev/null 2>&1
ls -la $O/${TAG}_* | awk '{print $5, $9}'
