# End-of-round validation on one B200: GPU tests, smoke, the default bench (CPU baseline + other configurations) and its
# reference arm, the ncu launch list of a short bench run and one full capture of the edge kernel and of the projection.
set -u
cd /root/repo
TAG=${1:-r02z}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/${TAG}_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 700 python bench.py > $O/${TAG}_bench_arxiv_n1_full.json 2> $O/${TAG}_bench.err || tail -3 $O/${TAG}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open('gpurun_out/%s_bench_arxiv_n1_full.json' % tag)); r = d['roofline']
    print('N=1 value %.3e ms/step %.3f (repeats %s) e2e %.3f edge_us %.1f proj_us %.1f frac %.3f cpu %.3e on %s threads' % (
        d['value'], d['ms_per_step'], [round(v, 3) for v in d['timing']['ms_per_step_repeats']], d['e2e']['ms_per_step'],
        r['us_per_launch'], r['projection_us_per_launch'], r['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
    for k, v in d['other_workloads'].items():
        print('   %-14s us/eval %8.1f  ms/step %.3f  frac %.4f  cpu %s' % (k, v.get('us_per_rhs_eval', -1), v.get('ms_per_step', -1), (v.get('roofline') or {}).get('frac', -1), (v.get('cpu_baseline') or {}).get('value')))
    ref = json.load(open('gpurun_out/%s_bench_reference.json' % tag))
    print('reference arm value %.3e (%s threads)' % (ref['value'], ref['cpu_baseline']['cores']))
except Exception as e:
    print('bench summary FAILED', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rhs_|linear_t|project_|lseg_|split_|segment_|edge_' -c 150 --csv --log-file $O/${TAG}_launches_arxiv.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rhs_ring_kernel -s 6 -c 1 -f -o $O/${TAG}_rhs_ring \
  python bench.py --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tm_kernel -s 6 -c 1 -f -o $O/${TAG}_linear_tm \
  python bench.py --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
ls -la $O/${TAG}_* | awk '{print $5, $9}'
