# End-of-round validation on one B200: GPU tests, smoke, the default bench (with CPU baseline) and its reference arm, the
# other BASELINE configs, the ncu launch list of the bench command and one full capture of the two main kernels.
# Everything lands in gpurun_out/ (scratch); the summaries that matter are copied into profiles/ by hand.
set -u
cd /root/repo
TAG=${1:-r01f}
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2 > $O/${TAG}_pytest.txt; cat $O/${TAG}_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 300 python bench.py > $O/${TAG}_bench_arxiv_n1_full.json 2> $O/${TAG}_bench.err || tail -3 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
for w in Cora Photo Cora-BLEND; do
  timeout 120 python bench.py --workload $w --skip-cpu > $O/${TAG}_bench_$w.json 2>> $O/${TAG}_bench.err
done
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
for n in ('bench_arxiv_n1_full', 'bench_reference', 'bench_Cora', 'bench_Photo', 'bench_Cora-BLEND'):
    try:
        d = json.load(open('gpurun_out/%s_%s.json' % (tag, n)))
        r = d.get('roofline') or {}
        print('%-22s value %.3e ms/step %s e2e %s edge_us %s proj_us %s frac %s cpu %s' % (
            n, d['value'], d.get('ms_per_step'), (d.get('e2e') or {}).get('ms_per_step'), r.get('us_per_launch'),
            r.get('projection_us_per_launch'), r.get('frac'), (d.get('cpu_baseline') or {}).get('value')))
    except Exception as e:
        print(n, 'FAILED', e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_launches_arxiv.csv \
  python bench.py --steps 1 --warmup 1 --skip-cpu > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rhs_window_kernel -s 6 -c 1 -f -o $O/${TAG}_rhs_window \
  python bench.py --steps 1 --warmup 1 --skip-cpu > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:project_qk -s 6 -c 1 -f -o $O/${TAG}_project \
  python bench.py --steps 1 --warmup 1 --skip-cpu > /dev/null 2>&1
ls -la $O/${TAG}_* | awk '{print $5, $9}'
