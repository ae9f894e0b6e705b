set -u
cd /root/repo
O=gpurun_out
TAG=${1:-r02q}
for w in Cora Photo Cora-GRAND-l; do
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
python - "$O/${TAG}_launches_$w.csv" "$w" <<'PY'
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hi + 2:]:
    if len(r) > vi and ('grand' in r[ki] or 'dopri5' in r[ki] or 'lincomb' in r[ki] or 'sumsq' in r[ki]):
        try: agg[r[ki][:64]].append(float(r[vi].replace(',', '')))
        except: pass
print('==', sys.argv[2])
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:9]:
    print('  %-66s n=%4d avg %7.1f us min %7.1f' % (k, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3))
PY
done
