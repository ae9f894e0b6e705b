# A/B of edge-kernel build variants (graph-neural-pde_b200/build/variants/lib_*.so) on the default bench
set -u
cd /root/repo
V=graph-neural-pde_b200/build/variants
run() { # name, env...
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/ab4_$name.json 2> gpurun_out/ab4_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open("gpurun_out/ab4_%s.json"%n)); r=d["roofline"]
    print("%-16s ms/step %.2f edge us %.1f e2e %.2f stage %s"%(n,d["ms_per_step"],r["us_per_launch"],d["e2e"]["ms_per_step"],r.get("stage_mode")))
except Exception as e:
    print(n,"FAILED",e)
PY
}
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
run default
run nopf GRAND_LIB=/root/repo/$V/lib_nopf.so
run b4mb12 GRAND_LIB=/root/repo/$V/lib_b4mb12.so
run b4mb10 GRAND_LIB=/root/repo/$V/lib_b4mb10.so
run b4mb12nopf GRAND_LIB=/root/repo/$V/lib_b4mb12nopf.so
timeout 100 python tools/edge_probe.py 2>&1 | grep "scale 1/1.*d 128"
