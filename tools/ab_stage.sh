set -u
cd /root/repo
V=graph-neural-pde_b200/build/variants
run() { # name, env...
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/ab2_$name.json 2> gpurun_out/ab2_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open("gpurun_out/ab2_%s.json"%n)); r=d["roofline"]
    print("%-16s ms/step %.2f edge us %.1f e2e %.2f stage %s"%(n,d["ms_per_step"],r["us_per_launch"],d["e2e"]["ms_per_step"],r.get("stage_mode")))
except Exception as e:
    print(n,"FAILED",e)
PY
}
GRAND_STAGE=rowbuf timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
run direct GRAND_STAGE=direct
run rowbuf GRAND_STAGE=rowbuf
run rowbuf_mb6 GRAND_STAGE=rowbuf GRAND_LIB=/root/repo/$V/lib_mb6.so
run rowbuf_b4mb12 GRAND_STAGE=rowbuf GRAND_LIB=/root/repo/$V/lib_b4mb12.so
run direct_b4mb12 GRAND_STAGE=direct GRAND_LIB=/root/repo/$V/lib_b4mb12.so
run rowbuf_b4mb8 GRAND_STAGE=rowbuf GRAND_LIB=/root/repo/$V/lib_b4mb8.so
