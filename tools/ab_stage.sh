# A/B of edge-kernel build variants on the default bench (run on the GPU box: gpurun -- 'bash tools/ab_stage.sh TAG').
# Times the in-tree library and every graph-neural-pde_b200/build/variants/lib_*.so (tools/build_variants.sh), after the
# GPU test-suite has passed with the in-tree library; variants are also put through the parity tests when CHECK=1.
set -u
cd /root/repo
TAG=${1:-ab}
V=graph-neural-pde_b200/build/variants
run() { # name, env...
  name=$1; shift
  env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python - "$TAG" "$name" <<'PY'
import json,sys
tag,n=sys.argv[1:3]
try:
    d=json.load(open("gpurun_out/%s_%s.json"%(tag,n))); r=d["roofline"]
    print("%-16s ms/step %.2f edge us %.1f proj us %.1f e2e %.2f"%(n,d["ms_per_step"],r["us_per_launch"],r["projection_us_per_launch"],d["e2e"]["ms_per_step"]))
except Exception as e:
    print(n,"FAILED",e)
PY
}
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
run default
for lib in $V/lib_*.so; do
  [ -e "$lib" ] || continue
  name=$(basename $lib .so); name=${name#lib_}
  if [ "${CHECK:-0}" = "1" ]; then
    GRAND_LIB=/root/repo/$lib timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1
  fi
  run $name GRAND_LIB=/root/repo/$lib
done
