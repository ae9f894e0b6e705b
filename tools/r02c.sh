set -u
cd /root/repo
O=gpurun_out
TAG=${1:-r02c}
V=graph-neural-pde_b200/build/variants
L=graph-neural-pde_b200/libgrand_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > $O/${TAG}_pytest.txt; cat $O/${TAG}_pytest.txt
args="base=$L window=$L,GRAND_EDGE=window"
for lib in $V/lib_*.so; do name=$(basename $lib .so); args="$args ${name#lib_}=$lib"; done
timeout 600 python tools/edge_ab.py $args > $O/${TAG}_edge_ab.txt 2> $O/${TAG}_edge_ab.err; cat $O/${TAG}_edge_ab.txt | grep -v '^{'; tail -3 $O/${TAG}_edge_ab.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rhs_stream_kernel -s 5 -c 1 -f -o $O/${TAG}_rhs_stream \
  python tools/edge_ab.py base=$L > /dev/null 2>&1
ls -la $O/${TAG}_* | awk '{print $5, $9}'
