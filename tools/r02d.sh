set -u
cd /root/repo
O=gpurun_out
TAG=${1:-r02d}
L=graph-neural-pde_b200/libgrand_b200.so
for md in 13161 2000 300 40; do
  echo "== max degree $md"
  timeout 300 python tools/edge_ab.py --max-degree $md base=$L window=$L,GRAND_EDGE=window 2>> $O/${TAG}.err | grep -v '^{' | tee -a $O/${TAG}_hub.txt
done
