set -u
cd /root/repo
O=gpurun_out
TAG=${1:-r02g}
V=graph-neural-pde_b200/build/variants
L=graph-neural-pde_b200/libgrand_b200.so
for md in 13161 40; do
  echo "== max degree $md"
  timeout 300 python tools/edge_ab.py --max-degree $md base=$L old320=$V/lib_r02b320.so combined=$V/lib_combined.so w320=$V/lib_w320.so 2>> $O/${TAG}.err | grep -v '^{' | tee -a $O/${TAG}_ab.txt
done
