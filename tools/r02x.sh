set -u
cd /root/repo
O=gpurun_out; mkdir -p $O
GRAND_PROJ=tm timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tm_kernel -s 6 -c 1 -f -o $O/r02x_linear_tm \
  python bench.py --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
GRAND_PROJ=tc timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 6 -c 1 -f -o $O/r02x_linear_tc \
  python bench.py --steps 1 --warmup 1 --skip-cpu --no-others --repeats 1 > /dev/null 2>&1
ls -la $O/r02x_* | awk '{print $5, $9}'
