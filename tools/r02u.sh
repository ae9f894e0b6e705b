set -u
cd /root/repo
mkdir -p gpurun_out
timeout 240 python tools/proj_ab.py 2>&1 | tail -25 | tee gpurun_out/r02u_proj_ab.txt
timeout 300 python -m pytest tests/test_gpu_backward.py -x -q -k "adaptive_adjoint" 2>&1 | tail -5
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
