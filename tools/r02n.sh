set -u
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_backward.py -x -q -k "general or attention_block or trains" 2>&1 | tail -8
