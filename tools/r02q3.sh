set -u
cd /root/repo
timeout 80 python -m pytest tests/test_gpu_parity.py -q -k "blocks_golden" --durations=3 2>&1 | tail -8
