set -u
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 110 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -14 | tee $O/r02q2_pytest.txt
