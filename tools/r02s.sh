set -u
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in Cora Cora-GRAND-l Cora-BLEND; do
for c in 1 0; do
GRAND_COOP=$c timeout 100 python bench.py --workload $w --steps 10 --skip-cpu --no-others --repeats 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w COOP=$c: ms/step %.3f evals %d us/eval %.1f launches/step %d'%(d['ms_per_step'],d['rhs_evals_per_step'],d['us_per_rhs_eval'],d['gpu_launches_per_step']))"
done; done
