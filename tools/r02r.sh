set -u
cd /root/repo
for w in Cora Photo Cora-GRAND-l; do
for seg in 256 64 32; do
GRAND_SEG_EDGES=$seg timeout 100 python bench.py --workload $w --steps 10 --skip-cpu --no-others --repeats 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w seg $seg: ms/step %.3f evals %d us/eval %.1f launches/step %d'%(d['ms_per_step'],d['rhs_evals_per_step'],d['us_per_rhs_eval'],d['gpu_launches_per_step']))"
done; done
