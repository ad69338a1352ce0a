cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2_ae.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_ae.json')); print(round(d['ms_per_step'],4), d['gpu_launches_per_step'])"
