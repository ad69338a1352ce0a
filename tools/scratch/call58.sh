cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_boundary_gpu.py -q -x 2>&1 | tail -3
for i in 1 2 3; do
  timeout 300 python bench.py --steps 30 --warmup 8 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('train', round(d['ms_per_step'],4), round(d['value'],1), [ (c['class'][:14], round(c['ms_per_step'],3)) for c in d['kernel_classes'] if 'wgrad' in c['class']])"
done
