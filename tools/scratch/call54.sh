cd $GRAFT_REPO_ROOT
VCD_CONV_CLUSTER=1 timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py -q -x 2>&1 | tail -5
for S in 0 1 2 0 1 2; do
  VCD_CONV_CLUSTER=$S timeout 300 python bench.py --steps 30 --warmup 8 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cluster $S', round(d['ms_per_step'],4), round(d['value'],1), [ (c['class'][:14], round(c['ms_per_step'],3)) for c in d['kernel_classes'] if 'conv_c>=128' in c['class']])"
done
for S in 0 1 2; do
  VCD_CONV_CLUSTER=$S timeout 300 python bench.py --workload infer_10s --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('infer cluster $S', round(d['ms_per_step'],3), round(d['value'],1), [ (c['class'][:14], round(c['ms_per_step'],3)) for c in d['kernel_classes'] if 'conv_c>=128' in c['class']])"
done
