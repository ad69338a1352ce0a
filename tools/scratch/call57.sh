cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_options_gpu.py -q -x -k cluster 2>&1 | tail -3
for S in -1 0; do
  VCD_CONV_CLUSTER=$S timeout 300 python bench.py --steps 30 --warmup 8 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('train cluster $S', round(d['ms_per_step'],4), round(d['value'],1))"
done
for S in -1 0 2 -1 0; do
  VCD_CONV_CLUSTER=$S timeout 300 python bench.py --workload infer_10s --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('infer cluster $S', round(d['ms_per_step'],3), round(d['value'],1), [ (c['class'][:14], round(c['ms_per_step'],3), round(c['tflops'],0)) for c in d['kernel_classes'] if 'conv_c' in c['class']])"
done
