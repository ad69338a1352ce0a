cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -15
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
run() { name=$1; shift; env "$@" $B --dump-launches gpurun_out/r2_q_$name.csv > gpurun_out/r2_q_$name.json 2> gpurun_out/r2_q_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_q_$name.json')); print('$name', round(d['ms_per_step'],4), [ (c['class'][:8], round(c['ms_per_step'],3)) for c in d['kernel_classes']])"; }
run fuse A=1
run nofuse VCD_WG_FUSE=0
VCD_PHASES=1 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -4
