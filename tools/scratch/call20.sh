cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py tests/test_boundary_gpu.py tests/test_mel_gpu.py -m gpu -q -x 2>&1 | tail -5
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_p_$name.json 2> gpurun_out/r2_p_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_p_$name.json')); print('$name', round(d['ms_per_step'],4), 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],4))"; }
run whole A=1
run perseg VCD_BWD_WHOLE=0
run whole_wg148 VCD_WGRAD_CTAS_SMALL=148 VCD_WGRAD_CTAS_BIG=64
run whole_wgtail VCD_WGRAD_CTAS_TAIL=148
VCD_PHASES=1 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -6
