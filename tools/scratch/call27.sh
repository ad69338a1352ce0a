cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -4
VCD_PAIR_MT=2 timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
VCD_PAIR_MT=4 timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_v_$name.json 2> gpurun_out/r2_v_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_v_$name.json')); print('$name', round(d['ms_per_step'],4))"; }
run pair_auto A=1
run pair_mt1 VCD_PAIR_MT=1
run pair_mt2 VCD_PAIR_MT=2
run nopair VCD_PAIR=0
I="timeout 300 python bench.py --workload infer_10s --steps 3 --warmup 2 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
runi() { name=$1; shift; env "$@" $I --dump-launches gpurun_out/r2_v_$name.csv > gpurun_out/r2_v_$name.json 2> gpurun_out/r2_v_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_v_$name.json')); print('$name', round(d['ms_per_step'],3), round(d['value'],1))"; }
runi inf_auto A=1
runi inf_mt1 VCD_PAIR_MT=1
runi inf_mt2 VCD_PAIR_MT=2
runi inf_nopair VCD_PAIR=0
