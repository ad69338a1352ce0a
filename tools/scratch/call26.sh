cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -4
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_u_$name.json 2> gpurun_out/r2_u_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_u_$name.json')); print('$name', round(d['ms_per_step'],4))"; }
run pair A=1
run nopair VCD_PAIR=0
I="timeout 300 python bench.py --workload infer_10s --steps 3 --warmup 2 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
runi() { name=$1; shift; env "$@" $I --dump-launches gpurun_out/r2_u_$name.csv > gpurun_out/r2_u_$name.json 2> gpurun_out/r2_u_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_u_$name.json')); print('$name', round(d['ms_per_step'],3), round(d['value'],1))"; }
runi inf_pair A=1
runi inf_pair_na6 VCD_PAIR_NA=6
runi inf_pair_na3 VCD_PAIR_NA=3
