cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py tests/test_options_gpu.py -m gpu -q -x 2>&1 | tail -4
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_z_$name.json 2> gpurun_out/r2_z_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_z_$name.json')); print('$name', round(d['ms_per_step'],4), d['gpu_launches_per_step'])"; }
run pair A=1
I="timeout 300 python bench.py --workload infer_10s --steps 4 --warmup 2 --no-cpu-baseline --no-extra --no-e2e"
runi() { name=$1; shift; env "$@" $I > gpurun_out/r2_z_$name.json 2> gpurun_out/r2_z_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_z_$name.json')); print('$name', round(d['ms_per_step'],3), round(d['value'],1))"; }
runi inf_pair A=1
runi inf_pair_na2 VCD_PAIR_NA=2
runi inf_pair_na4 VCD_PAIR_NA=4
runi inf_nopair VCD_PAIR=0
