cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --workload train_48k_b32 --steps 15 --warmup 4 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_ab_$name.json 2> gpurun_out/r2_ab_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_ab_$name.json')); print('$name', round(d['ms_per_step'],4), d['gpu_launches_per_step'])"; }
run pairbwd A=1
run nopairbwd VCD_PAIR_BWD=0
run nopair VCD_PAIR=0
run pairbwd2 A=1
run nopairbwd2 VCD_PAIR_BWD=0
