cd $GRAFT_REPO_ROOT
timeout 200 ./build_tools/load_rate > gpurun_out/r2_load_rate2.txt 2>&1
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_i_$name.json 2> gpurun_out/r2_i_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_i_$name.json')); print('$name', round(d['ms_per_step'],4))"; }
run base A=1
run skip_wgrad VCD_DEBUG_SKIP=4
run skip_dgrad VCD_DEBUG_SKIP=2
run skip_fwd VCD_DEBUG_SKIP=1
run skip_bwd VCD_DEBUG_SKIP=6
run skip_all VCD_DEBUG_SKIP=7
VCD_PHASES=1 VCD_DEBUG_SKIP=4 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -8
VCD_PHASES=1 VCD_DEBUG_SKIP=2 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -8
