cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_ad_$name.json 2> gpurun_out/r2_ad_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_ad_$name.json')); print('$name', round(d['ms_per_step'],4), d['gpu_launches_per_step'])"; }
run equal A=1
run mix50 VCD_SHARE_MIX=50
run share0 VCD_SHARE=0
run equal_all VCD_SHARE=2
