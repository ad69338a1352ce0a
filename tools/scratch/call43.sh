cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_ac_$name.json 2> gpurun_out/r2_ac_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_ac_$name.json')); print('$name', round(d['ms_per_step'],4), d['gpu_launches_per_step'])"; }
run share1 A=1
run share0 VCD_SHARE=0
run share2 VCD_SHARE=2
run share1_nopair VCD_PAIR=0
run share1_nopairbwd VCD_PAIR_BWD=0
run share1_mt2 VCD_PAIR_MT=2
run share1_mt1 VCD_PAIR_MT=1
echo "== share1"; VCD_GRAPHS=0 VCD_PHASES=1 VCD_BWD_WHOLE=0 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -11
