cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_w_$name.json 2> gpurun_out/r2_w_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_w_$name.json')); print('$name', round(d['ms_per_step'],4))"; }
run base A=1
run na2 VCD_CONV_NA_SMALL=2
run na3 VCD_CONV_NA_SMALL=3
run na2_ne3 VCD_CONV_NA_SMALL=2 VCD_CONV_NE=3
run na2_ne2_s48 VCD_CONV_NA_SMALL=2 VCD_WGRAD_CTAS_SMALL=48
run na2_pairna2 VCD_CONV_NA_SMALL=2 VCD_PAIR_NA=2
run na2_pairmt1 VCD_CONV_NA_SMALL=2 VCD_PAIR_MT=1
run na2_smem160 VCD_CONV_NA_SMALL=2 VCD_CONV_SMEM_KB=160
run na2_wgsm64 VCD_CONV_NA_SMALL=2 VCD_WGRAD_SMEM_KB_SMALL=64
run base_again A=1
