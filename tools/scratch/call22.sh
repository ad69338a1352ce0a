cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_r_$name.json 2> gpurun_out/r2_r_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_r_$name.json')); print('$name', round(d['ms_per_step'],4))"; }
run base A=1
run base2 A=1
run pdl3 VCD_PDL=3
run pdl3_late3 VCD_PDL=3 VCD_PDL_LATE=3
run pdl3_late0 VCD_PDL=3 VCD_PDL_LATE=0
run occ2 VCD_CONV_OCC2=1
run wgsmem_small VCD_WGRAD_SMEM_KB=96
run wgctas_b48 VCD_WGRAD_CTAS_BIG=48
run wgctas_s48 VCD_WGRAD_CTAS_SMALL=48
run wgctas_s96 VCD_WGRAD_CTAS_SMALL=96
run na2 VCD_CONV_NA_SMALL=2
run ne3 VCD_CONV_NE=3
