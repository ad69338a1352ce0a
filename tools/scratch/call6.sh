cd $GRAFT_REPO_ROOT
export VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2 VCD_CONV_UW32=0
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_f_$name.json 2> gpurun_out/r2_f_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_f_$name.json')); print('$name', round(d['ms_per_step'],4), [ (c['class'][:8], round(c['ms_per_step'],3)) for c in d['kernel_classes']])"; }
run base A=1
run big108 VCD_CONV_SMEM_KB_BIG=108
run big108_b1 VCD_CONV_SMEM_KB_BIG=108 VCD_CONV_BUFS1_BIG=1
run big140_b1 VCD_CONV_SMEM_KB_BIG=140 VCD_CONV_BUFS1_BIG=1
run b1 VCD_CONV_BUFS1_BIG=1
run big108_b1_wg100 VCD_CONV_SMEM_KB_BIG=108 VCD_CONV_BUFS1_BIG=1 VCD_WGRAD_SMEM_KB=100
run wg100 VCD_WGRAD_SMEM_KB=100
run wgsmall64 VCD_WGRAD_SMEM_KB_SMALL=64
run pdl3 VCD_PDL=3
run pdl0 VCD_PDL=0
VCD_PHASES=1 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -8
VCD_PHASES=1 VCD_CONV_SMEM_KB_BIG=108 VCD_CONV_BUFS1_BIG=1 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -8
