cd $GRAFT_REPO_ROOT
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_d_$name.json 2> gpurun_out/r2_d_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_d_$name.json')); print('$name', round(d['ms_per_step'],4), [ (c['class'][:8], round(c['ms_per_step'],3)) for c in d['kernel_classes']])"; }
OLD="VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2 VCD_CONV_UW32=0"
run old $OLD
run old_occ2 $OLD VCD_CONV_OCC2=1
run old_mt2 $OLD VCD_CONV_MT_SMALL=2
run old_mt2_occ2 $OLD VCD_CONV_MT_SMALL=2 VCD_CONV_OCC2=1
run e4_mt2 VCD_CONV_NE=4 VCD_CONV_NA_SMALL=3 VCD_CONV_MT_SMALL=2
run e4_mt4 VCD_CONV_NE=4 VCD_CONV_NA_SMALL=3 VCD_CONV_MT_SMALL=4
run e3_mt2_occ2 VCD_CONV_NE=3 VCD_CONV_NA_SMALL=2 VCD_CONV_MT_SMALL=2 VCD_CONV_OCC2=1
run f0uw32_mt2 VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2 VCD_CONV_MT_SMALL=2
run e6_mt2 VCD_CONV_NE=6 VCD_CONV_NA_SMALL=4 VCD_CONV_MT_SMALL=2
run old_wg12 $OLD VCD_WGRAD_CTAS_BIG=12
run old_wg20 $OLD VCD_WGRAD_CTAS_BIG=20
run old_wgs32 $OLD VCD_WGRAD_CTAS_SMALL=32
run old_wgs96 $OLD VCD_WGRAD_CTAS_SMALL=96
