cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py tests/test_mel_gpu.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_t7.log
VCD_CONV_MT_SMALL=2 timeout 600 python -m pytest tests/test_bf16_local_parity_gpu.py -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r2_t7.log
VCD_CONV_MT_SMALL=4 VCD_CONV_NE=3 timeout 600 python -m pytest tests/test_bf16_local_parity_gpu.py -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r2_t7.log
cat gpurun_out/r2_t7.log
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
run() { name=$1; shift; env "$@" $B --dump-launches gpurun_out/r2_g_$name.csv > gpurun_out/r2_g_$name.json 2> gpurun_out/r2_g_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_g_$name.json')); print('$name', round(d['ms_per_step'],4), [ (c['class'][:8], round(c['ms_per_step'],3)) for c in d['kernel_classes']])"; }
OLD="VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2 VCD_CONV_UW32=0 VCD_CONV_LEAN=0"
run old $OLD
run lean_def A=1
run lean_ne4 VCD_CONV_NE=4
run lean_mt2 VCD_CONV_MT_SMALL=2 VCD_CONV_NE=3
run lean_mt4 VCD_CONV_MT_SMALL=4 VCD_CONV_NE=3 VCD_CONV_NA_SMALL=3
run lean_mt2_na2 VCD_CONV_MT_SMALL=2 VCD_CONV_NE=2 VCD_CONV_NA_SMALL=2
run lean_f0only VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2
run lean_f0only_mt2 VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2 VCD_CONV_MT_SMALL=2
VCD_KTRACE=resblocks.9.convs1.0:fwd timeout 120 python tools/ktrace.py 2>&1 | tail -22
