cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -5
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
run() { name=$1; shift; env "$@" $B --dump-launches gpurun_out/r2_l_$name.csv > gpurun_out/r2_l_$name.json 2> gpurun_out/r2_l_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_l_$name.json')); print('$name', round(d['ms_per_step'],4), [ (c['class'][:8], round(c['ms_per_step'],3)) for c in d['kernel_classes']])"; }
run def A=1
run nacc2 VCD_CONV_NACC=2
run nacc2_noalt VCD_CONV_NACC=2 VCD_CONV_EPI_ALT=0 VCD_CONV_NE=2
run nacc8 VCD_CONV_NACC=8
run noalt VCD_CONV_EPI_ALT=0
run ne2 VCD_CONV_NE=2
run mt2 VCD_CONV_MT_SMALL=2
run mt2_nacc8 VCD_CONV_MT_SMALL=2 VCD_CONV_NACC=8
VCD_PHASES=1 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -8
echo "== serial fwd c1 k3 C32"; VCD_SERIAL=1 VCD_KTRACE=resblocks.9.convs1.0:fwd timeout 120 python tools/ktrace.py 2>&1 | tail -36
echo "== serial dgrad c1 k11 C32"; VCD_SERIAL=1 VCD_KTRACE=resblocks.11.convs1.1:dgrad timeout 120 python tools/ktrace.py 2>&1 | tail -36
