cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -25 > gpurun_out/r2_t3.log
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --profile-classes"
$B > gpurun_out/r2_c_def.json 2> gpurun_out/r2_c_def.err
VCD_CONV_UW32=0 $B > gpurun_out/r2_c_uw16.json 2> gpurun_out/r2_c_uw16.err
VCD_CONV_UW32=0 VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2 $B > gpurun_out/r2_c_old.json 2> gpurun_out/r2_c_old.err
VCD_CONV_OCC2=1 $B > gpurun_out/r2_c_occ2.json 2> gpurun_out/r2_c_occ2.err
VCD_KTRACE=resblocks.9.convs1.0:fwd timeout 120 python tools/ktrace.py > gpurun_out/r2_ktrace_c32k3.log 2>&1
VCD_KTRACE=resblocks.11.convs2.0:dgrad timeout 120 python tools/ktrace.py > gpurun_out/r2_ktrace_c32k11d.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -s 255 -c 300 --csv --log-file gpurun_out/r2_launches_c.csv python tools/one_step.py --steps 2 > gpurun_out/r2_ncu_c.log 2>&1
tail -4 gpurun_out/r2_t3.log
for f in def uw16 old occ2; do python -c "
import json,sys
d=json.load(open('gpurun_out/r2_c_$f.json')); print('$f', round(d['ms_per_step'],4), [ (c['class'][:8], round(c['ms_per_step'],3)) for c in d['kernel_classes']])"; done
cat gpurun_out/r2_ktrace_c32k3.log | tail -25
