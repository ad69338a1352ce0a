cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -40 > gpurun_out/r2_t2.log
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --profile-classes"
$B > gpurun_out/r2_b_def.json 2> gpurun_out/r2_b_def.err
VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2 $B > gpurun_out/r2_b_old.json 2> gpurun_out/r2_b_old.err
VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=6 $B > gpurun_out/r2_b_na6.json 2> gpurun_out/r2_b_na6.err
VCD_CONV_ESMEM=1 VCD_CONV_NA_SMALL=6 VCD_CONV_NE=3 $B > gpurun_out/r2_b_e3.json 2> gpurun_out/r2_b_e3.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -s 270 -c 290 --csv --log-file gpurun_out/r2_launches_b.csv python tools/one_step.py --steps 2 > gpurun_out/r2_ncu_b.log 2>&1
tail -6 gpurun_out/r2_t2.log
for f in def old na6 e3; do python -c "
import json,sys
d=json.load(open('gpurun_out/r2_b_$f.json')); print('$f', round(d['ms_per_step'],4), [ (c['class'][:8], round(c['ms_per_step'],3)) for c in d['kernel_classes']])"; done
