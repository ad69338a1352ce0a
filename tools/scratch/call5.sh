cd $GRAFT_REPO_ROOT
export VCD_CONV_ESMEM=0 VCD_CONV_NA_SMALL=2 VCD_CONV_UW32=0
M="gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__cycles_active.avg,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum"
VCD_GRAPHS=0 timeout 600 ncu --replay-mode range --metrics $M --clock-control none --csv --log-file gpurun_out/r2_range_nograph.csv python tools/one_step.py --steps 4 --range-last > gpurun_out/r2_range_nograph.log 2>&1
timeout 600 ncu --replay-mode range --metrics $M --clock-control none --csv --log-file gpurun_out/r2_range_graph.csv python tools/one_step.py --steps 4 --range-last > gpurun_out/r2_range_graph.log 2>&1
timeout 600 ncu --replay-mode application --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_app.csv -c 5 python tools/one_step.py --steps 2 > gpurun_out/r2_app.log 2>&1
tail -5 gpurun_out/r2_range_nograph.log; tail -30 gpurun_out/r2_range_nograph.csv | cut -c1-400; tail -5 gpurun_out/r2_range_graph.log; tail -8 gpurun_out/r2_range_graph.csv | cut -c1-300
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e --profile-classes"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_e_$name.json 2> gpurun_out/r2_e_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_e_$name.json')); print('$name', round(d['ms_per_step'],4), [ (c['class'][:8], round(c['ms_per_step'],3)) for c in d['kernel_classes']])"; }
run lr22 VCD_WGRAD_LOAD_RATE=22
run lr10 VCD_WGRAD_LOAD_RATE=10
run lr5 VCD_WGRAD_LOAD_RATE=5
run serial VCD_SERIAL=1
