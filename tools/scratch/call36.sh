cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --dump-launches gpurun_out/r02_prof_launches_train.csv > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cpu.json 2> gpurun_out/r02_bench_reference_cpu.err
timeout 400 python bench.py --workload infer_10s --steps 5 --warmup 3 --no-extra --no-cpu-baseline --dump-launches gpurun_out/r02_prof_launches_infer.csv > gpurun_out/r02_bench_infer_10s.json 2> gpurun_out/r02_bench_infer_10s.err
timeout 400 python bench.py --workload train_48k_b32 --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_train_48k_b32.json 2> gpurun_out/r02_bench_train_48k.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_ncu_launches.csv python tools/one_step.py --steps 3 > gpurun_out/r02_ncu_launches.log 2>&1
N="timeout 600 ncu --set full --clock-control none --import-source on"
$N -k regex:pair_kernel -s 24 -c 2 -o gpurun_out/r02_prof_pair_train -f python tools/one_step.py --steps 2 > /dev/null 2>&1
$N -k regex:pair_kernel -s 29 -c 2 -o gpurun_out/r02_prof_pair_infer -f python tools/one_step.py --B 64 --T 938 --fwd-only --steps 2 > /dev/null 2>&1
$N -k regex:wgrad_kernel -s 100 -c 2 -o gpurun_out/r02_prof_wgrad -f python tools/one_step.py --steps 3 > /dev/null 2>&1
$N -k regex:conv_kernel -s 20 -c 3 -o gpurun_out/r02_prof_conv_infer_big -f python tools/one_step.py --B 64 --T 938 --fwd-only --steps 2 > /dev/null 2>&1
ls -la gpurun_out/r02_*
head -c 600 gpurun_out/r02_bench_1gpu.json
