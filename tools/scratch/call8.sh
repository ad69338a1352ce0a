cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_t8.log
cat gpurun_out/r2_t8.log
timeout 120 ./build_tools/load_rate > gpurun_out/r2_load_rate.txt 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 600 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2_launches_h.csv python tools/one_step.py --steps 3 > gpurun_out/r2_ncu_h.log 2>&1
tail -3 gpurun_out/r2_ncu_h.log
head -c 1500 gpurun_out/r2_bench_default.json
