cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --dump-launches gpurun_out/r02_prof_launches_train.csv > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cpu.json 2> gpurun_out/r02_bench_reference_cpu.err
timeout 400 python bench.py --workload infer_10s --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_infer_10s.json 2> gpurun_out/r02_bench_infer_10s.err
timeout 400 python bench.py --workload train_48k_b32 --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_train_48k_b32.json 2> gpurun_out/r02_bench_train_48k.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_ncu_launches.csv python tools/one_step.py --steps 3 > gpurun_out/r02_ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_mel_launches.csv python tools/mel_step.py 3 > gpurun_out/r02_ncu_mel.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_fft_kernel -s 2 -c 1 -o gpurun_out/r02_prof_mel_fft -f python tools/mel_step.py 3 > gpurun_out/r02_prof_mel_fft.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches_per_step']); print(d['configs']['mel_loss_tail'])"
