cd $GRAFT_REPO_ROOT
for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_${N}gpu.json') if l.startswith('{')][-1]); print($N, round(d['ms_per_step'],4), round(d['value'],1), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],4))"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_1gpu_b.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu_b.json')); print(1, round(d['ms_per_step'],4), round(d['value'],1))"
