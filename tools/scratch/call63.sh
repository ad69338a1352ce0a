cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-extra > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_4gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'])"
