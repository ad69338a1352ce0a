cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'])"
