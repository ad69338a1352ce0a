cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_boundary_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py BASE_CFG 2>&1 | grep -E "rel-l2|Error|error" | head
B="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_s_$name.json 2> gpurun_out/r2_s_$name.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_s_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],4), round(d['value'],1), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],4))"; }
run whole2 A=1
run perseg2 VCD_BWD_WHOLE=0
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2_s_one.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_s_one.json')); print('n1', round(d['ms_per_step'],4), round(d['value'],1))"
