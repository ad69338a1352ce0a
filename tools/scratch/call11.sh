cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_boundary_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 300 python tools/host_profile.py > gpurun_out/r2_host_profile2.txt 2>&1; head -c 1500 gpurun_out/r2_host_profile2.txt
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_j_$name.json 2> gpurun_out/r2_j_$name.err; python -c "
import json
d=json.load(open('gpurun_out/r2_j_$name.json')); print('$name', round(d['ms_per_step'],4), 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],4))"; }
run base A=1
run skip_all VCD_DEBUG_SKIP=7
