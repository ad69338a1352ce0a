cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_mel_loss_gpu.py -q -x 2>&1 | tail -25
timeout 300 python - <<'PY' 2>&1 | tail -5
import torch, sys
sys.argv=['bench.py','--no-extra']
import bench
args = bench.parse_args()
b = bench.Bench(args)
print(b.mel_tail_record(16, 16384))
PY
