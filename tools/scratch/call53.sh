cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_mel_loss_gpu.py -q -x 2>&1 | tail -3
timeout 300 python - <<'PY' 2>&1 | tail -3
import sys
sys.argv=['bench.py','--no-extra']
import bench
b = bench.Bench(bench.parse_args())
r = b.mel_tail_record(16, 16384)
print(r['ms_per_call'], r['dense_gemm_path']['ms_per_call'], r['cpu_oracle']['ms_per_call'])
PY

