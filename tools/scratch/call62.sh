cd $GRAFT_REPO_ROOT
cp vcvits_b200/libvcd.so /tmp/cur.so
for round in 1 2 3; do
for which in cur prev; do
  if [ $which = prev ]; then cp vcvits_b200/libvcd_prev.so vcvits_b200/libvcd.so; else cp /tmp/cur.so vcvits_b200/libvcd.so; fi
  timeout 300 python bench.py --steps 30 --warmup 8 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$which', round(d['ms_per_step'],4), round(d['value'],1))"
done
done
cp /tmp/cur.so vcvits_b200/libvcd.so
timeout 600 python -m pytest tests/test_bf16_local_parity_gpu.py -q -x -k "config2 or small" 2>&1 | tail -2
