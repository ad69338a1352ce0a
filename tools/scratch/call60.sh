cd $GRAFT_REPO_ROOT
cp vcvits_b200/libvcd.so /tmp/new.so
timeout 900 python -m pytest tests/test_options_gpu.py -q -x -k cluster 2>&1 | tail -2
for round in 1 2; do
for which in new old; do
  if [ $which = old ]; then cp vcvits_b200/libvcd_precluster.so vcvits_b200/libvcd.so; else cp /tmp/new.so vcvits_b200/libvcd.so; fi
  timeout 300 python bench.py --steps 30 --warmup 8 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$which', round(d['ms_per_step'],4), round(d['value'],1))"
done
done
cp /tmp/new.so vcvits_b200/libvcd.so
