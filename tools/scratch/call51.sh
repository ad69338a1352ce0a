cd $GRAFT_REPO_ROOT
for S in 1 2 1 2 3 1 2; do
  VCD_CONV_SHARE=$S timeout 300 python bench.py --steps 30 --warmup 8 --no-extra --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('share $S', round(d['ms_per_step'],4), round(d['value'],1), d['clocks']['sm_mhz'])"
done
