cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r2_x_default.json 2> gpurun_out/r2_x_default.err
python -c "
import json
d=json.load(open('gpurun_out/r2_x_default.json'))
print('default', d['ms_per_step'], d['value'], 'e2e', d['e2e'])
print(json.dumps(d['roofline'], indent=1))
for c in d['roofline_classes']: print(c)
print({k:(v['ms_per_step'], v['value']) for k,v in d['configs'].items()})
"
timeout 300 python bench.py --workload infer_10s --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_x_infer.json 2> gpurun_out/r2_x_infer.err
python -c "
import json
d=json.load(open('gpurun_out/r2_x_infer.json'))
print('infer', d['ms_per_step'], d['value'])
for c in d['roofline_classes']: print(c)
"
