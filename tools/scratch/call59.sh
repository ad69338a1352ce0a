cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --dump-launches gpurun_out/r02_prof_launches_train.csv > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches_per_step'], d['roofline']['dominant_class'], d['roofline']['frac']); print(d['configs']['mel_loss_tail']['ms_per_call'], d['configs']['infer_10s']['value'], d['configs']['train_48k_b32']['value'])"
