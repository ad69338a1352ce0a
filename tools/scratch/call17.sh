cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --workload infer_10s --steps 3 --warmup 2 --no-cpu-baseline --no-extra --no-e2e --profile-classes --dump-launches gpurun_out/r2_n_infer.csv > gpurun_out/r2_n_infer.json 2> gpurun_out/r2_n_infer.err
tail -c 1500 gpurun_out/r2_n_infer.err
