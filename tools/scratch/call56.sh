cd $GRAFT_REPO_ROOT
VCD_CONV_CLUSTER=0 VCD_SERIAL=1 python tools/scratch/cl_debug.py /tmp/ref.npz 2>&1 | tail -1
for M in 6 1; do
echo "== mode $M"
VCD_CONV_CLUSTER=$M VCD_SERIAL=1 python tools/scratch/cl_debug.py /tmp/c.npz 2>&1 | tail -1
python tools/scratch/cl_cmp.py /tmp/ref.npz /tmp/c.npz 2>&1 | head -4
done
VCD_CONV_CLUSTER=1 timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py tests/test_parity_gpu.py tests/test_options_gpu.py -q -x 2>&1 | tail -3
