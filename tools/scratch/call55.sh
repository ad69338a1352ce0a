cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" timeout 900 python -m pytest tests/test_bf16_local_parity_gpu.py -q -x -k "config2" 2>&1 | grep -E "AssertionError:|passed|failed" | head -3; }
run VCD_CONV_CLUSTER=8
run VCD_CONV_CLUSTER=7
