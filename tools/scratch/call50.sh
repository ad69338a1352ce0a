cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_mel_loss_gpu.py -q -x -k sliced 2>&1 | tail -3
for L in resblocks.9.convs1.1:fwd resblocks.11.convs1.1:fwd resblocks.0.convs1.1:fwd resblocks.2.convs1.1:fwd resblocks.2.convs1.1:dgrad resblocks.5.convs1.1:fwd; do
  echo "== $L"; VCD_PAIR=0 VCD_KTRACE=$L timeout 120 python tools/ktrace.py 2>&1 | tail -28
done
