cd $GRAFT_REPO_ROOT
echo "== serial fwd c1 k3 C32"; VCD_SERIAL=1 VCD_KTRACE=resblocks.9.convs1.0:fwd timeout 120 python tools/ktrace.py 2>&1 | tail -36
echo "== serial dgrad c1 k11 C32"; VCD_SERIAL=1 VCD_KTRACE=resblocks.11.convs1.1:dgrad timeout 120 python tools/ktrace.py 2>&1 | tail -36
echo "== serial fwd c2 k7 C64"; VCD_SERIAL=1 VCD_KTRACE=resblocks.7.convs2.1:fwd timeout 120 python tools/ktrace.py 2>&1 | tail -36
