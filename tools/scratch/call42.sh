cd $GRAFT_REPO_ROOT
echo "== pair on"; VCD_GRAPHS=0 VCD_PHASES=1 VCD_BWD_WHOLE=0 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -13
echo "== pair off"; VCD_PAIR=0 VCD_GRAPHS=0 VCD_PHASES=1 VCD_BWD_WHOLE=0 timeout 120 python tools/one_step.py --steps 4 2>&1 | tail -13
