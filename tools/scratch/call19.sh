cd $GRAFT_REPO_ROOT
N="timeout 900 ncu --set full --clock-control none --import-source on"
$N -k regex:conv_kernel -s 136 -c 2 -o gpurun_out/r2_prof_inf_c32 -f python tools/one_step.py --B 64 --T 938 --fwd-only --steps 2 > gpurun_out/r2_prof4.log 2>&1
tail -n 3 gpurun_out/r2_prof4.log
