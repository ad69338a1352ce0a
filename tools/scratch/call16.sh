cd $GRAFT_REPO_ROOT
N="timeout 600 ncu --set full --clock-control none --import-source on"
$N -k regex:conv_kernel -s 322 -c 1 -o gpurun_out/r2_prof_conv_c256k11 -f python tools/one_step.py --steps 3 > gpurun_out/r2_prof1.log 2>&1
$N -k regex:conv_kernel -s 367 -c 1 -o gpurun_out/r2_prof_conv_c32k3 -f python tools/one_step.py --steps 3 > gpurun_out/r2_prof2.log 2>&1
$N -k regex:conv_kernel -s 340 -c 1 -o gpurun_out/r2_prof_conv_c128 -f python tools/one_step.py --steps 3 > gpurun_out/r2_prof3.log 2>&1
tail -3 gpurun_out/r2_prof1.log gpurun_out/r2_prof2.log gpurun_out/r2_prof3.log
ls -la gpurun_out/*.ncu-rep
