export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_w1 --launch-skip 10 -c 1 -f -o gpurun_out/r02d_w1 \
  python tools/profile_case.py 8192 config3 1 > gpurun_out/ncu_r02d.log 2>&1
tail -3 gpurun_out/ncu_r02d.log
