# source-level ncu capture of one round of the phased driver on the final code (prep, w1 sweep, forward)
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
python tools/profile_case.py 8192 config3 1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_phase|k_sweep_w1" --launch-skip 41 -c 4 -f -o gpurun_out/r02x_round \
  python tools/profile_case.py 8192 config3 1 > gpurun_out/ncu_r02x.log 2>&1
tail -3 gpurun_out/ncu_r02x.log
