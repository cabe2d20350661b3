# source-level ncu capture of one round of the phased driver (prep, sweep, forward): bash tools/dev_ncu_phases.sh
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
python tools/profile_case.py 8192 config3 1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_phase --launch-skip 31 -c 3 -f -o gpurun_out/r01i_phases \
  python tools/profile_case.py 8192 config3 1 > gpurun_out/ncu_phases.log 2>&1
tail -3 gpurun_out/ncu_phases.log
