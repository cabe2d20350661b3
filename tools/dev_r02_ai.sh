# ncu --set full of one round with k_lr_w1 (prep, sweep x2, linear rollout, forward)
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
ncu --set full --clock-control none --import-source on -k regex:"k_phase|k_sweep_w1|k_lr_w1" --launch-skip 51 -c 5 -f -o gpurun_out/r02ai_round \
  python tools/profile_case.py 8192 config3 1 > gpurun_out/ncu_r02ai.log 2>&1
tail -2 gpurun_out/ncu_r02ai.log
