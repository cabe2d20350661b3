timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "solve_modes_agree or parity_at_scale_config3 or full_size or receding" 2>&1 | tail -3
for v in 1 1; do python tools/profile_case.py 16384 config3 2 | tail -1; done
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_phase|k_sweep_w1|k_lr_w1" --launch-skip 51 -c 5 python tools/profile_case.py 8192 config3 1 2>&1 | grep -E "k_lr_w1|k_phase|k_sweep|duration|inst_executed" | head -20
