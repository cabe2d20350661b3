timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "solve_modes_agree or parity_at_scale_config3 or regularisation_retry or full_size_config3" 2>&1 | tail -4
python tools/profile_case.py 16384 config3 3 | tail -2
HSDDP_W1_MIN_BLOCKS=1184 python tools/profile_case.py 16384 config3 2 | tail -1
HSDDP_SOLVE_MODE=2 python tools/profile_case.py 65536 config3 2 | tail -1
