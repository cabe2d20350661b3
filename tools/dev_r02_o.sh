# round 2, GPU call O: linear rollout fed by TMA bulk copies
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "step_level or config1_single or config3_mixed or receding or horizon_sweep or solve_modes" 2>&1 | tail -4
python tools/profile_case.py 16384 config3 3 | tail -2
python tools/profile_case.py 1 config1 4 | tail -1
python tools/profile_case.py 2048 config3 3 | tail -1
