# round 2, GPU call J: linear-rollout recursion with uniform lane roles: parity, throughput, latency
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "step_level or config1_single or config3_mixed or receding or parity_at_scale_config3 or horizon_sweep" 2>&1 | tail -6
python tools/profile_case.py 16384 config3 2 | tail -1
python tools/profile_case.py 1 config1 4 | tail -1
python tools/profile_case.py 2048 config3 3 | tail -1
