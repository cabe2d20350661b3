# round 2, GPU call I: receding-horizon update on the device
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "receding_horizon or config1_single or config3_mixed or warm_start or device_schedule_builder or mpc_command" 2>&1 | tail -15
echo "#### 16384 (forward kernel grew a stack frame: any cost?)"
python tools/profile_case.py 16384 config3 2 | tail -1
