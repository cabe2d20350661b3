# round 2, GPU call K: after the register-pressure fixes: parity, throughput, latency; persistent kernel at 5 blocks / SM
P=$PWD/hkd-mpc_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "step_level or config1_single or config3_mixed or receding or parity_at_scale_config3 or horizon_sweep or solve_modes" 2>&1 | tail -4
python tools/profile_case.py 16384 config3 2 | tail -1
python tools/profile_case.py 1 config1 4 | tail -1
for v in - mb5; do
  if [ "$v" = "-" ]; then export HSDDP_LIB=$P/libhsddp_b200.so; else export HSDDP_LIB=$P/libhsddp_b200_$v.so; fi
  echo "== $v"
  python tools/profile_case.py 2048 config3 4 | tail -2
  python tools/profile_case.py 4096 config3 3 | tail -1
  HSDDP_SOLVE_MODE=1 python tools/profile_case.py 8192 config3 3 | tail -1
done
