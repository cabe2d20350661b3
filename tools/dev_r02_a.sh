# round 2, GPU call A: parity with Qxx-in-P3 as the default, A/B of the 4x4-pivot elimination on top of it,
# cycle accounting (profile build) alone and under load, source-level ncu capture of one round of the phased driver
P=$PWD/hkd-mpc_b200
echo "#### parity tests (default build)"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" 2>&1 | tail -3
echo "#### A/B"
bash tools/dev_ab_lat.sh - b4 2>&1 | grep -v "^$"
echo "#### cycle accounting, one problem (latency kernel)"
HSDDP_LIB=$P/libhsddp_b200_prof.so python tools/profile_case.py 1 config1 2 | tail -19
echo "#### cycle accounting, 16384 problems (phased)"
HSDDP_LIB=$P/libhsddp_b200_prof.so python tools/profile_case.py 16384 config3 1 | tail -19
echo "#### ncu"
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
ncu --set full --clock-control none --import-source on -k regex:k_phase --launch-skip 31 -c 3 -f -o gpurun_out/r02a_phases \
  python tools/profile_case.py 8192 config3 1 > gpurun_out/ncu_r02a.log 2>&1
tail -3 gpurun_out/ncu_r02a.log
ls -la gpurun_out/
