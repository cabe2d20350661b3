# round 2, GPU call E: two-warp sweep kernel: parity in phased mode, A/B of the three sweep kernels, ncu of the new one
echo "#### parity (phased-mode tests), sweep kind 2"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "solve_modes_agree or parity_at_scale_config3 or regularisation_retry" 2>&1 | tail -5
echo "#### A/B 16384"
for v in 0 1 2 0 1 2; do echo "kind $v"; HSDDP_SWEEP_KIND=$v python tools/profile_case.py 16384 config3 2 | tail -1; done
echo "#### ncu"
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_w2 --launch-skip 10 -c 1 -f -o gpurun_out/r02e_w2 \
  python tools/profile_case.py 8192 config3 1 > gpurun_out/ncu_r02e.log 2>&1
tail -2 gpurun_out/ncu_r02e.log
