# round 2, GPU call C: warp-per-problem sweep kernel: parity in phased mode, A/B against the block-per-problem sweep
echo "#### parity (phased-mode tests)"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "solve_modes_agree or parity_at_scale_config3 or regularisation_retry" 2>&1 | tail -5
echo "#### A/B 16384"
for v in 0 1 0 1; do HSDDP_SWEEP_W1=$v python tools/profile_case.py 16384 config3 2 | tail -1; done
echo "#### A/B 4096 phased"
for v in 0 1; do HSDDP_SOLVE_MODE=2 HSDDP_SWEEP_W1=$v python tools/profile_case.py 4096 config3 2 | tail -1; done
