# linear rollout as its own one-warp-per-problem kernel in the phased driver (k_lr_w1): parity tests, A/B
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for v in 0 1 0 1; do echo "HSDDP_LR_W1=$v"; HSDDP_LR_W1=$v python tools/profile_case.py 16384 config3 2 | tail -1; done
for v in 0 1; do echo "HSDDP_LR_W1=$v, 65536"; HSDDP_LR_W1=$v python tools/profile_case.py 65536 config3 1 | tail -1; done
for v in 0 1; do echo "HSDDP_LR_W1=$v, 8192 mode 2"; HSDDP_SOLVE_MODE=2 HSDDP_LR_W1=$v python tools/profile_case.py 8192 config3 2 | tail -1; done
