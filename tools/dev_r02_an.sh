# new GPU test; persistent / phased crossover on the final code (auto switches at 6,216 problems)
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "linear_rollout_kernel" 2>&1 | tail -5
for n in 3072 4096 5120 6144; do for m in 1 2 3; do echo "n $n mode $m"; HSDDP_SOLVE_MODE=$m python tools/profile_case.py $n config3 3 | tail -1; done; done
