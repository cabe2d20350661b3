# new GPU tests; heavy-first on / off on the strong-scaling shards; persistent / phased crossover on the final code
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "linear_rollout_kernel or heaviest or solve_modes_agree or receding" 2>&1 | tail -5
for v in 0 1; do echo "HSDDP_HEAVY_FIRST=$v"; HSDDP_HEAVY_FIRST=$v python tools/strong_shards.py 2048 1 | cut -c1-75; done
for n in 1024 4096 6144; do for v in 0 1; do echo "n $n heavy-first $v"; HSDDP_HEAVY_FIRST=$v HSDDP_SOLVE_MODE=1 python tools/profile_case.py $n config3 3 | tail -1; done; done
for n in 4096 5120 6144; do echo "n $n phased"; HSDDP_SOLVE_MODE=2 python tools/profile_case.py $n config3 3 | tail -1; done
