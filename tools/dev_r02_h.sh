# round 2, GPU call H: asynchronous phased driver + queue order: full GPU tests, then timings
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "#### 16384 auto"
python tools/profile_case.py 16384 config3 3 | tail -2
echo "#### 16384 kinds 0 / 1 / auto, groups 4"
for v in 0 1 2; do HSDDP_PHASED_GROUPS=4 HSDDP_SWEEP_KIND=$v python tools/profile_case.py 16384 config3 2 | tail -1; done
echo "#### 2048 (strong-scaling shard): persistent without / with queue order, phased"
HSDDP_QUEUE_ORDER=0 HSDDP_SOLVE_MODE=1 python tools/profile_case.py 2048 config3 4 | tail -2
HSDDP_SOLVE_MODE=1 python tools/profile_case.py 2048 config3 4 | tail -3
HSDDP_SOLVE_MODE=2 python tools/profile_case.py 2048 config3 3 | tail -2
HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1 python tools/profile_case.py 2048 config3 3 | tail -1
echo "#### 4096: persistent with order, phased"
HSDDP_SOLVE_MODE=1 python tools/profile_case.py 4096 config3 3 | tail -2
HSDDP_SOLVE_MODE=2 python tools/profile_case.py 4096 config3 3 | tail -1
echo "#### 8192"
HSDDP_SOLVE_MODE=1 python tools/profile_case.py 8192 config3 3 | tail -1
HSDDP_SOLVE_MODE=2 python tools/profile_case.py 8192 config3 3 | tail -1
