for w in 800 1036 1554 2072 3000; do echo "w1_min_blocks $w"; HSDDP_W1_MIN_BLOCKS=$w python tools/profile_case.py 16384 config3 2 | tail -1; done
for g in 3 5 6; do echo "groups $g"; HSDDP_PHASED_GROUPS=$g python tools/profile_case.py 16384 config3 2 | tail -1; done
echo "persistent with queue order at 16384"; HSDDP_SOLVE_MODE=1 python tools/profile_case.py 16384 config3 3 | tail -1
echo "phased at 6144 / persistent at 6144"; HSDDP_SOLVE_MODE=2 python tools/profile_case.py 6144 config3 3 | tail -1; HSDDP_SOLVE_MODE=1 python tools/profile_case.py 6144 config3 3 | tail -1
