for g in 2 4 8 16; do echo "2048 phased groups $g (min group 128)"; HSDDP_SOLVE_MODE=2 HSDDP_PHASED_MIN_GROUP=128 HSDDP_PHASED_GROUPS=$g python tools/profile_case.py 2048 config3 3 | tail -1; done
for g in 8 16; do echo "4096 phased groups $g (min group 128)"; HSDDP_SOLVE_MODE=2 HSDDP_PHASED_MIN_GROUP=128 HSDDP_PHASED_GROUPS=$g python tools/profile_case.py 4096 config3 3 | tail -1; done
echo "2048 phased groups 8, sweep kind 0"; HSDDP_SWEEP_KIND=0 HSDDP_SOLVE_MODE=2 HSDDP_PHASED_MIN_GROUP=128 HSDDP_PHASED_GROUPS=8 python tools/profile_case.py 2048 config3 3 | tail -1
