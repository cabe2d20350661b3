# single solve through the different drivers
for m in 0 1 2; do echo "mode $m"; HSDDP_SOLVE_MODE=$m python tools/profile_case.py 1 config2 4 | tail -2; done
echo "mode 2, one group, always four-warp sweep"; HSDDP_SOLVE_MODE=2 HSDDP_SWEEP_KIND=0 python tools/profile_case.py 1 config2 4 | tail -2
echo "mode 2, w1 sweep"; HSDDP_SOLVE_MODE=2 HSDDP_SWEEP_KIND=1 python tools/profile_case.py 1 config2 4 | tail -2
echo "cluster off"; HSDDP_CLUSTER_LS=0 python tools/profile_case.py 1 config2 4 | tail -2
