echo "#### carve-out unified"
for v in 0 1 2; do echo "kind $v"; HSDDP_SWEEP_KIND=$v python tools/profile_case.py 16384 config3 2 | tail -1; done
echo "#### carve-out default"
for v in 0 1; do echo "kind $v"; HSDDP_NO_CARVEOUT=1 HSDDP_SWEEP_KIND=$v python tools/profile_case.py 16384 config3 2 | tail -1; done
echo "#### groups, kind 1"
for g in 1 2 4 16; do echo "groups $g"; HSDDP_PHASED_GROUPS=$g HSDDP_SWEEP_KIND=1 python tools/profile_case.py 16384 config3 2 | tail -1; done
echo "#### groups, kind 0"
for g in 1 4; do echo "groups $g"; HSDDP_PHASED_GROUPS=$g HSDDP_SWEEP_KIND=0 python tools/profile_case.py 16384 config3 2 | tail -1; done
