for g in 2 3 4 6; do for w in 500 800 1184; do echo "groups $g w1_min $w"; HSDDP_PHASED_GROUPS=$g HSDDP_W1_MIN_BLOCKS=$w python tools/profile_case.py 16384 config3 2 | tail -1; done; done
echo "kind 1 (always w1), groups 4"; HSDDP_SWEEP_KIND=1 python tools/profile_case.py 16384 config3 2 | tail -1
