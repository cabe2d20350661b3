# phased driver knobs on the final code (with k_lr_w1): groups x one-warp-sweep threshold
for g in 3 4 6 8; do for w in 800 1036 1184 1600; do echo "groups $g w1_min $w"; HSDDP_PHASED_GROUPS=$g HSDDP_W1_MIN_BLOCKS=$w python tools/profile_case.py 16384 config3 2 | tail -1; done; done
