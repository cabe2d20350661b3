# phased driver on mid-size batches: number of index ranges (groups)
for n in 2048 3072 4096 6144; do for g in 1 2 4; do echo "n $n phased groups $g"; HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=$g HSDDP_PHASED_MIN_GROUP=256 python tools/profile_case.py $n config3 3 | tail -1; done; done
