# L2 prefetch of the arrays the prep / forward kernels stream through
L=$PWD/hkd-mpc_b200
for v in _nopf "" _nopf ""; do echo "lib '$v'"; HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 16384 config3 2 | tail -1; done
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_phase" --launch-skip 31 -c 3 python tools/profile_case.py 8192 config3 1 2>&1 | grep -E "k_phase|duration|inst_executed|long_score|bytes_read|hit_rate"
