# why do batched loads in the flat passes not pay?  timing and ncu of the prep / forward kernels of one round, batch 1 / 2 / 4
L=$PWD/hkd-mpc_b200
for v in "" _fb2 _fb4; do echo "lib '$v'"; HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 16384 config3 2 | tail -1; done
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
for v in "" _fb4; do echo "ncu lib '$v'"
HSDDP_LIB=$L/libhsddp_b200$v.so ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_phase" --launch-skip 31 -c 3 python tools/profile_case.py 8192 config3 1 2>&1 | grep -E "k_phase|duration|inst_executed|long_score|issue_active|bytes_read|icc|no_instr|hit_rate"
done
