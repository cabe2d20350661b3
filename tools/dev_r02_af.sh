# per-phase cycle accounting of ONE solve in the latency build (k_solve_lat: the concurrent line search switched off)
L=$PWD/hkd-mpc_b200
HSDDP_CLUSTER_LS=0 python tools/profile_case.py 1 config2 2 | tail -1
HSDDP_CLUSTER_LS=0 HSDDP_LIB=$L/libhsddp_b200_prof.so python tools/profile_case.py 1 config2 1
