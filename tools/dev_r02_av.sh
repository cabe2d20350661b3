# selective batching of the flat passes: SMALL (tiny bodies) x MID (cost / control passes); big LQ / barrier passes stay at 1
L=$PWD/hkd-mpc_b200
for v in "" _s2m1 _s4m1 _s2m2 _s4m2 ""; do echo "lib '$v'"; HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 16384 config3 2 | tail -1; done
for v in "" _s4m1 _s4m2; do echo "lib '$v' 2048 / 1"; HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 2048 config3 2 | tail -1; HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 1 config2 3 | tail -1; done
