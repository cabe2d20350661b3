# A/B: '#pragma unroll 4' on the flat passes of the forward / cost code; per-phase cycle accounting of one solve and of 2,048
L=$PWD/hkd-mpc_b200
for v in "" _unroll "" _unroll; do echo "lib '$v'"; HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 16384 config3 2 | tail -1; done
for v in "" _unroll; do echo "lib '$v' 2048, 1"; HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 2048 config3 2 | tail -1; HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 1 config2 3 | tail -1; done
echo "prof n=1"; HSDDP_LIB=$L/libhsddp_b200_prof.so python tools/profile_case.py 1 config2 1
echo "prof n=2048"; HSDDP_LIB=$L/libhsddp_b200_prof.so python tools/profile_case.py 2048 config3 1
echo "prof n=16384"; HSDDP_LIB=$L/libhsddp_b200_prof.so python tools/profile_case.py 16384 config3 1
