# A/B of the batch size of flat_pass (1 = the plain loops again, with the terminal records on the fourth warp)
L=$PWD/hkd-mpc_b200
for v in _fb1 _fb2; do echo "lib '$v'"
  HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 16384 config3 2 | tail -1
  HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 2048 config3 2 | tail -1
  HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 1 config2 3 | tail -1
done
