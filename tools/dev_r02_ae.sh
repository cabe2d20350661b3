# A/B: linear-rollout chunk size (2, 3 stages per ring half instead of 4); resident blocks the prep / forward kernels are compiled for (5, 7 instead of 6)
L=$PWD/hkd-mpc_b200
for v in "" _lr2 _lr3 _mb5 _mb7 ""; do echo "lib '$v'"
  HSDDP_LIB=$L/libhsddp_b200$v.so python tools/profile_case.py 16384 config3 2 | tail -1
done
