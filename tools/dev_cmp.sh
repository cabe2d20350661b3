set -x
P=hkd-mpc_b200
for v in "" _ring; do
  export HSDDP_LIB=$PWD/$P/libhsddp_b200$v.so
  python tools/profile_case.py 4096 config2 3
  python tools/profile_case.py 4096 config3 2
  python tools/profile_case.py 1 config2 3
done
for v in _prof _ringprof; do
  export HSDDP_LIB=$PWD/$P/libhsddp_b200$v.so
  python tools/profile_case.py 4096 config3 1
done
export HSDDP_LIB=$PWD/$P/libhsddp_b200.so
for b in 3 4 5; do HSDDP_BLOCKS_PER_SM=$b python tools/profile_case.py 4096 config3 2; done
