# A/B of the phased driver's group count on the headline workload: bash tools/dev_groups.sh
P=$PWD/hkd-mpc_b200
for rep in 1 2; do
  echo "== prev G=4"; HSDDP_LIB=$P/libhsddp_b200_prev.so python tools/profile_case.py 16384 config3 2
  for G in 2 4 8; do
    echo "== new G=$G"; HSDDP_PHASED_GROUPS=$G python tools/profile_case.py 16384 config3 2
  done
done
