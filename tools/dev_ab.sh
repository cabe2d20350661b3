# A/B of library variants on the headline workload: bash tools/dev_ab.sh <suffix|-> ...
P=$PWD/hkd-mpc_b200
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "-" ]; then export HSDDP_LIB=$P/libhsddp_b200.so; else export HSDDP_LIB=$P/libhsddp_b200_$v.so; fi
  echo "== variant $v"
  python tools/profile_case.py 16384 config3 2
done
done
