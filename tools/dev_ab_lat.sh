# A/B of library variants: headline throughput + single-solve latency.  bash tools/dev_ab_lat.sh <suffix|-> ...
P=$PWD/hkd-mpc_b200
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "-" ]; then export HSDDP_LIB=$P/libhsddp_b200.so; else export HSDDP_LIB=$P/libhsddp_b200_$v.so; fi
  echo "== variant $v"
  python tools/profile_case.py 16384 config3 2 | tail -1
  python tools/profile_case.py 1 config1 4 | tail -2
done
done
