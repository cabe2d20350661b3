# A/B throughput of library variants: bash tools/dev_cmp.sh <suffix> <suffix> ...   ("-" = the product build)
P=$PWD/hkd-mpc_b200
for v in "$@"; do
  if [ "$v" = "-" ]; then export HSDDP_LIB=$P/libhsddp_b200.so; else export HSDDP_LIB=$P/libhsddp_b200_$v.so; fi
  echo "== variant $v"
  python tools/profile_case.py 4096 config2 2
  python tools/profile_case.py 4096 config3 2
  python tools/profile_case.py 1 config2 2
done
