# persistent (mode 1) vs phased (mode 2) across batch sizes: bash tools/dev_sizes.sh
for n in 2048 4096 8192 12288; do
  for m in 1 2; do
    echo "== n=$n mode=$m"; HSDDP_SOLVE_MODE=$m python tools/profile_case.py $n config3 2 | tail -1
  done
done
echo "== round trace"; HSDDP_DEBUG=1 python tools/profile_case.py 16384 config3 1 2> gpurun_out/rounds.log | tail -1
