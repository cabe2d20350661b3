# End-of-round validation on one B200: GPU tests, smoke, bench line, then the ncu launch list of the same bench command.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || { tail -5 gpurun_out/bench_n1.err; exit 1; }
cat gpurun_out/bench_n1.json
python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2700 --csv \
  --log-file gpurun_out/r01j_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
