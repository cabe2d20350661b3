# End-of-round evidence on one B200 (final code): GPU tests, smoke, bench line, the ncu launch list of the bench command,
# ncu --set full of one round's kernels, the configuration sweep.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || { tail -5 gpurun_out/bench_final.err; exit 1; }
tail -c 1500 gpurun_out/bench_final.json
python bench.py --steps 2 --warmup 1 --no-latency --no-cpu-baseline --no-generic > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 7000 --csv \
  --log-file gpurun_out/r02zy_launches.csv python bench.py --steps 2 --warmup 1 --no-latency --no-cpu-baseline --no-generic > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
export HSDDP_SOLVE_MODE=2 HSDDP_PHASED_GROUPS=1
ncu --set full --clock-control none --import-source on -k regex:"k_phase|k_sweep_w1|k_lr_w1" --launch-skip 51 -c 5 -f -o gpurun_out/r02zy_round \
  python tools/profile_case.py 8192 config3 1 > gpurun_out/ncu_r02zy.log 2>&1
tail -2 gpurun_out/ncu_r02zy.log
unset HSDDP_SOLVE_MODE HSDDP_PHASED_GROUPS
python tools/config_sweep.py > gpurun_out/r02zy_config_sweep.jsonl 2> gpurun_out/config_sweep.err
cat gpurun_out/r02zy_config_sweep.jsonl | cut -c1-330
