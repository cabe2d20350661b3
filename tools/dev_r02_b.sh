# round 2, GPU call B: the full GPU test-suite with the at-scale parity tests, then the bench line
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err || tail -20 gpurun_out/bench_r02b.err
cat gpurun_out/bench_r02b.json
