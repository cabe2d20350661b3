# N4: generic SinglePhase<xs,us,ys> sweeps: parity tests, throughput; then the whole GPU suite
timeout 600 python -m pytest tests/test_generic_phase.py -x -q -m gpu -s 2>&1 | tail -8
python tools/generic_phase_bench.py 4096 60
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
