# N4: generic SinglePhase<xs,us,ys> sweeps: parity tests, throughput
timeout 600 python -m pytest tests/test_generic_phase.py -x -q -m gpu -s 2>&1 | tail -8
HSDDP_VERBOSE=1 python tools/generic_phase_bench.py 4096 60 2>&1 | sort -u | cut -c1-420
python tools/generic_phase_bench.py 16384 60 2>&1 | cut -c1-420
