for n in 4096 8192 16384; do for m in 1 2; do HSDDP_SOLVE_MODE=$m python tools/profile_case.py $n config3 2 | tail -1 | sed "s/^/mode $m: /"; done; done
for m in 1 2; do HSDDP_SOLVE_MODE=$m python tools/profile_case.py 65536 config3 1 | sed "s/^/mode $m: /"; done
python -m pytest tests -m gpu -x -q -k "solve_modes" 2>&1 | tail -2
