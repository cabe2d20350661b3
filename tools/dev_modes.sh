for m in 1 2 1 2; do HSDDP_SOLVE_MODE=$m python tools/profile_case.py 16384 config3 2 | sed "s/^/mode $m: /"; done
for m in 1 2; do HSDDP_SOLVE_MODE=$m python tools/profile_case.py 65536 config3 1 | sed "s/^/mode $m: /"; done
