for g in 4 6 8; do HSDDP_PHASED_GROUPS=$g HSDDP_SOLVE_MODE=2 python tools/profile_case.py 16384 config3 2 | tail -1 | sed "s/^/phased groups $g: /"; done
for g in 4 8; do HSDDP_PHASED_GROUPS=$g HSDDP_SOLVE_MODE=2 python tools/profile_case.py 12288 config3 2 | tail -1 | sed "s/^/phased groups $g: /"; done
HSDDP_SOLVE_MODE=1 python tools/profile_case.py 12288 config3 2 | tail -1 | sed "s/^/persistent: /"
for g in 8; do HSDDP_PHASED_GROUPS=$g HSDDP_SOLVE_MODE=2 python tools/profile_case.py 65536 config3 1 | tail -1 | sed "s/^/phased groups $g: /"; done
