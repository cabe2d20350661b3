echo "persistent 2048"; HSDDP_SOLVE_MODE=1 python tools/profile_case.py 2048 config3 3 | tail -1
for r in 10 15 20 25 30; do for g in 2 4; do echo "hybrid rounds $r groups $g"; HSDDP_SOLVE_MODE=3 HSDDP_HYBRID_ROUNDS=$r HSDDP_HYBRID_GROUPS=$g python tools/profile_case.py 2048 config3 3 | tail -1; done; done
echo "hybrid rounds 20 groups 8"; HSDDP_SOLVE_MODE=3 HSDDP_HYBRID_ROUNDS=20 HSDDP_HYBRID_GROUPS=8 python tools/profile_case.py 2048 config3 3 | tail -1
echo "4096: persistent, hybrid 20/4, 25/8"; HSDDP_SOLVE_MODE=1 python tools/profile_case.py 4096 config3 3 | tail -1
HSDDP_SOLVE_MODE=3 HSDDP_HYBRID_ROUNDS=20 HSDDP_HYBRID_GROUPS=4 python tools/profile_case.py 4096 config3 3 | tail -1
HSDDP_SOLVE_MODE=3 HSDDP_HYBRID_ROUNDS=25 HSDDP_HYBRID_GROUPS=8 python tools/profile_case.py 4096 config3 3 | tail -1
echo "1024: persistent, hybrid 15/2"; HSDDP_SOLVE_MODE=1 python tools/profile_case.py 1024 config3 3 | tail -1
HSDDP_SOLVE_MODE=3 HSDDP_HYBRID_ROUNDS=15 HSDDP_HYBRID_GROUPS=2 python tools/profile_case.py 1024 config3 3 | tail -1
