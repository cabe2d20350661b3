python bench.py --problems 2048 --no-cpu-baseline --no-latency --no-strong --steps 5 --warmup 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench 2048:', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
python tools/profile_case.py 2048 config3 5 | tail -3
python - <<'PY'
import importlib, sys, time
sys.path.insert(0,'.')
pkg=importlib.import_module('hkd-mpc_b200'); wl=importlib.import_module('hkd-mpc_b200.workloads')
for first in (0, 2048*3, 2048*7):
    w=wl.config3(pkg,2048,first=first)
    B=pkg.MultiPhaseDDPBatch(0); B.set_problems(w.schedules,w.schedule_id); B.set_initial_condition(w.x0)
    ms=[]
    for r in range(5):
        B.reset(); B.event_record(0); B.solve_async(); B.event_record(1); B.sync(); ms.append(round(B.event_elapsed_ms(0,1),2))
    print('shard first', first, 'solve ms', ms, 'iters', B.info()['n_iter'].sum())
PY
