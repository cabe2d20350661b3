python tools/strong_shards.py 2048 1
python - <<'PY'
import importlib, sys, numpy as np
sys.path.insert(0, '.')
pkg = importlib.import_module("hkd-mpc_b200"); wl = importlib.import_module("hkd-mpc_b200.workloads")
for first in (2048, 4096):
    w = wl.config3(pkg, 2048, 0.6, first=first)
    B = pkg.MultiPhaseDDPBatch(0); B.set_problems(w.schedules, w.schedule_id); B.set_initial_condition(w.x0)
    B.reset(); B.solve(); i = B.info()
    print(first, "sweeps/iter pct", np.percentile(i['n_sweeps'] / np.maximum(i['n_iter'], 1), [50, 90, 99, 100]).round(2), "max sweeps", i['n_sweeps'].max(),
          "trials/iter pct", np.percentile(i['n_trials'] / np.maximum(i['n_iter'], 1), [50, 90, 99, 100]).round(2), "status", np.bincount(i['status'], minlength=4))
PY
