# ncu --set full of the generic backward sweep <36,12,12> and <24,24,0> (one launch each)
cat > /tmp/gen_one.py <<'PY'
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("hkd-mpc_b200"); wl = importlib.import_module("hkd-mpc_b200.workloads")
for xs, us, ys in ((36, 12, 12), (24, 24, 0)):
    n, N = 1776, 30
    one = wl.random_phase(xs, us, ys, N, 1, n=8)
    B = pkg.SinglePhaseBatch(xs, us, ys, N, n)
    for nm in B.INPUTS:
        B.set(nm, np.ascontiguousarray(np.broadcast_to(one[nm][None], (n // 8,) + one[nm].shape).reshape((n,) + one[nm].shape[1:])))
    B.backward_sweep(1e-3); print(xs, us, ys, B.last_ms())
PY
python /tmp/gen_one.py
ncu --set full --clock-control none --import-source on -k regex:"k_generic_backward_sweep" -f -o gpurun_out/r02aq_generic python /tmp/gen_one.py > gpurun_out/ncu_r02aq.log 2>&1
tail -2 gpurun_out/ncu_r02aq.log
