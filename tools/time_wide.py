"""Development aid: the wide kernel (variant 9) against the one-CTA / cluster kernels on BASELINE config 4 (401x201x200)
and on a shared-memory grid, for batch sizes between 1 and 148."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
steps = int(os.environ.get("WIDE_STEPS", "200"))
def run(m1, m2, scheme, nopt, N):
    num = hadi.make_numerics(m1, m2, 0.8, hadi.EUROPEAN, hadi.CALL, scheme, None)
    pts, n = hadi.make_points([100.0 + 0.1 * k for k in range(nopt)], 1.0, N)
    bt = ctx.batch(mdl, num, pts, n)
    ts = []
    for r in range(3):
        bt.launch(); v = bt.fetch(); ts.append(bt.elapsed_ms())
    bt.destroy()
    return min(ts), v
for (m1, m2, N) in ((400, 200, steps), (100, 50, 50)):
    for scheme, name in ((hadi.CRAIG_SNEYD, "CS"), (hadi.DOUGLAS, "DO")):
        for nopt in ((1, 4, 8, 18, 37) if os.environ.get("WIDE_QUICK") else (1, 2, 4, 8, 18, 37, 74, 148)):
            os.environ.pop("HADI_FORCE_VARIANT", None)
            os.environ["HADI_WIDE_MAX_ITEMS"] = "0"
            t0, v0 = run(m1, m2, scheme, nopt, N)
            os.environ["HADI_FORCE_VARIANT"] = "9"
            t1, v1 = run(m1, m2, scheme, nopt, N)
            print("%s %dx%dx%d n=%3d: others %.3f ms, wide %.3f ms (%.3f ms/solve, %.1f us/step) equal=%s price %.16g" % (
                name, m1 + 1, m2 + 1, N, nopt, t0, t1, t1 / nopt, 1e3 * t1 / N / max(1, (nopt + 147) // 148), bool(np.array_equal(v0, v1)), v1[0]), flush=True)
