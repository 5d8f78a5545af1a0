"""Development aid: the wide kernel against the shared-memory kernel on small batches of config-2 items (101x51x50,
American + dividends) — the regime of a rank that holds a slice of a sharded chain."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
def run(nopt, style, dv):
    num = hadi.make_numerics(100, 50, 0.8, style, hadi.CALL, hadi.DOUGLAS, dv)
    pts, n = hadi.make_points([70.0 + 0.12 * k for k in range(nopt)], 1.0, 50)
    bt = ctx.batch(mdl, num, pts, n)
    ts = []
    for r in range(4):
        bt.launch(); v = bt.fetch().copy(); ts.append(bt.elapsed_ms())
    info = bt.kernel_info
    bt.destroy()
    return min(ts), v, info
for style, dv, name in ((1, divs, "American+dividends"), (0, None, "European")):
    for nopt in (1, 2, 4, 8, 16, 32, 63, 74, 100, 148):
        os.environ.pop("HADI_FORCE_VARIANT", None)
        t0, v0, i0 = run(nopt, style, dv)
        os.environ["HADI_FORCE_VARIANT"] = "9"
        t1, v1, i1 = run(nopt, style, dv)
        print("%s n=%3d: default %.3f ms %s, wide %.3f ms %s equal=%s" % (name, nopt, t0, i0, t1, i1, bool(np.array_equal(v0, v1))), flush=True)
