"""Development aid: team-size sweep of the wide kernel (HADI_WIDE_TEAM caps the CTAs per solve)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
os.environ["HADI_FORCE_VARIANT"] = "9"
def run(m1, m2, scheme, nopt, N):
    num = hadi.make_numerics(m1, m2, 0.8, hadi.EUROPEAN, hadi.CALL, scheme, None)
    pts, n = hadi.make_points([100.0 + 0.1 * k for k in range(nopt)], 1.0, N)
    bt = ctx.batch(mdl, num, pts, n)
    ts = []
    for r in range(3):
        bt.launch(); v = bt.fetch(); ts.append(bt.elapsed_ms())
    bt.destroy()
    return min(ts), v
for (m1, m2, N, teams) in ((400, 200, 50, (148, 134, 101, 74, 67, 51, 37)), (100, 50, 50, (101, 51, 34, 26, 17, 13, 7)), (50, 25, 20, (51, 26, 13, 7, 4))):
    for scheme in (1, 0):
        for n in (1, 2):
            out = []
            for G in teams:
                os.environ["HADI_WIDE_TEAM"] = str(G)
                t, v = run(m1, m2, scheme, n, N)
                out.append("G<=%d: %.3f" % (G, t))
            print("%dx%dx%d scheme %d n=%d ms: %s" % (m1 + 1, m2 + 1, N, scheme, n, "  ".join(out)), flush=True)
