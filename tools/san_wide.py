"""Small wide-kernel workload (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
hadi = ge.load_hadi()
ctx = hadi.Context(0)
mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
os.environ["HADI_FORCE_VARIANT"] = "9"
for m1, m2, n, N, scheme in ((36, 18, 1, 2, 0), (100, 50, 2, 2, 0), (100, 50, 5, 2, 1), (400, 200, 2, 2, 0)):
    num = hadi.make_numerics(m1, m2, 0.8, 0, 0, scheme, None)
    pts, n = hadi.make_points([95 + 2.0 * i for i in range(n)], 1.0, N)
    a = ctx.price_batch(mdl, num, pts, n)["prices"].copy()
    print(m1, m2, n, a, flush=True)
    if os.environ.get("HADI_WIDE_STOP"): break
