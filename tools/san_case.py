"""Small split-schedule workload for compute-sanitizer (memcheck / racecheck): 51x26 and 101x51 American + dividends,
more solves than persistent CTAs so that solves are cut and handed over."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
hadi = ge.load_hadi()
ctx = hadi.Context(0)
mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
for m1, m2, n, N in ((50, 25, 520, 8), (100, 50, 330, 7)):
    num = hadi.make_numerics(m1, m2, 0.8, 1, 0, 0, divs)
    pts, n = hadi.make_points([75 + 50.0 * i / n for i in range(n)], 1.0, [N + (i % 3) for i in range(n)])
    a = ctx.price_batch(mdl, num, pts, n)["prices"].copy()
    os.environ["HADI_NO_SPLIT"] = "1"
    b = ctx.price_batch(mdl, num, pts, n)["prices"].copy()
    del os.environ["HADI_NO_SPLIT"]
    print(m1, m2, n, "split == whole:", bool(np.array_equal(a, b)), flush=True)
