import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle.reflib import RefLib
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
R = RefLib(omp=True)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
for name, K in (("same", [100.0] * 1000), ("two", [100.0, 130.0] * 500), ("chain", [70 + 0.06 * i for i in range(1000)])):
    ref = R.solve_batch(K, 50, 1 / 50, m1=100, m2=50, theta=0.8, style=1, divs=DIVS, **BASE)["prices"]
    num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, DIVS)
    pts, nn = hadi.make_points(K, 1.0, 50)
    bt = ctx.batch(mdl, num, pts, nn)
    tot = 0; worst = 0.0
    for rep in range(4):
        bt.launch(); g = bt.fetch()
        bad = g != ref
        tot += int(bad.sum()); worst = max(worst, float(np.max(np.abs(g - ref))))
    print(name, "mismatches in 4 reps of 1000:", tot, "worst abs diff", worst)
