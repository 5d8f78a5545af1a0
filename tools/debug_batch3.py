import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle.reflib import OracleLib
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
O = OracleLib()
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
for N in (50, 12, 11, 10):
    o = O.solve(100.0, N, 1.0 / 50, m1=100, m2=50, theta=0.8, style=1, divs=DIVS, **BASE)
    num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, DIVS)
    pts, nn = hadi.make_points([100.0] * 600, N / 50.0, N, 1.0 / 50)
    for rep in range(3):
        g = ctx.price_batch(mdl, num, pts, nn, want_U=True, want_lambda=True)
        nbad = 0
        for k in range(nn):
            if not np.array_equal(g["U"][k], o["U"]):
                nbad += 1
                if nbad <= 2:
                    d = (g["U"][k] != o["U"]).reshape(51, 101)
                    dl = (g["lambda"][k] != o["lambda"]).reshape(51, 101)
                    rows = np.nonzero(d.any(axis=1))[0]; cols = np.nonzero(d.any(axis=0))[0]
                    print("N", N, "item", k, "U diff count", int(d.sum()), "rows", rows.min(), rows.max(), "cols", cols.min(), cols.max(),
                          "max", float(np.max(np.abs(g["U"][k] - o["U"]))), "lambda diffs", int(dl.sum()))
        print("N", N, "rep", rep, "bad items", nbad)
