"""Quick GPU parity + timing probe (development aid; the real tests are tests/test_*.py)."""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("hadi", os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200", "hadi.py"))
hadi = importlib.util.module_from_spec(spec)
spec.loader.exec_module(hadi)
from oracle.reflib import OracleLib

O = OracleLib()
ctx = hadi.Context(0)
base = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
nbad = 0
for (m1, m2, N) in [(50, 25, 20), (100, 50, 20), (20, 10, 7)]:
    for style in (0, 1):
        for dv in (None, divs):
            for put in (0, 1):
                for rf in (0.0, 0.01):
                    b = dict(base)
                    b["r_f"] = rf
                    o = O.solve(93.0, N, 1 / N, m1=m1, m2=m2, theta=0.8, style=style, divs=dv, payoff_put=put, **b)
                    mdl = hadi.make_model(**b)
                    num = hadi.make_numerics(m1, m2, 0.8, style, put, 0, dv)
                    pts, n = hadi.make_points([93.0], 1.0, N, 1 / N)
                    g = ctx.price_batch(mdl, num, pts, n, want_U=True, want_lambda=True)
                    ok = (o["price"] == g["prices"][0] and np.array_equal(o["U"], g["U"][0]) and
                          (style == 0 or np.array_equal(o["lambda"], g["lambda"][0])))
                    if not ok:
                        nbad += 1
                        d = np.abs(o["U"] - g["U"][0])
                        print("DIFF", m1, m2, N, style, dv is not None, put, rf, o["price"], g["prices"][0], d.max(),
                              (d > 0).sum())
print("single-solve parity mismatches:", nbad)

# Jacobian parity
mdl = hadi.make_model(**base)
num = hadi.make_numerics(25, 20, 0.8)
pts, n = hadi.make_points([90.0, 100.0, 107.5], 1.0, 20)
J, b0 = ctx.jacobian_batch(mdl, num, pts, n)
Jo, bo = O.jacobian_batch([90.0, 100.0, 107.5], 20, 1 / 20, m1=25, m2=20, theta=0.8, **base)
print("jacobian equal:", np.array_equal(J, Jo), np.array_equal(b0, bo), np.abs(J - Jo).max())

# config 2 timing: 500 American calls with dividends, 100x50x50
strikes = [70 + 0.12 * i for i in range(500)]
num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, divs)
pts, n = hadi.make_points(strikes, 1.0, 50)
t0 = time.time()
g = ctx.price_batch(mdl, num, pts, n)
t1 = time.time()
g = ctx.price_batch(mdl, num, pts, n)
t2 = time.time()
print("config2 e2e first %.3f ms, second %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
bt = ctx.batch(mdl, num, pts, n)
for r in range(5):
    bt.launch()
    v = bt.fetch()
    print("config2 kernel ms:", bt.elapsed_ms())
po = O.price_batch(strikes[:8], 50, 1 / 50, m1=100, m2=50, theta=0.8, style=1, divs=divs, **base)
print("config2 first 8 equal:", np.array_equal(po, v[:8]), v[:3], po[:3])
flops = 500 * 50 * 5151 * 72
ms = bt.elapsed_ms()
print("algorithmic GFLOP/s: %.1f" % (flops / ms / 1e6))
# European S grid 500 options N=20 (reference's benchmark shape)
num = hadi.make_numerics(50, 25, 0.8)
pts, n = hadi.make_points(strikes, 1.0, 20)
bt2 = ctx.batch(mdl, num, pts, n)
for r in range(3):
    bt2.launch(); bt2.fetch(); print("S-grid EU 500x20 kernel ms:", bt2.elapsed_ms())
# LM calibration 10x10
import math
mats = [1.0 + i * 0.25 if i < 8 else 3.0 + (i - 8) * 0.5 for i in range(10)]
ks = [95.0 + s for s in range(10)]
K = []; T = []; N = []
for Tm in mats:
    for k in ks:
        K.append(k); T.append(Tm); N.append(max(20, int(Tm * 20)))
pts, n = hadi.make_points(K, T, N)
market = [hadi.bs_call(100.0, k, 0.025, 0.2, t) for k, t in zip(K, T)]
for (m1, m2) in [(50, 25), (100, 50)]:
    num = hadi.make_numerics(m1, m2, 0.8)
    t0 = time.time()
    res = ctx.calibrate(mdl, num, pts, n, market, 15, 0.1 * math.sqrt(n), 0.1 * (1 + math.log(n)))
    print("LM", m1, m2, "wall ms %.2f" % ((time.time() - t0) * 1e3), res)
print("launches", ctx.kernel_launches)
