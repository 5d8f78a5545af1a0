"""Development aid: timing of BASELINE config 4 (401x201 grid, N=200) on the global-state kernel."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
for scheme, name in ((hadi.CRAIG_SNEYD, "CS"), (hadi.DOUGLAS, "DO")):
    for nopt in (1, 8, 148):
        num = hadi.make_numerics(400, 200, 0.8, hadi.EUROPEAN, hadi.CALL, scheme, None)
        pts, n = hadi.make_points([100.0 + 0.1 * k for k in range(nopt)], 1.0, 200)
        bt = ctx.batch(mdl, num, pts, n)
        ts = []
        for r in range(3):
            bt.launch(); v = bt.fetch(); ts.append(bt.elapsed_ms())
        P = 401 * 201
        flops = nopt * 200 * P * (111 if scheme else 64)
        byts = nopt * 200 * P * 8 * (9 if scheme else 5)
        ms = min(ts)
        print("%s 401x201x200 n=%d: %.2f ms (%.2f ms/solve), %.3f TFLOP/s alg, %.1f GB/s alg-bytes, price %.16g" % (name, nopt, ms, ms / nopt, flops / ms / 1e9, byts / ms / 1e6, v[0]))
        bt.destroy()
