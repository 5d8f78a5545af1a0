"""Small driver used under ncu: n Craig-Sneyd solves on the 401x201 grid, N steps.  usage: prof_large.py n N"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("hadi", os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200", "hadi.py"))
hadi = importlib.util.module_from_spec(spec)
spec.loader.exec_module(hadi)
n, N = int(sys.argv[1]), int(sys.argv[2])
ctx = hadi.Context(0)
mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
num = hadi.make_numerics(400, 200, 0.8, hadi.EUROPEAN, hadi.CALL, hadi.CRAIG_SNEYD, None)
pts, n = hadi.make_points([100.0 + 0.1 * k for k in range(n)], 1.0 * N / 200, N)
bt = ctx.batch(mdl, num, pts, n)
for r in range(3):
    bt.launch()
    v = bt.fetch()
    print("ms", bt.elapsed_ms(), "v0", v[0])
