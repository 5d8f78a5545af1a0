"""Small driver used under ncu: one prepared batch, a few launches.  usage: prof_case.py n N m1 m2 style ndiv reps"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("hadi", os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200", "hadi.py"))
hadi = importlib.util.module_from_spec(spec)
spec.loader.exec_module(hadi)

n, N, m1, m2, style, ndiv, reps = (int(x) for x in sys.argv[1:8])
ctx = hadi.Context(0)
mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4) if ndiv else None
num = hadi.make_numerics(m1, m2, 0.8, style, 0, 0, divs)
strikes = [70 + 60.0 * i / max(n, 1) for i in range(n)]
pts, n = hadi.make_points(strikes, 1.0, N)
bt = ctx.batch(mdl, num, pts, n)
for r in range(reps):
    bt.launch()
    v = bt.fetch()
    print("ms", bt.elapsed_ms(), "v0", v[0])
