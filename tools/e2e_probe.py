"""Development aid: where the end-to-end time of hadi_price_batch goes (HADI_HOST_TIMING=1 prints the C++ stages)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
ctx = hadi.Context(0)
mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, divs)
pts, n = hadi.make_points([70 + 0.12 * i for i in range(500)], 1.0, 50)
for r in range(6):
    t0 = time.perf_counter()
    ctx.price_batch(mdl, num, pts, n)
    print("python wall %.1f us" % ((time.perf_counter() - t0) * 1e6), flush=True)
