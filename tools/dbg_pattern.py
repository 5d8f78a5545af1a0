"""Development aid: where does the GPU grid differ from the oracle after N steps? usage: dbg_pattern.py m1 m2 N style"""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("hadi", os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200", "hadi.py"))
hadi = importlib.util.module_from_spec(spec); spec.loader.exec_module(hadi)
from oracle.reflib import OracleLib
m1, m2, N, style = (int(x) for x in sys.argv[1:5])
if len(sys.argv) > 5: hadi.LIB_PATH = os.path.join(os.path.dirname(hadi.LIB_PATH), sys.argv[5])
O = OracleLib(); ctx = hadi.Context(0)
b = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
o = O.solve(93.0, N, 1 / 20, m1=m1, m2=m2, theta=0.8, style=style, divs=None, payoff_put=0, **b)
mdl = hadi.make_model(**b); num = hadi.make_numerics(m1, m2, 0.8, style, 0, 0, None)
pts, n = hadi.make_points([93.0], N / 20, N, 1 / 20)
g = ctx.price_batch(mdl, num, pts, n, want_U=True, want_lambda=True)
U = g["U"][0].reshape(m2 + 1, m1 + 1); Uo = o["U"].reshape(m2 + 1, m1 + 1)
d = (U != Uo)
print("mismatch count", d.sum(), "of", d.size)
print("rows with mismatches:", np.where(d.any(axis=1))[0].tolist())
print("cols with mismatches:", np.where(d.any(axis=0))[0].tolist())
j = int(np.where(d.any(axis=1))[0][0]) if d.any() else 0
print("row", j, "gpu", U[j, :12], "\n      ref", Uo[j, :12])
print("row", j, "first bad col", np.where(d[j])[0][:10])
