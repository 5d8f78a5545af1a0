import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle.reflib import RefLib
hadi = ge.load_hadi()
L = hadi.lib(); L.hadi_exact_reruns.restype = C.c_longlong; L.hadi_exact_reruns.argtypes = [C.c_void_p]
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
K = [70 + 0.12 * i for i in range(500)]
os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
R = RefLib(omp=True)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
for style in (1, 0):
    for dv in (DIVS, None):
        ref = R.solve_batch(K, 50, 1 / 50, m1=100, m2=50, theta=0.8, style=style, divs=dv, **BASE)["prices"]
        num = hadi.make_numerics(100, 50, 0.8, style, 0, 0, dv)
        pts, nn = hadi.make_points(K, 1.0, 50)
        tot = 0
        bt = ctx.batch(mdl, num, pts, nn)
        for rep in range(4):
            bt.launch(); g = bt.fetch()
            tot += int((g != ref).sum())
        cyc = (C.c_longlong * 8)(); L.hadi_batch_phase_cycles(bt._h, cyc)
        print("variant", os.environ.get("HADI_FORCE_VARIANT"), "style", style, "div", dv is not None, "mismatches in 4 reps:", tot, "exact reruns:", L.hadi_exact_reruns(ctx._h))
