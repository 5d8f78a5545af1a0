"""Development aid: count corrupted items of an identical American+dividend batch for a given library."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
libname = sys.argv[1]
hadi.LIB_PATH = os.path.join(os.path.dirname(hadi.LIB_PATH), libname)
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
os.environ["HADI_FORCE_VARIANT"] = os.environ.get("RM_VARIANT", "0")
N = int(os.environ.get("RM_N", "20"))
for style, divs in ((1, DIVS), (1, None), (0, DIVS)):
    num = hadi.make_numerics(100, 50, 0.8, style, 0, 0, divs)
    pts1, n1 = hadi.make_points([100.0], N / 50.0, N, 1.0 / 50)
    ref = ctx.price_batch(mdl, num, pts1, n1)["prices"][0]
    pts, nn = hadi.make_points([100.0] * 1200, N / 50.0, N, 1.0 / 50)
    bt = ctx.batch(mdl, num, pts, nn)
    bad = []
    for rep in range(10):
        bt.launch(); g = bt.fetch()
        bad.append(int((g != ref).sum()))
    print(libname, "style", style, "divs", divs is not None, "bad per rep of 1200:", bad, flush=True)
