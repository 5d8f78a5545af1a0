"""Development aid: hunt the intermittent batch corruption.  Runs many identical items through the
batched kernel under several variants / grid caps and counts items whose full grid differs from
the first item's."""
import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
NITEMS = int(os.environ.get("RM_ITEMS", "900"))
N = int(os.environ.get("RM_N", "20"))
def run(variant, cap, style, divs, reps=3):
    if variant is None: os.environ.pop("HADI_FORCE_VARIANT", None)
    else: os.environ["HADI_FORCE_VARIANT"] = str(variant)
    if cap is None: os.environ.pop("HADI_MAX_CTAS", None)
    else: os.environ["HADI_MAX_CTAS"] = str(cap)
    num = hadi.make_numerics(100, 50, 0.8, style, 0, 0, divs)
    pts1, n1 = hadi.make_points([100.0], N / 50.0, N, 1.0 / 50)
    ref = ctx.price_batch(mdl, num, pts1, n1, want_U=True)["U"][0]
    pts, nn = hadi.make_points([100.0] * NITEMS, N / 50.0, N, 1.0 / 50)
    bad = []
    for rep in range(reps):
        g = ctx.price_batch(mdl, num, pts, nn, want_U=True)
        bad.append(int(sum(not np.array_equal(g["U"][k], ref) for k in range(nn))))
    print("variant", variant, "cap", cap, "style", style, "divs", divs is not None, "N", N, "bad per rep", bad, flush=True)
for variant in (0, 4, 2):
    for cap in (None, 148, 74):
        for style, divs in ((0, None), (1, None), (0, DIVS), (1, DIVS)):
            run(variant, cap, style, divs)
