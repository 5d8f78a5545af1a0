"""Development aid: find the first phase at which items of an identical batch diverge
(libhadi_debug.so, HADI_DEBUG_STOP=step:phase)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
hadi.LIB_PATH = os.path.join(os.path.dirname(hadi.LIB_PATH), "libhadi_debug.so")
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
NITEMS = int(os.environ.get("RM_ITEMS", "900"))
N = 12
os.environ["HADI_FORCE_VARIANT"] = os.environ.get("RM_VARIANT", "0")
num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, DIVS)
pts1, n1 = hadi.make_points([100.0], N / 50.0, N, 1.0 / 50)
pts, nn = hadi.make_points([100.0] * NITEMS, N / 50.0, N, 1.0 / 50)
names = {1: "div1", 2: "div2", 3: "div3", 4: "E", 5: "S1", 6: "R", 7: "S2", 8: "P"}
stops = [(9, 8), (9, 108), (10, 1), (10, 2), (10, 3), (10, 4), (10, 5), (10, 6), (10, 7), (10, 8), (10, 108), (11, 4), (12, 8)]
for (st, ph) in stops:
    os.environ["HADI_DEBUG_STOP"] = "%d:%d" % (st, ph)
    ref = ctx.price_batch(mdl, num, pts1, n1, want_U=True, want_lambda=True)
    for rep in range(3):
        g = ctx.price_batch(mdl, num, pts, nn, want_U=True, want_lambda=True)
        badU = [k for k in range(nn) if not np.array_equal(g["U"][k], ref["U"][0])]
        badY = [k for k in range(nn) if not np.array_equal(g["lambda"][k], ref["lambda"][0], equal_nan=True)]
        msg = ""
        for k in (badU[:2] + [b for b in badY[:2] if b not in badU[:2]]):
            dU = (g["U"][k] != ref["U"][0]).reshape(51, 101)
            dY = ~((g["lambda"][k] == ref["lambda"][0]) | (np.isnan(g["lambda"][k]) & np.isnan(ref["lambda"][0])))
            dY = dY.reshape(51, 101)
            def box(d):
                if not d.any(): return "-"
                r = np.nonzero(d.any(axis=1))[0]; c = np.nonzero(d.any(axis=0))[0]
                return "n=%d rows %d..%d cols %d..%d" % (int(d.sum()), r.min(), r.max(), c.min(), c.max())
            msg += "\n     item %d U[%s] %s[%s]" % (k, box(dU), "lam" if ph >= 100 else "Y", box(dY))
        print("stop step %d after %s%s rep %d: bad U %d, bad %s %d%s" % (st, names[ph % 100], " (lam dump)" if ph >= 100 else "", rep, len(badU), "lam" if ph >= 100 else "Y", len(badY), msg), flush=True)
