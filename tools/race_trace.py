"""Development aid: per-phase trace at the price node; report where bad items first leave the reference trace."""
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
use_div = len(sys.argv) > 2 and sys.argv[2] == "div"
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
os.environ["HADI_FORCE_VARIANT"] = "0"
N = 20
num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, DIVS if use_div else None)
pts1, n1 = hadi.make_points([100.0], N / 50.0, N, 1.0 / 50)
r = ctx.price_batch(mdl, num, pts1, n1, want_U=True)
ref = r["U"][0][: 24 * (N + 1)].copy().view(np.uint64); refp = r["prices"][0]
pts, nn = hadi.make_points([100.0] * 1200, N / 50.0, N, 1.0 / 50)
names = {1: "div1", 2: "div2", 3: "div3", 4: "E", 5: "S1", 6: "R", 7: "S2", 8: "P"}
shown = 0
for rep in range(6):
    g = ctx.price_batch(mdl, num, pts, nn, want_U=True)
    bad = np.nonzero(g["prices"] != refp)[0]
    print(libname, "rep", rep, "bad prices", len(bad), flush=True)
    for k in bad[:6]:
        tr = g["U"][k][: 24 * (N + 1)].copy().view(np.uint64)
        d = np.nonzero(tr != ref)[0]
        if len(d) == 0:
            print("   item", k, "trace identical but price differs", g["prices"][k], refp); continue
        f = d[0]; n = f // 24; ph = (f % 24) // 3; what = ("U", "Y", "lam")[f % 3]
        print("   item %d first diff at step %d after %s in %s ; price %.17g vs %.17g ; next: %s" % (k, n, names.get(ph, ph), what, g["prices"][k], refp, [(int(ff // 24), names.get((ff % 24) // 3), ("U", "Y", "lam")[ff % 3]) for ff in d[1:5]]))
