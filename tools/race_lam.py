"""Development aid: report disagreements between the shared and the global copy of lambda."""
import os, sys, struct, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
libname = sys.argv[1]
hadi.LIB_PATH = os.path.join(os.path.dirname(hadi.LIB_PATH), libname)
L = hadi.lib()
L.hadi_batch_prof_raw.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
os.environ["HADI_FORCE_VARIANT"] = os.environ.get("RM_VARIANT", "0")
N = 20
num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, None)
pts1, n1 = hadi.make_points([100.0], N / 50.0, N, 1.0 / 50)
ref = ctx.price_batch(mdl, num, pts1, n1)["prices"][0]
pts, nn = hadi.make_points([100.0] * 1200, N / 50.0, N, 1.0 / 50)
bt = ctx.batch(mdl, num, pts, nn)
tot = 0
for rep in range(5):
    bt.launch(); g = bt.fetch(); tot += int((g != ref).sum())
buf = (C.c_longlong * (8 * 296))()
n = L.hadi_batch_prof_raw(bt._h, buf, 296)
a = np.frombuffer(buf, dtype=np.int64).reshape(296, 8)[:n]
print(libname, "bad prices in 5 reps:", tot, "CTAs with mismatches:", int((a[:, 0] > 0).sum()), "total mismatches", int(a[:, 0].sum()))
for c in np.nonzero(a[:, 0] > 0)[0][:25]:
    r = a[c]
    lg = struct.unpack("d", struct.pack("q", int(r[4])))[0]; ls = struct.unpack("d", struct.pack("q", int(r[5])))[0]
    print("  cta %3d count %5d first: step %2d i %3d j %2d global %.17g shared %.17g smid %3d item %d" % (c, r[0], r[1], r[2], r[3], lg, ls, r[6], r[7]))
