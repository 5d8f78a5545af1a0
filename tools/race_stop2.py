"""Development aid: characterise the corrupted state right after div3."""
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
NITEMS = 900
N = 12
os.environ["HADI_FORCE_VARIANT"] = "0"
num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, DIVS)
pts1, n1 = hadi.make_points([100.0], N / 50.0, N, 1.0 / 50)
pts, nn = hadi.make_points([100.0] * NITEMS, N / 50.0, N, 1.0 / 50)
def single(st, ph):
    os.environ["HADI_DEBUG_STOP"] = "%d:%d" % (st, ph)
    r = ctx.price_batch(mdl, num, pts1, n1, want_U=True, want_lambda=True)
    return r["U"][0].reshape(51, 101), r["lambda"][0].reshape(51, 101)
cands = {}
for st in (7, 8, 9):
    U, Y = single(st, 8); cands["U@%d:P" % st] = U; cands["Y@%d:P" % st] = Y
    _, lam = single(st, 108); cands["lam@%d:P" % st] = lam
for ph, nm in ((1, "div1"), (2, "div2"), (3, "div3")):
    U, Y = single(10, ph); cands["U@10:%s" % nm] = U; cands["Y@10:%s" % nm] = Y
cands["zero"] = np.zeros((51, 101))
refU, refY = cands["U@10:div3"], cands["Y@10:div3"]
os.environ["HADI_DEBUG_STOP"] = "10:3"
for rep in range(4):
    g = ctx.price_batch(mdl, num, pts, nn, want_U=True, want_lambda=True)
    for k in range(nn):
        U = g["U"][k].reshape(51, 101); Y = g["lambda"][k].reshape(51, 101)
        if np.array_equal(U, refU) and np.array_equal(Y, refY):
            continue
        dU = U != refU; dY = Y != refY
        print("rep", rep, "item", k, "U bad", int(dU.sum()), "Y bad", int(dY.sum()))
        for nm, c in cands.items():
            mu = int((U[dU] == c[dU]).sum()) if dU.any() else 0
            my = int((Y[dY] == c[dY]).sum()) if dY.any() else 0
            fu = int((U == c).sum()); fy = int((Y == c).sum())
            print("    %-12s matches bad-U nodes %5d  bad-Y nodes %5d | whole-grid U %5d Y %5d" % (nm, mu, my, fu, fy))
        # rows / cols pattern of bad nodes
        print("    bad-U per row:", dU.sum(axis=1).tolist())
        print("    bad-Y per row:", dY.sum(axis=1).tolist())
        js, is_ = np.nonzero(dY)
        print("    some bad Y:", [(int(j), int(i), float(Y[j, i]), float(refY[j, i])) for j, i in list(zip(js, is_))[:6]])
        js, is_ = np.nonzero(dU)
        print("    some bad U:", [(int(j), int(i), float(U[j, i]), float(refU[j, i])) for j, i in list(zip(js, is_))[:6]])
