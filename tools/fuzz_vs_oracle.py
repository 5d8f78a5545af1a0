"""Development aid: randomised differential test of the library against the C restatement (oracle/hadi_oracle.c): random
grid shapes, styles, payoffs, dividend sets, schemes, boundary sets, step counts, strikes and rates; prices, full grids
and multipliers compared bit for bit.  usage: fuzz_vs_oracle.py [seconds] [seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
from oracle.reflib import OracleLib
O = OracleLib()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = hadi.Context(0)
t0 = time.time()
cases = bad = 0
kinds = {}
while time.time() - t0 < budget:
    m1 = int(rng.choice([rng.integers(6, 40), rng.integers(40, 140), rng.integers(140, 330)], p=[0.3, 0.5, 0.2]))
    m2 = int(rng.integers(4, min(m1, 90) + 1))
    if os.environ.get("FUZZ_LARGE"):   # grids beyond shared memory (one-CTA, 1024-thread and wide kernels)
        m1 = int(rng.integers(330, 760)); m2 = int(rng.integers(40, 160))
    scheme = int(rng.choice([0, 0, 0, 1, 2, 3]))
    style = int(rng.integers(0, 2)) if scheme == 0 else 0
    put = int(rng.integers(0, 2))
    bc = int(rng.integers(0, 2)) if (scheme == 0 and put) else 0
    nd = int(rng.integers(0, 4)) if scheme == 0 else 0
    divs = None
    if nd:
        dates = np.sort(rng.uniform(0.05, 0.95, nd))
        divs = (list(dates), list(rng.uniform(0.0, 1.0, nd)), list(rng.uniform(0.0, 0.03, nd)))
    div_all = int(rng.integers(0, 2)) if nd else 0
    N = int(rng.integers(1, 4 if os.environ.get("FUZZ_LARGE") else 9))
    T = float(rng.choice([0.25, 1.0, 2.0]))
    theta = float(rng.choice([0.5, 0.8, 0.8, 1.0]))
    base = dict(S0=float(rng.choice([100.0, 100.0, 80.0, 123.4])), V0=float(rng.choice([0.04, 0.09, 0.0225])),
                r_d=float(rng.choice([0.025, 0.025, 0.0, 0.06])), r_f=float(rng.choice([0.0, 0.01])),
                rho=float(rng.uniform(-0.9, 0.3)), sigma=float(rng.uniform(0.1, 0.6)), kappa=float(rng.uniform(0.5, 3.0)),
                eta=float(rng.uniform(0.02, 0.1)))
    nopt = int(rng.choice([1, 2, 5]))
    if os.environ.get("FUZZ_LARGE") and rng.uniform() < 0.3:
        nopt = int(rng.choice([9, 17, 33, 45]))   # teams of 16, 8, 4 CTAs (rows not resident) and the one-CTA kernel
    big = scheme == 0 and rng.uniform() < 0.03   # now and then a batch beyond the persistent grid (split schedule, variant 11)
    if big:
        m1, m2 = [(50, 25), (100, 50), (40, 20)][int(rng.integers(0, 3))]
        nopt = int(rng.integers(300, 1100))
    Ks = [float(k) for k in rng.uniform(80.0, 120.0, nopt)]
    mdl = hadi.make_model(**base)
    if os.environ.get("FUZZ_VERBOSE"):
        print("start m1=%d m2=%d scheme=%d style=%d put=%d bc=%d nd=%d all=%d N=%d n=%d" % (m1, m2, scheme, style, put, bc, nd, div_all, N, nopt), flush=True)
    try:
        num = hadi.make_numerics(m1, m2, theta, style, put, scheme, divs, boundary=bc, dividend_schedule=div_all)
        Ns = [N + int(x) for x in rng.integers(0, 4, nopt)] if big else [N] * nopt
        pts, n = hadi.make_points(Ks, T, Ns)
        g = ctx.price_batch(mdl, num, pts, n, want_U=not big, want_lambda=bool(style) and not big)
    except hadi.HadiError as e:
        kinds["refused: " + str(e)[:60]] = kinds.get("refused: " + str(e)[:60], 0) + 1
        continue
    cases += 1
    if os.environ.get("FUZZ_VERBOSE"):
        print("case", cases, "m1=%d m2=%d scheme=%d style=%d nd=%d N=%d n=%d" % (m1, m2, scheme, style, nd, N, nopt), "%.2fs" % (time.time() - t0), flush=True)
    kinds["big batches"] = kinds.get("big batches", 0) + int(big)
    info = "m1=%d m2=%d scheme=%d style=%d put=%d bc=%d nd=%d all=%d N=%d n=%d theta=%g S0=%g r_d=%g V0=%g" % (
        m1, m2, scheme, style, put, bc, nd, div_all, N, nopt, theta, base["S0"], base["r_d"], base["V0"])
    for k, K in enumerate(Ks):
        if (big and k % 37 != 0) or (not big and nopt > 5 and k % 8 != 0):
            continue
        o = O.solve(K, Ns[k], T / Ns[k], m1=m1, m2=m2, theta=theta, style=style, divs=divs, payoff_put=put, scheme=scheme, bc=bc,
                    div_all=div_all, want_lambda=bool(style), **base)
        ok = g["prices"][k] == o["price"] and (big or np.array_equal(g["U"][k], o["U"], equal_nan=True))
        if style and not big:
            ok = ok and np.array_equal(g["lambda"][k], o["lambda"], equal_nan=True)
        if not ok:
            bad += 1
            print("MISMATCH", info, "K=%r" % K, "price", g["prices"][k], o["price"], flush=True)
            break
print("cases %d, mismatches %d, exact re-solves %d, %s" % (cases, bad, ctx.exact_reruns, kinds))
