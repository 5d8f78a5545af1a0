"""Development aid: randomised differential test of hadi_jacobian_batch and hadi_calibrate against the C restatement:
random grids, styles, dividend sets, parameters, multi-maturity points; Jacobians, base prices and LM results bit for bit.
usage: fuzz_jacobian.py [seconds] [seed]"""
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
cases = bad = lms = 0
while time.time() - t0 < budget:
    m1 = int(rng.integers(8, 110)); m2 = int(rng.integers(4, min(m1, 55) + 1))
    style = int(rng.integers(0, 2)); put = int(rng.integers(0, 2))
    nd = int(rng.integers(0, 3))
    divs = None
    if nd:
        divs = (list(np.sort(rng.uniform(0.05, 0.95, nd))), list(rng.uniform(0.0, 1.0, nd)), list(rng.uniform(0.0, 0.03, nd)))
    base = dict(S0=100.0, V0=float(rng.uniform(0.02, 0.1)), r_d=0.025, r_f=float(rng.choice([0.0, 0.01])),
                rho=float(rng.uniform(-0.9, 0.3)), sigma=float(rng.uniform(0.1, 0.6)), kappa=float(rng.uniform(0.5, 3.0)),
                eta=float(rng.uniform(0.02, 0.1)))
    nopt = int(rng.integers(1, 9))
    if rng.uniform() < 0.04:   # a Jacobian batch that fills the small-CTA instantiation of the 51 x 26 grid (>= 888 items)
        m1, m2, nopt = 50, 25, int(rng.integers(150, 260))
    elif rng.uniform() < 0.04:   # and one beyond the persistent grid at 101 x 51 (split schedule)
        m1, m2, nopt = 100, 50, int(rng.integers(60, 140))
    Ks = [float(k) for k in rng.uniform(85.0, 115.0, nopt)]
    Ts = [float(t) for t in rng.choice([0.5, 1.0, 1.5], nopt)]
    Ns = [int(x) for x in rng.integers(2, 9, nopt)]
    eps = float(rng.choice([1e-6, 1e-5]))
    mdl = hadi.make_model(**base)
    num = hadi.make_numerics(m1, m2, 0.8, style, put, hadi.DOUGLAS, divs)
    pts, n = hadi.make_points(Ks, Ts, Ns)
    J, b = ctx.jacobian_batch(mdl, num, pts, n, eps)
    Jo, bo = O.jacobian_batch(Ks, Ns, [t / k for t, k in zip(Ts, Ns)], eps=eps, m1=m1, m2=m2, theta=0.8, style=style, divs=divs,
                              payoff_put=put, **base)
    cases += 1
    if not (np.array_equal(J, Jo, equal_nan=True) and np.array_equal(b, bo, equal_nan=True)):
        bad += 1
        print("JACOBIAN MISMATCH m1=%d m2=%d style=%d put=%d nd=%d n=%d eps=%g V0=%r" % (m1, m2, style, put, nd, nopt, eps, base["V0"]), flush=True)
        continue
    if rng.uniform() < 0.15 and np.all(np.isfinite(b)) and nopt >= 5:
        # a short LM run towards perturbed prices: reference schedule against the restatement, speculative against reference
        market = [float(x) * (1.0 + 0.02 * float(rng.uniform(-1, 1))) for x in b]
        a = ctx.calibrate(mdl, num, pts, n, market, 4, 1e-9, 1e-9, eps=eps)
        s = ctx.calibrate(mdl, num, pts, n, market, 4, 1e-9, 1e-9, eps=eps, schedule=hadi.LM_SCHEDULE_SPECULATIVE)
        o = O.calibrate(Ks, Ns, [t / k for t, k in zip(Ts, Ns)], market, max_iter=4, tol=1e-9, delta_tol=1e-9, eps=eps, m1=m1, m2=m2,
                        theta=0.8, style=style, divs=divs, payoff_put=put, **base)
        lms += 1
        okl = list(a["params"]) == list(o["params"]) and a["final_error"] == o["final_error"] and a["iterations"] == o["iterations"]
        oks = list(s["params"]) == list(a["params"]) and s["final_error"] == a["final_error"] and s["lam"] == a["lam"]
        if not (okl and oks):
            bad += 1
            print("LM MISMATCH m1=%d m2=%d style=%d nd=%d n=%d vs-oracle %s speculative %s" % (m1, m2, style, nd, nopt, okl, oks), a["params"], o["params"], flush=True)
print("jacobian cases %d, LM runs %d, mismatches %d" % (cases, lms, bad))
