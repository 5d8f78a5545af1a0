"""TEST INFRASTRUCTURE — generates tests/golden/*.json from the REAL reference.

The numbers are produced by oracle/_ref/libhadi_ref.so, i.e. by the reference's own unmodified
sources (see oracle/ref_driver.cpp).  Run it in the build container, where /root/reference is
mounted:   make -C oracle ref && python oracle/make_golden.py
The fixtures are committed; the GPU box never needs /root/reference.

Floats are written with repr(), which round-trips IEEE doubles exactly.
"""
import hashlib
import json
import math
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.reflib import RefLib  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04, theta=0.8)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02, 0.02, 0.02, 0.02])


def digest(a):
    """sha256 of the raw little-endian doubles, with -0.0 normalised to +0.0 (signed zeros are not
    part of the parity contract, SURVEY hard part 8)."""
    a = np.ascontiguousarray(a, dtype=np.float64) + 0.0
    return hashlib.sha256(a.tobytes()).hexdigest()


def market_fixtures(R):
    """Synthetic market, dividend-adjusted market and implied-vol inversion (src/bs.hpp:58-192)."""
    strikes = [90.0 + 2.5 * i for i in range(9)]
    cases = []
    for T in (0.5, 1.0, 2.75):
        plain = R.market(100.0, T, 0.025, strikes)
        div = R.market(100.0, T, 0.025, strikes, DIVS)
        cases.append(dict(S0=100.0, T=T, r_d=0.025, strikes=strikes, plain=[repr(float(x)) for x in plain],
                          dividends=[repr(float(x)) for x in div]))
    iv = []
    for (S, K, r, T, target) in [(100.0, 100.0, 0.025, 1.0, 9.0), (100.0, 90.0, 0.025, 0.5, 12.5),
                                 (100.0, 110.0, 0.025, 2.75, 11.0), (97.1, 100.0, 0.025, 1.0, 6.4),
                                 (100.0, 100.0, 0.025, 1.0, float(R.bs_call(100.0, 100.0, 0.025, 0.2, 1.0)))]:
        for eps in (0.01, 1e-8):
            iv.append(dict(S=S, K=K, r=r, T=T, target=repr(target), eps=eps,
                           newton=repr(float(R.reverse_bs(S, K, r, T, 0.5, target, eps))),
                           bisect=repr(float(R.reverse_bs_dic(S, K, r, T, target, eps, 0.001, 1.0))),
                           vega=repr(float(R.bs_vega(S, K, r, 0.2, T)))))
    # a deep in-the-money target at a tiny maturity: vega < 1e-10 at the first Newton step -> bisection fallback
    S, K, r, T, target, eps = 100.0, 20.0, 0.025, 0.001, 80.2, 1e-6
    iv.append(dict(S=S, K=K, r=r, T=T, target=repr(target), eps=eps, fallback=True,
                   newton=repr(float(R.reverse_bs(S, K, r, T, 0.5, target, eps))),
                   bisect=repr(float(R.reverse_bs_dic(S, K, r, T, target, eps, 0.001, 1.0))),
                   vega=repr(float(R.bs_vega(S, K, r, 0.5, T)))))
    json.dump(dict(divs=DIVS, market=cases, implied_vol=iv), open(os.path.join(OUT, "market.json"), "w"), indent=0)


def lm_trajectory(R, K, T, N, market, *, m1, m2, style, divs, multi, max_iter, tol, dtol, init):
    """The reference's LM loop (src/heston_calibration.cpp:204-417 single maturity, :2692-2831 multi-maturity,
    :3568-3716 American + dividends) driven through the reference's own compute_jacobian* /
    compute_base_prices* / compute_parameter_update_on_device.  Returns the trajectory and the final state the
    drivers print (params, final_error, iteration_count, lambda)."""
    K, T = np.asarray(K, dtype=np.float64), np.asarray(T, dtype=np.float64)
    N = np.asarray(N, dtype=np.int32)
    dt = T / N
    cur = dict(init)
    lam = 0.01
    traj = []
    final_error, iters, converged, solves = 100.0, 0, 0, 0
    kw = dict(m1=m1, m2=m2, style=style, divs=divs, multi=multi)
    if multi:
        kw["maturities"] = T
    for it in range(max_iter):
        r = R.solve_batch(K, N, dt, jac=1, **kw, **cur)
        solves += 6 * K.size
        res = market - r["prices"]
        delta = R.lm_update(r["J"], res, lam)
        nw = dict(cur)
        nw["kappa"] = max(1e-3, cur["kappa"] + delta[0])
        nw["eta"] = max(1e-2, cur["eta"] + delta[1])
        nw["sigma"] = max(1e-2, cur["sigma"] + delta[2])
        nw["rho"] = min(1.0, max(-1.0, cur["rho"] + delta[3]))
        nw["V0"] = max(1e-2, cur["V0"] + delta[4])
        dn = 0.0
        for d in delta:
            dn += d * d
        dn = math.sqrt(dn)
        err = 0.0
        for x in res:
            err += x * x
        step = dict(iter=it, lam=lam, err=repr(float(err)), delta_norm=repr(float(dn)),
                    delta=[repr(float(x)) for x in delta],
                    clamped=[k for k in ("kappa", "eta", "sigma", "rho", "V0")
                             if nw[k] != cur[k] + delta[("kappa", "eta", "sigma", "rho", "V0").index(k)]])
        if dn < dtol or err < tol:
            cur = nw
            final_error, iters, converged = err, it + 1, 1
            traj.append(step)
            break
        r2 = R.solve_batch(K, N, dt, **kw, **nw)
        solves += K.size
        nerr = 0.0
        for x in market - r2["prices"]:
            nerr += x * x
        step["new_err"] = repr(float(nerr))
        step["accepted"] = bool(nerr < err)
        if nerr < err:
            cur = nw
            lam = max(lam / 10.0, 1e-7)
        else:
            lam = min(lam * 10.0, 1e7)
        final_error, iters = min(nerr, err), it + 1
        traj.append(step)
    return dict(n=int(K.size), tol=tol, delta_tol=dtol, max_iter=max_iter, iterations=iters, converged=converged,
                pde_solves=solves, final_error=repr(float(final_error)), lam=repr(float(lam)), trajectory=traj,
                params=[repr(float(cur[k])) for k in ("kappa", "eta", "sigma", "rho", "V0")])


def more_fixtures():
    """Round-2 fixtures (python oracle/make_golden.py more): Jacobians of all four entry-point families on the
    101x51 grid, the LM calibration of BASELINE config 3 (10 strikes x 10 maturities) on both grids, and the
    trajectories of the reference's two other shipped LM drivers (test_calibration_european,
    src/heston_calibration.cpp:26; test_calibration_american_divident_multi_maturity, :3245).  Computed with the
    OpenMP flavour of oracle/_ref (each thread owns its option instance: bit-identical to the serial build)."""
    R = RefLib(omp=True)
    R.set_threads(os.cpu_count() or 1)
    init = {k: BASE[k] for k in ("S0", "V0", "r_d", "r_f", "rho", "sigma", "kappa", "eta", "theta")}
    # ---- Jacobians at 101x51 (N = 20), three strikes per family
    jac = []
    for (style, dv) in [(0, None), (1, None), (0, DIVS), (1, DIVS)]:
        strikes = [92.0, 100.0, 111.0]
        r = R.solve_batch(strikes, 20, 1.0 / 20, m1=100, m2=50, style=style, divs=dv, jac=1, eps=1e-6, **BASE)
        jac.append(dict(m1=100, m2=50, N=20, style=style, div=dv is not None, strikes=strikes, eps=1e-6,
                        base=[repr(float(x)) for x in r["prices"]],
                        J=[[repr(float(x)) for x in row] for row in r["J"]]))
    json.dump(dict(base=BASE, divs=DIVS, cases=jac), open(os.path.join(OUT, "jacobians_101x51.json"), "w"), indent=0)
    out = {}
    # ---- BASELINE config 3: strikes 95 + s, maturities of src/heston_calibration.cpp:2485-2489
    mats = [1.0 + i * 0.25 if i < 8 else 3.0 + (i - 8) * 0.5 for i in range(10)]
    K, T, N = [], [], []
    for Tm in mats:
        for s in range(10):
            K.append(95.0 + 1.0 * s)
            T.append(Tm)
            N.append(max(20, int(Tm * 20)))
    market = np.array([R.bs_call(100.0, k, 0.025, 0.2, t) for k, t in zip(K, T)])
    n = len(K)
    for name, (m1, m2) in (("config3_51x26", (50, 25)), ("config3_101x51", (100, 50))):
        out[name] = lm_trajectory(R, K, T, N, market, m1=m1, m2=m2, style=0, divs=None, multi=1, max_iter=15,
                                  tol=0.1 * math.sqrt(n), dtol=0.1 * (1.0 + math.log(n)), init=init)
        out[name].update(m1=m1, m2=m2, style=0)
    # ---- test_calibration_european: 60 strikes 70..129, T = 1, N = 20, 51x26, tol = delta_tol = 0.1
    K = [100.0 * 0.7 + i * 1 for i in range(60)]
    market = np.array([R.bs_call(100.0, k, 0.025, 0.2, 1.0) for k in K])
    out["shipped_european"] = lm_trajectory(R, K, [1.0] * 60, [20] * 60, market, m1=50, m2=25, style=0, divs=None,
                                            multi=0, max_iter=15, tol=0.1, dtol=0.1, init=init)
    out["shipped_european"].update(m1=50, m2=25, style=0,
                                   survey="4 iterations, kappa=47.9119 eta=0.0312648 sigma=0.01 rho=-1 v0=0.0946873 err=0.0836893")
    # ---- test_calibration_american_divident_multi_maturity: 3 maturities x 60 strikes, market from the model
    divs = ([0.2, 0.4, 0.6, 0.8], [0.10] * 4, [0.0005] * 4)
    K, T, N = [], [], []
    for Tm in (1.0, 1.5, 2.0):
        for i in range(60):
            K.append(100.0 * 0.7 + i * 1)
            T.append(Tm)
            N.append(max(20, int(Tm * 20)))
    Na = np.asarray(N, dtype=np.int32)
    Ta = np.asarray(T)
    gen = dict(init, kappa=3.0, eta=0.1, sigma=0.05, rho=0.2, V0=0.06)
    market = R.solve_batch(K, Na, Ta / Na, maturities=Ta, m1=50, m2=25, style=1, divs=divs, multi=1, **gen)["prices"].copy()
    out["shipped_american_dividend"] = lm_trajectory(R, K, T, N, market, m1=50, m2=25, style=1, divs=divs, multi=1,
                                                     max_iter=20, tol=0.1, dtol=0.3 * 0.1, init=init)
    out["shipped_american_dividend"].update(
        m1=50, m2=25, style=1, divs=divs, market_params=[3.0, 0.1, 0.05, 0.2, 0.06],
        market_sha256=digest(market),
        survey="18 iterations, kappa=0.52875 eta=0.114716 sigma=0.0901574 rho=0.120892 v0=0.0800021 err=0.19322")
    json.dump(out, open(os.path.join(OUT, "lm_more.json"), "w"), indent=0)
    # ---- Modified Craig-Sneyd as the reference ships it (src/solver.hpp:917-1075): price and digest of the full grid
    Rs = RefLib()
    mcs = []
    for (m1, m2, N) in [(50, 25, 20), (100, 50, 20), (64, 32, 9)]:
        price, U = Rs.host_scheme(2, K=100.0, T=1.0, m1=m1, m2=m2, N=N, want_U=True, **BASE)
        mcs.append(dict(m1=m1, m2=m2, N=N, K=100.0, T=1.0, price=repr(float(price)), U_sha256=digest(U)))
    json.dump(dict(base=BASE, cases=mcs), open(os.path.join(OUT, "mcs.json"), "w"), indent=0)
    for k, v in out.items():
        print(k, v["iterations"], v["converged"], v["params"], v["final_error"], v["pde_solves"])


def main():
    R = RefLib()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "more":
        more_fixtures()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "market":
        market_fixtures(R)
        return
    market_fixtures(R)

    # ---- grids (src/grid.cpp:16-96)
    grids = []
    for (m1, m2, K, S0, V0) in [(50, 25, 100.0, 100.0, 0.04), (100, 50, 93.0, 100.0, 0.04),
                                (100, 50, 70.0, 100.0, 0.04 + 1e-6), (20, 10, 120.0, 100.0, 0.09)]:
        s, ds, v, dv = R.grid(m1, m2, K, S0, V0)
        grids.append(dict(m1=m1, m2=m2, K=K, S0=S0, V0=V0, s=[repr(float(x)) for x in s],
                          v=[repr(float(x)) for x in v]))
    json.dump(grids, open(os.path.join(OUT, "grids.json"), "w"), indent=0)

    # ---- single solves: price, digest of the full U, exercise region (lambda > 0) digest and count
    solves = []
    for (m1, m2, N) in [(50, 25, 20), (100, 50, 20), (100, 50, 50), (20, 10, 7), (64, 32, 9)]:
        for style in (0, 1):
            for dv in (None, DIVS):
                for put in (0, 1):
                    for rf in (0.0, 0.01):
                        for K in (93.0, 100.0):
                            if (m1, m2, N) == (100, 50, 50) and (rf != 0.0 or K != 100.0):
                                continue
                            if (m1, m2, N) == (64, 32, 9) and (put or rf != 0.0):
                                continue
                            b = dict(BASE)
                            b["r_f"] = rf
                            r = R.solve_batch([K], N, 1.0 / N, m1=m1, m2=m2, style=style, divs=dv, payoff_put=put,
                                              want_U=True, want_lambda=True, **b)
                            lam = r["lambda"][0]
                            mask = (lam > 0).astype(np.uint8)
                            solves.append(dict(m1=m1, m2=m2, N=N, T=1.0, style=style, div=dv is not None, put=put,
                                               r_f=rf, K=K, price=repr(float(r["prices"][0])),
                                               U_sha256=digest(r["U"][0]),
                                               lam_sha256=digest(lam) if style else None,
                                               exercise_count=int(mask.sum()) if style else 0,
                                               exercise_sha256=hashlib.sha256(mask.tobytes()).hexdigest() if style else None))
    json.dump(dict(base=BASE, divs=DIVS, cases=solves), open(os.path.join(OUT, "solves.json"), "w"), indent=0)

    # ---- the survey's pinned values re-derived from the reference build
    named = {}
    named["EU_call_S_N20"] = R.solve_batch([100.0], 20, 1 / 20, m1=50, m2=25, **BASE)["prices"][0]
    named["EU_call_M_N20"] = R.solve_batch([100.0], 20, 1 / 20, m1=100, m2=50, **BASE)["prices"][0]
    named["AMDIV_call_M_N50"] = R.solve_batch([100.0], 50, 1 / 50, m1=100, m2=50, style=1, divs=DIVS, **BASE)["prices"][0]
    named["AMDIV_put_M_N50"] = R.solve_batch([100.0], 50, 1 / 50, m1=100, m2=50, style=1, divs=DIVS, payoff_put=1, **BASE)["prices"][0]
    named["CS_shuffled_S_N20"] = R.host_scheme(1, K=100.0, T=1.0, m1=50, m2=25, N=20, **BASE)
    named["CS_shuffled_M_N20"] = R.host_scheme(1, K=100.0, T=1.0, m1=100, m2=50, N=20, **BASE)
    named["DO_host_S_N20"] = R.host_scheme(0, K=100.0, T=1.0, m1=50, m2=25, N=20, **BASE)
    json.dump({k: repr(float(v)) for k, v in named.items()}, open(os.path.join(OUT, "named.json"), "w"), indent=0)

    # ---- Jacobians (src/jacobian_computation.cpp:204-364 and variants)
    jac = []
    for (m1, m2, N, style, dv, strikes) in [(25, 20, 20, 0, None, [90.0, 100.0, 107.5]),
                                           (50, 25, 20, 0, None, [95.0, 100.0]),
                                           (50, 25, 20, 1, DIVS, [95.0, 105.0]),
                                           (50, 25, 20, 1, None, [100.0]),
                                           (50, 25, 20, 0, DIVS, [100.0])]:
        r = R.solve_batch(strikes, N, 1.0 / N, m1=m1, m2=m2, style=style, divs=dv, jac=1, eps=1e-6, **BASE)
        jac.append(dict(m1=m1, m2=m2, N=N, style=style, div=dv is not None, strikes=strikes, eps=1e-6,
                        base=[repr(float(x)) for x in r["prices"]],
                        J=[[repr(float(x)) for x in row] for row in r["J"]]))
    # multi-maturity (src/heston_calibration.cpp:2174)
    Ns = np.array([20, 25, 30], dtype=np.int32)
    Ts = np.array([1.0, 1.25, 1.5])
    r = R.solve_batch([95.0, 100.0, 105.0], Ns, Ts / Ns, maturities=Ts, m1=30, m2=15, jac=1, multi=1, **BASE)
    jac.append(dict(m1=30, m2=15, N=[int(x) for x in Ns], T=[float(x) for x in Ts], style=0, div=False,
                    strikes=[95.0, 100.0, 105.0], eps=1e-6, multi=1, base=[repr(float(x)) for x in r["prices"]],
                    J=[[repr(float(x)) for x in row] for row in r["J"]]))
    json.dump(dict(base=BASE, divs=DIVS, cases=jac), open(os.path.join(OUT, "jacobians.json"), "w"), indent=0)

    # ---- LM update (src/jacobian_computation.cpp:107-195) on a fixed pseudo-random system
    rng = np.random.default_rng(20261018)
    Jr = rng.normal(size=(37, 5))
    rr = rng.normal(size=37)
    lm = dict(J=[[repr(float(x)) for x in row] for row in Jr], r=[repr(float(x)) for x in rr], cases=[])
    for lam in (0.01, 1e-7, 10.0):
        lm["cases"].append(dict(lam=lam, delta=[repr(float(x)) for x in R.lm_update(Jr, rr, lam)]))
    A = Jr[:5, :].T @ Jr[:5, :] + np.eye(5)
    b = rr[:5]
    lm["solve5"] = dict(A=[[repr(float(x)) for x in row] for row in A], b=[repr(float(x)) for x in b],
                        x=[repr(float(x)) for x in R.solve5(A, b)])
    json.dump(lm, open(os.path.join(OUT, "lm_update.json"), "w"), indent=0)

    # ---- Black-Scholes helper (src/bs.hpp:44)
    bs = [dict(S=100.0, K=K, r=0.025, vol=0.2, T=T, price=repr(float(R.bs_call(100.0, K, 0.025, 0.2, T))))
          for K in (90.0, 95.0, 100.0, 104.5, 110.0) for T in (0.25, 1.0, 2.75)]
    json.dump(bs, open(os.path.join(OUT, "bs.json"), "w"), indent=0)

    # ---- LM trajectory of the reference's shipped multi-maturity driver set-up
    # (src/heston_calibration.cpp:2428-2925: 10 maturities x 20 strikes, 51x26 grid).  The loop below is
    # the same LM loop driven through the reference's compute_jacobian_multi_maturity /
    # compute_base_prices_multi_maturity / compute_parameter_update_on_device.
    mats = [1.0 + i * 0.25 if i < 8 else 3.0 + (i - 8) * 0.5 for i in range(10)]
    strikes = [100.0 * 0.95 + i * 0.5 for i in range(20)]
    K, T, N = [], [], []
    for Tm in mats:
        for s in strikes:
            K.append(s)
            T.append(Tm)
            N.append(max(20, int(Tm * 20)))
    K, T, N = np.array(K), np.array(T), np.array(N, dtype=np.int32)
    dt = T / N
    market = np.array([R.bs_call(100.0, k, 0.025, 0.2, t) for k, t in zip(K, T)])
    n = K.size
    tol, dtol = 0.1 * math.sqrt(n), 0.1 * (1.0 + math.log(n))
    cur = dict(BASE)
    lam = 0.01
    traj = []
    for it in range(15):
        r = R.solve_batch(K, N, dt, maturities=T, m1=50, m2=25, jac=1, multi=1, **cur)
        res = market - r["prices"]
        delta = R.lm_update(r["J"], res, lam)
        nw = dict(cur)
        nw["kappa"] = max(1e-3, cur["kappa"] + delta[0])
        nw["eta"] = max(1e-2, cur["eta"] + delta[1])
        nw["sigma"] = max(1e-2, cur["sigma"] + delta[2])
        nw["rho"] = min(1.0, max(-1.0, cur["rho"] + delta[3]))
        nw["V0"] = max(1e-2, cur["V0"] + delta[4])
        dn = 0.0
        for d in delta:
            dn += d * d
        dn = math.sqrt(dn)
        err = 0.0
        for x in res:
            err += x * x
        traj.append(dict(iter=it, lam=lam, err=repr(float(err)), delta_norm=repr(float(dn)),
                         delta=[repr(float(x)) for x in delta]))
        if dn < dtol or err < tol:
            cur = nw
            break
        r2 = R.solve_batch(K, N, dt, maturities=T, m1=50, m2=25, multi=1, **nw)
        nerr = 0.0
        for x in market - r2["prices"]:
            nerr += x * x
        if nerr < err:
            cur = nw
            lam = max(lam / 10.0, 1e-7)
        else:
            lam = min(lam * 10.0, 1e7)
    json.dump(dict(n=n, tol=tol, delta_tol=dtol, iterations=len(traj), trajectory=traj,
                   params=[repr(float(cur[k])) for k in ("kappa", "eta", "sigma", "rho", "V0")],
                   survey=dict(kappa="4.560782419250094", eta="0.03985674060896215", sigma="0.1047557213238867",
                               rho="-0.112460685320953", v0="0.04208380263222991", err="1.294560550586723")),
              open(os.path.join(OUT, "lm_multi_maturity.json"), "w"), indent=0)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
