"""The C restatement against the REAL reference (oracle/_ref: the reference's unmodified sources
compiled against the Kokkos stand-in) on seeded inputs.  Skipped where oracle/_ref is not built."""
import numpy as np

from conftest import BASE, DIVS


def test_random_solves(oracle, reflib):
    rng = np.random.default_rng(20261018)
    for trial in range(10):
        m2 = int(rng.integers(6, 20))
        m1 = 2 * m2 + int(rng.integers(0, 5))
        N = int(rng.integers(3, 15))
        K = float(np.round(rng.uniform(70, 130), 2))
        style = int(rng.integers(0, 2))
        put = int(rng.integers(0, 2))
        dv = DIVS if rng.integers(0, 2) else None
        b = dict(BASE)
        b.update(rho=float(rng.uniform(-0.95, 0.5)), sigma=float(rng.uniform(0.1, 0.9)),
                 kappa=float(rng.uniform(0.5, 4.0)), eta=float(rng.uniform(0.02, 0.2)),
                 r_f=float(rng.choice([0.0, 0.01])), V0=float(np.round(rng.uniform(0.02, 0.3), 3)))
        T = float(rng.choice([0.5, 1.0, 2.0]))
        r = reflib.solve_batch([K], N, T / N, m1=m1, m2=m2, theta=0.8, style=style, divs=dv, payoff_put=put,
                               want_U=True, want_lambda=True, **b)
        o = oracle.solve(K, N, T / N, m1=m1, m2=m2, theta=0.8, style=style, divs=dv, payoff_put=put, **b)
        assert r["prices"][0] == o["price"]
        assert np.array_equal(r["U"][0], o["U"])
        if style:
            assert np.array_equal(r["lambda"][0], o["lambda"])


def test_jacobian_and_lm(oracle, reflib):
    r = reflib.solve_batch([90.0, 100.0], 12, 1 / 12, m1=30, m2=15, theta=0.8, style=1, divs=DIVS, jac=1, **BASE)
    J, b0 = oracle.jacobian_batch([90.0, 100.0], 12, 1 / 12, m1=30, m2=15, theta=0.8, style=1, divs=DIVS, **BASE)
    assert np.array_equal(r["J"], J) and np.array_equal(r["prices"], b0)
    res = np.array([0.3, -0.2])
    assert np.array_equal(reflib.lm_update(J, res, 0.01), oracle.lm_update(J, res, 0.01))


def test_craig_sneyd(oracle, reflib):
    pr, U = reflib.host_scheme(1, K=100.0, T=1.0, m1=40, m2=20, N=11, theta=0.8, want_U=True, **BASE)
    o = oracle.solve(100.0, 11, 1.0 / 11, m1=40, m2=20, theta=0.8, scheme=1, **BASE)
    assert pr == o["price"] and np.array_equal(U, o["U"])
