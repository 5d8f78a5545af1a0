"""Host layer of libhadi.so (no GPU needed): grids, Black-Scholes helper, LM normal equations, the
Jacobian assembly and the multi-GPU partitioner, against the oracle and the golden vectors."""
import numpy as np
import pytest

from conftest import BASE, golden


def test_grid_matches_reference(hadi):
    for g in golden("grids.json"):
        s, v = hadi.grid(g["m1"], g["m2"], g["K"], g["S0"], g["V0"])
        assert [repr(float(x)) for x in s] == g["s"]
        assert [repr(float(x)) for x in v] == g["v"]


def test_bs_call(hadi, oracle):
    for c in golden("bs.json"):
        assert repr(hadi.bs_call(c["S"], c["K"], c["r"], c["vol"], c["T"])) == c["price"]


def test_lm_update_and_solve5(hadi):
    G = golden("lm_update.json")
    J = np.array([[float(x) for x in row] for row in G["J"]])
    r = np.array([float(x) for x in G["r"]])
    for c in G["cases"]:
        assert [repr(float(x)) for x in hadi.lm_update(J, r, c["lam"])] == c["delta"]
    A = np.array([[float(x) for x in row] for row in G["solve5"]["A"]])
    b = np.array([float(x) for x in G["solve5"]["b"]])
    assert [repr(float(x)) for x in hadi.solve5(A, b)] == G["solve5"]["x"]


def test_solve5_pivoting(hadi, oracle):
    # a zero leading entry forces a row swap (src/jacobian_computation.cpp:44-66)
    A = np.array([[0.0, 2, 1, 0, 1], [3, 1, 0, 2, 0], [1, 0, 4, 1, 1], [0, 1, 1, 5, 0], [2, 0, 0, 1, 6.0]])
    b = np.arange(1.0, 6.0)
    x = hadi.solve5(A, b)
    assert np.array_equal(x, oracle.solve5(A, b))
    assert np.allclose(A @ x, b, atol=1e-12)


def test_jacobian_assemble(hadi):
    rng = np.random.default_rng(3)
    v = rng.normal(size=6 * 7)
    J, base = hadi.jacobian_assemble(v, 1e-6)
    for k in range(7):
        assert base[k] == v[6 * k]
        for c in range(5):
            assert J[k, c] == (v[6 * k + 1 + c] - v[6 * k]) / 1e-6


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_covers_everything_once(hadi, world):
    rng = np.random.default_rng(world)
    for n in (0, 1, 5, 97, 600):
        costs = rng.integers(20, 200, size=n).astype(np.int32) * 5151
        prev_end = 0
        loads = []
        for r in range(world):
            b, e = hadi.partition(costs, world, r)
            assert b == prev_end and e >= b
            prev_end = e
            loads.append(int(costs[b:e].sum()))
        assert prev_end == n
        if n >= 8 * world:
            assert max(loads) <= 1.0 * sum(loads) / world + int(costs.max())  # balanced by cost


def test_item_costs(hadi):
    num = hadi.make_numerics(100, 50, 0.8)
    pts, n = hadi.make_points([90.0, 100.0, 110.0], [1.0, 2.0, 0.5], [20, 40, 20])
    c1 = hadi.item_costs(num, pts, n, hadi.MODE_PRICE)
    c6 = hadi.item_costs(num, pts, n, hadi.MODE_JACOBIAN)
    assert list(c1) == [20 * 5151, 40 * 5151, 20 * 5151]
    assert list(c6) == [x for x in c1 for _ in range(6)]


def test_bad_arguments_are_rejected_without_a_gpu(hadi):
    import ctypes

    L = hadi.lib()
    assert L.hadi_solve5(None, None, None) == hadi.ERR_ARG
    b, e = ctypes.c_int(), ctypes.c_int()
    assert L.hadi_partition(5, None, 0, 0, ctypes.byref(b), ctypes.byref(e)) == hadi.ERR_ARG
    assert L.hadi_price_batch(None, None, None, 0, None, None, None, None) == hadi.ERR_ARG


def test_market_generators_and_implied_vol(hadi):
    """synthetic market (plain / dividend-adjusted spot), vega, Newton and bisection
    implied-vol inversion, bit-equal to the reference's BlackScholes class (src/bs.hpp:58-192)."""
    G = golden("market.json")
    for c in G["market"]:
        n = len(c["strikes"])
        pts, _ = hadi.make_points(c["strikes"], c["T"], 20)
        plain = hadi.market_prices(c["S0"], c["r_d"], 0.2, pts, n)
        div = hadi.market_prices(c["S0"], c["r_d"], 0.2, pts, n, divs=G["divs"])
        assert [repr(float(x)) for x in plain] == c["plain"]
        assert [repr(float(x)) for x in div] == c["dividends"]
        # the dividend-adjusted spot is what the generator priced at
        sa = hadi.dividend_adjusted_spot(c["S0"], c["T"], c["r_d"], G["divs"])
        assert repr(hadi.bs_call(sa, c["strikes"][0], c["r_d"], 0.2, c["T"])) == c["dividends"][0]
    for c in G["implied_vol"]:
        t = float(c["target"])
        assert repr(hadi.bs_implied_vol(c["S"], c["K"], c["r"], c["T"], 0.5, t, c["eps"])) == c["newton"]
        assert repr(hadi.bs_implied_vol_bisect(c["S"], c["K"], c["r"], c["T"], t, c["eps"], 0.001, 1.0)) == c["bisect"]
        vol = 0.5 if c.get("fallback") else 0.2
        assert repr(hadi.bs_vega(c["S"], c["K"], c["r"], vol, c["T"])) == c["vega"]


def test_market_helpers_against_the_reference(hadi, reflib):
    rng = np.random.default_rng(11)
    for _ in range(40):
        S, K = 100.0 * (1 + 0.1 * rng.normal()), 100.0 * (1 + 0.2 * rng.normal())
        T, vol = float(rng.uniform(0.1, 3.0)), float(rng.uniform(0.08, 0.6))
        target = hadi.bs_call(S, K, 0.025, vol, T)
        assert hadi.bs_vega(S, K, 0.025, vol, T) == reflib.bs_vega(S, K, 0.025, vol, T)
        assert hadi.bs_implied_vol(S, K, 0.025, T, 0.5, target, 0.01) == reflib.reverse_bs(S, K, 0.025, T, 0.5, target, 0.01)
        assert (hadi.bs_implied_vol_bisect(S, K, 0.025, T, target, 1e-9, 0.001, 1.0)
                == reflib.reverse_bs_dic(S, K, 0.025, T, target, 1e-9, 0.001, 1.0))


def test_calibration_report_writer(hadi, tmp_path):
    """The CSV the reference's LM drivers export (src/heston_calibration.cpp:467-508, 2857-2921): same header
    fields, same columns, doubles streamed at the default precision."""
    strikes = [95.0 + s for s in range(4)]
    mats = [1.0, 1.5]
    K = [k for _ in mats for k in strikes]
    T = [t for t in mats for _ in strikes]
    pts, n = hadi.make_points(K, T, [max(20, int(20 * t)) for t in T])
    market = hadi.market_prices(100.0, 0.025, 0.2, pts, n)
    fitted = market + 0.05
    initial = hadi.make_model(100.0, 0.04, 0.025, 0.0, 1.5, 0.04, 0.3, -0.9)
    res = dict(params=[4.5, 0.04, 0.1, -0.11, 0.042], final_error=1.25, iterations=3, pde_solves=160)
    p1 = str(tmp_path / "multi.csv")
    hadi.write_calibration_csv(p1, 1, 100.0, 0.025, len(mats), len(strikes), pts, market, fitted, initial, res, 0.5)
    lines = open(p1).read().splitlines()
    assert lines[0] == ("# Calibration with 2 maturities, 4 strikes per maturity, Time=0.5 s, FinalError=1.25, "
                        "IterationCount=3, TotalPdeSolves=160, init_kappa=1.5, init_eta=0.04, init_sigma=0.3, "
                        "init_rho=-0.9, init_v0=0.04, kappa=4.5, eta=0.04, sigma=0.1, rho=-0.11, v0=0.042")
    assert lines[1] == "Maturity,Strike,MarketPrice,FittedPrice,MarketIV,FittedIV,IVDifference"
    assert len(lines) == 2 + n
    miv, fiv, dif = hadi.implied_vols(100.0, 0.025, pts, n, market, fitted, 0.01)
    row = lines[2 + 5].split(",")
    assert row[0] == "1.5" and row[1] == "96"
    assert row[2] == "%g" % market[5] and row[3] == "%g" % fitted[5]
    assert row[4] == "%g" % miv[5] and row[5] == "%g" % fiv[5] and row[6] == "%g" % dif[5]
    # single-maturity format
    pts1, n1 = hadi.make_points(strikes, 1.0, 20)
    p0 = str(tmp_path / "single.csv")
    hadi.write_calibration_csv(p0, 0, 100.0, 0.025, 1, n1, pts1, market[:n1], fitted[:n1], initial, res, 0.25)
    lines = open(p0).read().splitlines()
    assert lines[0].startswith("# 4 options, Time=0.25 s, FinalError=1.25, iterationCount=3, TotalPdeSolves=160, init_kappa=1.5")
    assert lines[1] == "Strike,MarketPrice,FittedPrice,IVDifference" and len(lines) == 2 + n1
    assert lines[2].split(",")[0] == "95"


def test_jacobian_assembly_of_every_mode(hadi):
    """Layouts of include/hadi.h: forward (6 values per option), interpolated V0 (5 items x 3 values),
    central (11 values); and the v-bracket of the interpolation (src/device_solver.cpp:1735-1754)."""
    rng = np.random.default_rng(5)
    n, eps5 = 4, np.array([1e-6, 2e-6, 3e-6, 4e-6, 5e-6])
    v = rng.normal(size=6 * n)
    J, base = hadi.jacobian_assemble_ex(v, hadi.MODE_JACOBIAN, eps5)
    for k in range(n):
        assert base[k] == v[6 * k]
        for c in range(5):
            assert J[k, c] == (v[6 * k + 1 + c] - v[6 * k]) / eps5[c]
    J1, base1 = hadi.jacobian_assemble(v, 1e-6)
    J2, base2 = hadi.jacobian_assemble_ex(v, hadi.MODE_JACOBIAN, 1e-6)
    assert np.array_equal(J1, J2) and np.array_equal(base1, base2)
    v = rng.normal(size=15 * n)
    w = 0.37
    J, base = hadi.jacobian_assemble_ex(v, hadi.MODE_JACOBIAN_INTERP, eps5, w)
    for k in range(n):
        o = v[15 * k:15 * k + 15]
        assert base[k] == o[0]
        for c in range(4):
            assert J[k, c] == (o[3 * (1 + c)] - o[0]) / eps5[c]
        assert J[k, 4] == ((o[1] + w * (o[2] - o[1])) - o[0]) / eps5[4]
    v = rng.normal(size=11 * n)
    J, base = hadi.jacobian_assemble_ex(v, hadi.MODE_JACOBIAN_CENTRAL, eps5)
    for k in range(n):
        assert base[k] == v[11 * k]
        for c in range(5):
            assert J[k, c] == (v[11 * k + 1 + c] - v[11 * k + 6 + c]) / (2.0 * eps5[c])
    _, vg = hadi.grid(50, 25, 100.0, 100.0, 0.04)
    lo, hi, wt = hadi.jacobian_v0_weight(25, 0.04, 1e-6)
    assert vg[lo] == 0.04 and hi == lo + 1 and wt == (0.04 + 1e-6 - vg[lo]) / (vg[hi] - vg[lo])
    # a bump beyond the last v-node finds no bracket: the reference leaves index 0 / weight 0
    assert hadi.jacobian_v0_weight(25, 0.04, 10.0) == (0, 0, 0.0)
    num = hadi.make_numerics(100, 50, 0.8)
    pts, n3 = hadi.make_points([90.0, 100.0], [1.0, 2.0], [20, 40])
    assert hadi.item_costs(num, pts, n3, hadi.MODE_JACOBIAN_INTERP).size == 10
    assert hadi.item_costs(num, pts, n3, hadi.MODE_JACOBIAN_CENTRAL).size == 22


@pytest.mark.parametrize("case", ["config2", "mixed", "barely", "chain", "tiny_steps"])
def test_split_schedule_covers_every_step_once(hadi, case):
    """The split schedule of batches larger than the persistent grid (McNaughton's wrap-around rule): every time
    step of every solve appears exactly once, a solve is cut at most once, the first steps of a cut solve open
    the list of the CTA after the one its last steps close, hand-off slots pair up, and the heaviest CTA
    carries (nearly) total / slots instead of a whole number of solves."""
    ts, slots = dict(config2=([50] * 500, 296), mixed=(sorted([20 + (i % 3) * 5 for i in range(1000)], reverse=True), 444),
                     barely=([50] * 297, 296), chain=(sorted([max(20, int(20 * t)) for t in (0.25, 0.5, 1, 2)] * 2500,
                                                             reverse=True), 296),
                     tiny_steps=([4] * 40, 16))[case]
    segs, off, heaviest = hadi.plan_schedule(ts, slots)
    if case == "tiny_steps":
        assert segs == []   # solves of fewer than six steps are never cut: the CTAs pull whole items
        return
    assert segs, "a batch larger than the grid must be cut"
    assert off[0] == 0 and off[slots] == len(segs) and all(off[b] <= off[b + 1] for b in range(slots))
    slot_of = {}
    for b in range(slots):
        for q in range(off[b], off[b + 1]):
            slot_of[q] = b
    steps = [[] for _ in ts]
    producers, consumers = {}, {}
    for q, (item, n0, n1, hin, hout) in enumerate(segs):
        assert 1 <= n0 <= n1 <= ts[item]
        steps[item].append((n0, n1, q))
        if hout >= 0:
            assert n0 == 1 and n1 < ts[item] and hout not in producers
            producers[hout] = q
            assert q == off[slot_of[q]], "the first steps of a cut solve open their CTA's list"
        if hin >= 0:
            assert n0 > 1 and n1 == ts[item] and hin not in consumers
            consumers[hin] = q
            assert q == off[slot_of[q] + 1] - 1, "the last steps of a cut solve close their CTA's list"
    assert sorted(producers) == sorted(consumers) == list(range(len(producers)))
    for h, qp in producers.items():
        qc = consumers[h]
        assert segs[qp][0] == segs[qc][0] and segs[qp][2] + 1 == segs[qc][1]
        assert slot_of[qp] == slot_of[qc] + 1
    for item, parts in enumerate(steps):
        parts.sort()
        assert len(parts) in (1, 2) and parts[0][0] == 1 and parts[-1][1] == ts[item]
        if len(parts) == 2:
            assert parts[0][1] + 1 == parts[1][0]
    ideal = (sum(ts) + 1.5 * len(ts)) / slots
    whole = -(-len(ts) // slots) * (max(ts) + 1.5)
    assert heaviest <= max(1.06 * ideal + 4.0, max(ts) + 1.5) and heaviest <= whole
    # nothing to cut when every solve has its own CTA
    assert hadi.plan_schedule([50] * 200, 296)[0] == []
