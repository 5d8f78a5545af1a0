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
