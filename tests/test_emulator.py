"""The kernel's device code (csrc/hadi_phases.cuh), compiled for the host and run phase by phase with
a serial loop over thread ids, against the oracle: checks the CUDA path's index logic and operation
order without a GPU.  Bit-exact on the full grid, the multiplier and the exercise region."""
import numpy as np
import pytest

import emu
from conftest import BASE, DIVS


@pytest.mark.parametrize("m1,m2,N,nt", [(50, 25, 8, 256), (100, 50, 6, 320), (20, 10, 7, 128), (64, 32, 5, 416),
                                        (40, 40, 4, 1024)])
def test_emulated_kernel_matches_oracle(hadi, oracle, m1, m2, N, nt):
    for style in (0, 1):
        for dv in (None, DIVS):
            for put in (0, 1):
                b = dict(BASE)
                b["r_f"] = 0.01 if put else 0.0
                o = oracle.solve(93.0, N, 1.0 / N, m1=m1, m2=m2, theta=0.8, style=style, divs=dv, payoff_put=put, **b)
                e = emu.emu_solve(hadi, 93.0, N, 1.0 / N, m1=m1, m2=m2, theta=0.8, style=style, divs=dv,
                                  payoff_put=put, nt=nt, **b)
                assert o["price"] == e["price"]
                assert np.array_equal(o["U"], e["U"])
                if style:
                    assert np.array_equal(o["lambda"], e["lambda"])


def test_emulated_kernel_v0_plus_eps_grid(hadi, oracle):
    """The Jacobian's V0 column runs on the grid rebuilt for V0 + eps (src/jacobian_computation.cpp:339)."""
    b = dict(BASE)
    b["V0"] = BASE["V0"] + 1e-6
    o = oracle.solve(100.0, 6, 1.0 / 6, m1=30, m2=15, theta=0.8, **b)
    e = emu.emu_solve(hadi, 100.0, 6, 1.0 / 6, m1=30, m2=15, theta=0.8, nt=128, **b)
    assert o["price"] == e["price"] and np.array_equal(o["U"], e["U"])


@pytest.mark.parametrize("m1,m2,N,nt", [(50, 25, 6, 256), (30, 15, 5, 128), (64, 32, 4, 1024)])
def test_emulated_craig_sneyd_matches_oracle(hadi, oracle, m1, m2, N, nt):
    """Craig-Sneyd stages (csrc/hadi_phases_cs.cuh) against the oracle's restatement of CS_scheme_shuffled."""
    for put in (0, 1):
        b = dict(BASE)
        b["r_f"] = 0.01 if put else 0.0
        o = oracle.solve(97.0, N, 1.0 / N, m1=m1, m2=m2, theta=0.8, scheme=1, payoff_put=put, want_lambda=False, **b)
        e = emu.emu_solve(hadi, 97.0, N, 1.0 / N, m1=m1, m2=m2, theta=0.8, payoff_put=put, nt=nt, scheme=1, **b)
        assert o["price"] == e["price"]
        assert np.array_equal(o["U"], e["U"])
