"""TEST INFRASTRUCTURE — ctypes driver for tests/emu/libhadi_emu.so (CPU emulation of the kernel phases)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "pde-based-heston-solver-gpu-accelerated_b200", "csrc")
_dp = C.POINTER(C.c_double)


class Item(C.Structure):
    _fields_ = ([(k, C.c_double) for k in ("kappa", "eta", "sigma", "rho", "r_d", "r_f", "dt", "theta", "K", "ef")] +
                [(k, C.c_int) for k in ("N", "style", "payoff", "nd", "s_off", "v_off", "e_off", "idx_s", "idx_v",
                                        "out", "cost", "aux", "bc", "div_all")])


def build():
    so = os.path.join(_HERE, "libhadi_emu.so")
    srcs = [os.path.join(_HERE, "hadi_emu.cpp"), os.path.join(_CSRC, "hadi_phases.cuh"),
            os.path.join(_CSRC, "hadi_phases_cs.cuh")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I" + _CSRC,
                               srcs[0], "-o", so])
    return so


def emu_solve(hadi, K, N, dt, *, S0, V0, r_d, r_f, rho, sigma, kappa, eta, m1, m2, theta, style=0, payoff_put=0,
              divs=None, nt=384, V0_grid=None, scheme=0):
    """Run one solve through the emulated kernel.  `hadi` is the product's python binding (for hadi_grid)."""
    L = C.CDLL(build())
    assert L.hadi_emu_item_size() == C.sizeof(Item)
    Vg = V0 if V0_grid is None else V0_grid
    s, v = hadi.grid(m1, m2, K, S0, Vg)
    idx_s = next(i for i in range(m1 + 1) if abs(s[i] - S0) < 1e-10)
    idx_v = next((j for j in range(m2 + 1) if abs(v[j] - Vg) < 1e-10), 0)
    E = np.array([math.exp(r_f * dt * n) for n in range(N + 1)])
    ef = math.exp(-r_f * dt * (N - 1))
    if divs is not None and len(divs[0]) > 0:
        dd, da, dpc = (np.ascontiguousarray(x, dtype=np.float64) for x in divs)
        nd = dd.size
    else:
        dd = da = dpc = np.zeros(1)
        nd = 0
    it = Item(kappa, eta, sigma, rho, r_d, r_f, dt, theta, K, ef, N, style, payoff_put, nd, 0, 0, 0, idx_s, idx_v, 0,
              0, 0)
    P = (m1 + 1) * (m2 + 1)
    U, lam = np.zeros(P), np.zeros(P)
    price = C.c_double(0.0)
    f = lambda a: a.ctypes.data_as(_dp)
    if scheme == 1:
        rc = L.hadi_emu_solve_cs(C.byref(it), m1, m2, nt, f(s), f(v), f(E), C.byref(price), f(U))
    else:
        rc = L.hadi_emu_solve(C.byref(it), m1, m2, nt, f(s), f(v), f(E), nd, f(dd), f(da), f(dpc), C.byref(price),
                              f(U), f(lam))
    assert rc == 0
    return {"price": price.value, "U": U, "lambda": lam}
