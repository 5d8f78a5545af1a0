"""TEST INFRASTRUCTURE — ctypes bindings for oracle/_ref/libhadi_ref{,_omp}.so (the reference's own
sources compiled against oracle/kokkos_shim, see oracle/ref_driver.cpp) and for
oracle/libhadi_oracle.so (the plain-C restatement, oracle/hadi_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing in the product package imports it.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def ref_path(omp=False):
    return os.path.join(_HERE, "_ref", "libhadi_ref_omp.so" if omp else "libhadi_ref.so")


def have_ref(omp=False):
    return os.path.exists(ref_path(omp))


class RefLib:
    """The real reference (unmodified sources) behind a C ABI."""

    def __init__(self, omp=False):
        self.lib = C.CDLL(ref_path(omp))
        L = self.lib
        L.hadi_ref_grid.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                    C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, _dp]
        L.hadi_ref_solve_batch.argtypes = (
            [C.c_int, _dp, _dp, _ip, _dp] + [C.c_double] * 8 + [C.c_int, C.c_int, C.c_double, C.c_int,
                                                               C.c_int, C.c_int, _dp, _dp, _dp, C.c_int,
                                                               C.c_int, C.c_double, C.c_double, _dp, _dp,
                                                               _dp, _dp])
        L.hadi_ref_lm_update.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp]
        L.hadi_ref_solve5.argtypes = [_dp, _dp, _dp]
        L.hadi_ref_bs_call.argtypes = [C.c_double] * 5
        L.hadi_ref_bs_call.restype = C.c_double
        L.hadi_ref_host_scheme.argtypes = ([C.c_int] + [C.c_double] * 10 + [C.c_int, C.c_int, C.c_int,
                                                                            C.c_double, _dp, _dp])
        L.hadi_ref_dump_operators.argtypes = ([C.c_double, C.c_int] + [C.c_double] * 9 +
                                              [C.c_int, C.c_int, C.c_double] + [_dp] * 8)
        L.hadi_ref_run_shipped.argtypes = [C.c_int]
        L.hadi_ref_last_compute_seconds.restype = C.c_double
        L.hadi_ref_set_threads.argtypes = [C.c_int]

    def set_threads(self, n):
        """Host threads of the league loop (overrides OMP_NUM_THREADS); returns the count in effect."""
        return int(self.lib.hadi_ref_set_threads(int(n)))

    def threads(self):
        return int(self.lib.hadi_ref_threads())

    def last_compute_seconds(self):
        """Wall time of the reference entry point alone in the last solve_batch (driver set-up excluded)."""
        return float(self.lib.hadi_ref_last_compute_seconds())

    def grid(self, m1, m2, K, S0, V0, S=None, c=None, V=5.0, d=5.0 / 500):
        S = 8 * K if S is None else S
        c = K / 5 if c is None else c
        s, ds = np.zeros(m1 + 1), np.zeros(m1)
        v, dv = np.zeros(m2 + 1), np.zeros(m2)
        self.lib.hadi_ref_grid(m1, S, S0, K, c, m2, V, V0, d, _d(s), _d(ds), _d(v), _d(dv))
        return s, ds, v, dv

    def solve_batch(self, strikes, N, dt, *, S0, V0, r_d, r_f, rho, sigma, kappa, eta, m1, m2, theta,
                    style=0, payoff_put=0, divs=None, multi=0, jac=0, eps=1e-6, maturities=None,
                    want_U=False, want_lambda=False, V0_grid=None):
        strikes = np.ascontiguousarray(strikes, dtype=np.float64)
        n = strikes.size
        Ns = np.ascontiguousarray(np.broadcast_to(np.asarray(N, dtype=np.int32), (n,)))
        dts = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (n,)))
        mats = None if maturities is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(maturities, dtype=np.float64), (n,)))
        P = (m1 + 1) * (m2 + 1)
        prices = np.zeros(n)
        J = np.zeros((n, 5)) if jac else None
        U = np.zeros((n, P)) if want_U else None
        lam = np.zeros((n, P)) if want_lambda else None
        if divs is not None and len(divs[0]) > 0:
            dd, da, dp = (np.ascontiguousarray(x, dtype=np.float64) for x in divs)
            nd = dd.size
        else:
            dd = da = dp = None
            nd = 0
        rc = self.lib.hadi_ref_solve_batch(
            n, _d(strikes), _d(mats), _i(Ns), _d(dts), S0, V0, r_d, r_f, rho, sigma, kappa, eta, m1, m2,
            theta, style, payoff_put, nd, _d(dd), _d(da), _d(dp), multi, jac, eps,
            V0 if V0_grid is None else V0_grid, _d(prices), _d(J), _d(U), _d(lam))
        if rc != 0:
            raise RuntimeError("hadi_ref_solve_batch rc=%d" % rc)
        out = {"prices": prices}
        if jac:
            out["J"] = J
        if want_U:
            out["U"] = U
        if want_lambda:
            out["lambda"] = lam
        return out

    def lm_update(self, J, r, lam):
        J = np.ascontiguousarray(J, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        delta = np.zeros(5)
        self.lib.hadi_ref_lm_update(r.size, _d(J), _d(r), lam, _d(delta))
        return delta

    def solve5(self, A, b):
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(5)
        self.lib.hadi_ref_solve5(_d(A), _d(b), _d(x))
        return x

    def bs_call(self, S, K, r, vol, T):
        return self.lib.hadi_ref_bs_call(S, K, r, vol, T)

    def bs_vega(self, S, K, r, vol, T):
        self.lib.hadi_ref_bs_vega.restype = C.c_double
        self.lib.hadi_ref_bs_vega.argtypes = [C.c_double] * 5
        return self.lib.hadi_ref_bs_vega(S, K, r, vol, T)

    def reverse_bs(self, S, K, r, T, v0, target, eps):
        self.lib.hadi_ref_reverse_bs.restype = C.c_double
        self.lib.hadi_ref_reverse_bs.argtypes = [C.c_double] * 7
        return self.lib.hadi_ref_reverse_bs(S, K, r, T, v0, target, eps)

    def reverse_bs_dic(self, S, K, r, T, target, eps, a, b):
        self.lib.hadi_ref_reverse_bs_dic.restype = C.c_double
        self.lib.hadi_ref_reverse_bs_dic.argtypes = [C.c_double] * 8
        return self.lib.hadi_ref_reverse_bs_dic(S, K, r, T, target, eps, a, b)

    def market(self, S0, T, r_d, strikes, divs=None):
        k = np.ascontiguousarray(strikes, dtype=np.float64)
        out = np.zeros(k.size)
        if divs:
            dd, da, dp = (np.ascontiguousarray(x, dtype=np.float64) for x in divs)
            nd = dd.size
        else:
            dd = da = dp = None
            nd = 0
        self.lib.hadi_ref_market.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp, C.c_int, _dp, _dp,
                                             _dp, _dp]
        self.lib.hadi_ref_market(S0, T, r_d, k.size, _d(k), nd, _d(dd), _d(da), _d(dp), _d(out))
        return out

    def host_scheme(self, scheme, *, K, S0, V0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2, N, theta,
                    want_U=False):
        price = C.c_double(0.0)
        U = np.zeros((m1 + 1) * (m2 + 1)) if want_U else None
        rc = self.lib.hadi_ref_host_scheme(scheme, K, S0, V0, T, r_d, r_f, rho, sigma, kappa, eta, m1,
                                           m2, N, theta, C.byref(price), _d(U))
        if rc != 0:
            raise RuntimeError("hadi_ref_host_scheme rc=%d" % rc)
        return (price.value, U) if want_U else price.value

    def dump_operators(self, *, K, N, dt, S0, V0, r_d, r_f, rho, sigma, kappa, eta, m1, m2, theta):
        P = (m1 + 1) * (m2 + 1)
        a1l, a1m, a1u = np.zeros(P), np.zeros(P), np.zeros(P)
        a2 = np.zeros(5 * (m2 + 1))
        b, b1, b2 = np.zeros(P), np.zeros(P), np.zeros(P)
        a0 = np.zeros((m2 - 1) * (m1 - 1) * 9)
        self.lib.hadi_ref_dump_operators(K, N, dt, S0, V0, r_d, r_f, rho, sigma, kappa, eta, m1, m2,
                                         theta, _d(a1l), _d(a1m), _d(a1u), _d(a2), _d(b), _d(b1), _d(b2),
                                         _d(a0))
        return dict(a1_lower=a1l.reshape(m2 + 1, m1 + 1), a1_main=a1m.reshape(m2 + 1, m1 + 1),
                    a1_upper=a1u.reshape(m2 + 1, m1 + 1), a2=a2.reshape(5, m2 + 1), b=b, b1=b1, b2=b2,
                    a0=a0.reshape(m2 - 1, (m1 - 1) * 9))

    def run_shipped(self, which):
        return self.lib.hadi_ref_run_shipped(which)


# --------------------------------------------------------------------------- plain-C restatement
class _Model(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("S0", "V0", "r_d", "r_f", "kappa", "eta", "sigma", "rho")]


class _Numerics(C.Structure):
    _fields_ = [("m1", C.c_int), ("m2", C.c_int), ("theta", C.c_double), ("style", C.c_int),
                ("payoff_put", C.c_int), ("scheme", C.c_int), ("nd", C.c_int), ("div_dates", _dp),
                ("div_amounts", _dp), ("div_pcts", _dp), ("bc", C.c_int), ("div_all", C.c_int)]


class _LmOpts(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("tol", C.c_double), ("delta_tol", C.c_double),
                ("lambda0", C.c_double), ("eps", C.c_double)]


class _LmResult(C.Structure):
    _fields_ = [("params", C.c_double * 5), ("final_error", C.c_double), ("lambda_", C.c_double),
                ("delta_norm", C.c_double), ("iterations", C.c_int), ("converged", C.c_int),
                ("pde_solves", C.c_int)]


def oracle_path():
    return os.path.join(_HERE, "libhadi_oracle.so")


class OracleLib:
    """oracle/hadi_oracle.c — the CPU restatement ("port")."""

    def __init__(self):
        self.lib = C.CDLL(oracle_path())
        L = self.lib
        L.ho_grid_s.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.ho_grid_v.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.ho_solve.argtypes = [C.POINTER(_Model), C.POINTER(_Numerics), C.c_double, C.c_int, C.c_double,
                               C.c_double, _dp, _dp, _dp]
        L.ho_price_batch.argtypes = [C.POINTER(_Model), C.POINTER(_Numerics), C.c_int, _dp, _ip, _dp, _dp]
        L.ho_jacobian_batch.argtypes = [C.POINTER(_Model), C.POINTER(_Numerics), C.c_int, _dp, _ip, _dp,
                                        C.c_double, _dp, _dp]
        L.ho_solve5.argtypes = [_dp, _dp, _dp]
        L.ho_lm_update.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp]
        L.ho_calibrate.argtypes = [C.POINTER(_Model), C.POINTER(_Numerics), C.c_int, _dp, _ip, _dp, _dp,
                                   C.POINTER(_LmOpts), C.POINTER(_LmResult)]
        L.ho_bs_call.argtypes = [C.c_double] * 5
        L.ho_bs_call.restype = C.c_double

    @staticmethod
    def _mk(S0, V0, r_d, r_f, rho, sigma, kappa, eta, m1, m2, theta, style=0, payoff_put=0, scheme=0,
            divs=None, bc=0, div_all=0):
        mdl = _Model(S0, V0, r_d, r_f, kappa, eta, sigma, rho)
        keep = []
        if divs is not None and len(divs[0]) > 0:
            arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in divs]
            keep = arrs
            num = _Numerics(m1, m2, theta, style, payoff_put, scheme, arrs[0].size, _d(arrs[0]),
                            _d(arrs[1]), _d(arrs[2]), bc, div_all)
        else:
            num = _Numerics(m1, m2, theta, style, payoff_put, scheme, 0, None, None, None, bc, div_all)
        return mdl, num, keep

    def grid(self, m1, m2, K, S0, V0):
        s, ds = np.zeros(m1 + 1), np.zeros(m1)
        v, dv = np.zeros(m2 + 1), np.zeros(m2)
        self.lib.ho_grid_s(m1, 8 * K, S0, K, K / 5, _d(s), _d(ds))
        self.lib.ho_grid_v(m2, 5.0, V0, 5.0 / 500, _d(v), _d(dv))
        return s, ds, v, dv

    def solve(self, K, N, dt, want_U=True, want_lambda=True, **kw):
        mdl, num, keep = self._mk(**kw)
        P = (num.m1 + 1) * (num.m2 + 1)
        price = C.c_double(0.0)
        U = np.zeros(P) if want_U else None
        lam = np.zeros(P) if want_lambda else None
        rc = self.lib.ho_solve(C.byref(mdl), C.byref(num), K, N, dt, mdl.V0, C.byref(price), _d(U),
                               _d(lam))
        if rc != 0:
            raise RuntimeError("ho_solve rc=%d" % rc)
        return {"price": price.value, "U": U, "lambda": lam}

    @staticmethod
    def _batch_args(strikes, N, dt):
        strikes = np.ascontiguousarray(strikes, dtype=np.float64)
        n = strikes.size
        Ns = np.ascontiguousarray(np.broadcast_to(np.asarray(N, dtype=np.int32), (n,)))
        dts = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (n,)))
        return strikes, n, Ns, dts

    def price_batch(self, strikes, N, dt, **kw):
        mdl, num, keep = self._mk(**kw)
        strikes, n, Ns, dts = self._batch_args(strikes, N, dt)
        prices = np.zeros(n)
        rc = self.lib.ho_price_batch(C.byref(mdl), C.byref(num), n, _d(strikes), _i(Ns), _d(dts),
                                     _d(prices))
        if rc != 0:
            raise RuntimeError("ho_price_batch rc=%d" % rc)
        return prices

    def jacobian_batch(self, strikes, N, dt, eps=1e-6, **kw):
        mdl, num, keep = self._mk(**kw)
        strikes, n, Ns, dts = self._batch_args(strikes, N, dt)
        J, base = np.zeros((n, 5)), np.zeros(n)
        rc = self.lib.ho_jacobian_batch(C.byref(mdl), C.byref(num), n, _d(strikes), _i(Ns), _d(dts), eps,
                                        _d(J), _d(base))
        if rc != 0:
            raise RuntimeError("ho_jacobian_batch rc=%d" % rc)
        return J, base

    def lm_update(self, J, r, lam):
        J = np.ascontiguousarray(J, dtype=np.float64)
        r = np.ascontiguousarray(r, dtype=np.float64)
        delta = np.zeros(5)
        self.lib.ho_lm_update(r.size, _d(J), _d(r), lam, _d(delta))
        return delta

    def solve5(self, A, b):
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(5)
        self.lib.ho_solve5(_d(A), _d(b), _d(x))
        return x

    def calibrate(self, strikes, N, dt, market, *, max_iter, tol, delta_tol, lambda0=0.01, eps=1e-6, **kw):
        mdl, num, keep = self._mk(**kw)
        strikes, n, Ns, dts = self._batch_args(strikes, N, dt)
        market = np.ascontiguousarray(market, dtype=np.float64)
        opt = _LmOpts(max_iter, tol, delta_tol, lambda0, eps)
        res = _LmResult()
        self.lib.ho_calibrate(C.byref(mdl), C.byref(num), n, _d(strikes), _i(Ns), _d(dts), _d(market),
                              C.byref(opt), C.byref(res))
        return dict(params=list(res.params), final_error=res.final_error, lam=res.lambda_,
                    delta_norm=res.delta_norm, iterations=res.iterations, converged=res.converged,
                    pde_solves=res.pde_solves)

    def bs_call(self, S, K, r, vol, T):
        return self.lib.ho_bs_call(S, K, r, vol, T)
