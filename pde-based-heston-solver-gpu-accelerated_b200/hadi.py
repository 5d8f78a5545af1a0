"""ctypes binding of libhadi.so (the C ABI declared in include/hadi.h).

This module is plumbing for tests/ and bench.py: it loads the in-tree shared library, mirrors the
C structs and turns error codes into exceptions.  All numerical work happens in the CUDA kernel
behind the C ABI; there is no Python or CPU fallback — if the library (or a CUDA device) is missing
the calls fail loudly.
"""
import ctypes as C
import weakref
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# development aid: HADI_LIB=<file name in this directory> loads an experiment build (csrc/Makefile: make exp)
LIB_PATH = os.path.join(_HERE, os.environ.get("HADI_LIB", "libhadi.so"))

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

OK, ERR_ARG, ERR_GRID, ERR_CUDA, ERR_SMEM, ERR_COMM, ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6
EUROPEAN, AMERICAN = 0, 1
CALL, PUT = 0, 1
DOUGLAS, CRAIG_SNEYD, MODIFIED_CRAIG_SNEYD, HUNDSDORFER_VERWER = 0, 1, 2, 3
BC_REFERENCE_CALL, BC_PUT = 0, 1
DIVIDENDS_DEVICE, DIVIDENDS_ALL = 0, 1
MODE_PRICE, MODE_JACOBIAN, MODE_JACOBIAN_INTERP, MODE_JACOBIAN_CENTRAL = 0, 1, 2, 3
ITEMS_PER_OPTION = {0: 1, 1: 6, 2: 5, 3: 11}
VALUES_PER_ITEM = {0: 1, 1: 1, 2: 3, 3: 1}

# every symbol include/hadi.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "hadi_create", "hadi_destroy", "hadi_last_error", "hadi_version", "hadi_kernel_launches",
    "hadi_price_batch", "hadi_jacobian_batch", "hadi_batch_create", "hadi_batch_num_items",
    "hadi_batch_launch", "hadi_batch_values_dev", "hadi_batch_fetch", "hadi_batch_elapsed_ms",
    "hadi_batch_destroy", "hadi_jacobian_assemble", "hadi_partition", "hadi_item_costs", "hadi_solve5",
    "hadi_lm_update", "hadi_calibrate", "hadi_grid", "hadi_bs_call", "hadi_transfer_bytes", "hadi_measure_fp64",
    "hadi_batch_phase_cycles", "hadi_bs_vega", "hadi_bs_implied_vol", "hadi_bs_implied_vol_bisect",
    "hadi_dividend_adjusted_spot", "hadi_market_prices", "hadi_implied_vols", "hadi_write_calibration_csv",
    "hadi_batch_create_ex", "hadi_batch_values_per_item", "hadi_jacobian_assemble_ex", "hadi_jacobian_v0_weight",
    "hadi_jacobian_batch_ex", "hadi_calibrate_ex", "hadi_plan_schedule", "hadi_exact_reruns",
    "hadi_batch_exact_reruns", "hadi_batch_kernel_info", "hadi_nccl_unique_id", "hadi_comm_init", "hadi_comm_finalize", "hadi_comm_world",
    "hadi_comm_rank", "hadi_price_batch_sharded", "hadi_jacobian_batch_sharded", "hadi_batch_update_model",
]


class Model(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("S0", "V0", "r_d", "r_f", "kappa", "eta", "sigma", "rho")]


class Point(C.Structure):
    _fields_ = [("strike", C.c_double), ("maturity", C.c_double), ("time_steps", C.c_int),
                ("delta_t", C.c_double), ("global_index", C.c_int)]


class Numerics(C.Structure):
    _fields_ = [("m1", C.c_int), ("m2", C.c_int), ("theta", C.c_double), ("style", C.c_int),
                ("payoff", C.c_int), ("scheme", C.c_int), ("num_dividends", C.c_int),
                ("dividend_dates", _dp), ("dividend_amounts", _dp), ("dividend_percentages", _dp),
                ("boundary", C.c_int), ("dividend_schedule", C.c_int)]


class LmOptions(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("tol", C.c_double), ("delta_tol", C.c_double),
                ("lambda0", C.c_double), ("eps", C.c_double)]


class LmResult(C.Structure):
    _fields_ = [("params", C.c_double * 5), ("final_error", C.c_double), ("lambda_", C.c_double),
                ("delta_norm", C.c_double), ("iterations", C.c_int), ("converged", C.c_int),
                ("pde_solves", C.c_int), ("gpu_ms", C.c_double), ("exact_reruns", C.c_longlong)]


class JacobianOptions(C.Structure):
    _fields_ = [("mode", C.c_int), ("eps", C.c_double * 5), ("schedule", C.c_int)]


LM_SCHEDULE_REFERENCE, LM_SCHEDULE_SPECULATIVE = 0, 1


def make_jacobian_options(mode=MODE_JACOBIAN, eps=1e-6, schedule=LM_SCHEDULE_REFERENCE):
    jo = JacobianOptions()
    jo.mode = mode
    jo.schedule = schedule
    e = np.broadcast_to(np.asarray(eps, dtype=np.float64), (5,))
    for k in range(5):
        jo.eps[k] = float(e[k])
    return jo


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp, C.c_int, _dp, _ip, _ip, C.c_int)


class Comm(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("allgather", ALLGATHER_FN), ("user", C.c_void_p)]


class HadiError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("hadi error %d %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libhadi.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HadiError(ERR_CUDA, "libhadi.so is not built (run __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        L.hadi_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.hadi_destroy.argtypes = [C.c_void_p]
        L.hadi_destroy.restype = None
        L.hadi_last_error.argtypes = [C.c_void_p]
        L.hadi_last_error.restype = C.c_char_p
        L.hadi_version.restype = C.c_char_p
        L.hadi_kernel_launches.argtypes = [C.c_void_p]
        L.hadi_kernel_launches.restype = C.c_longlong
        L.hadi_nccl_unique_id.argtypes = [C.c_void_p]
        L.hadi_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.hadi_comm_finalize.argtypes = [C.c_void_p]
        L.hadi_comm_finalize.restype = None
        L.hadi_comm_world.argtypes = [C.c_void_p]
        L.hadi_comm_rank.argtypes = [C.c_void_p]
        L.hadi_price_batch_sharded.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                               C.POINTER(Point), _dp]
        L.hadi_jacobian_batch_sharded.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                                  C.POINTER(Point), C.POINTER(JacobianOptions), _dp, _dp]
        L.hadi_batch_update_model.argtypes = [C.c_void_p, C.POINTER(Model)]
        L.hadi_convergence_study.argtypes = [C.c_void_p, C.POINTER(Model), C.c_double, C.c_double, C.c_int, C.c_double,
                                             C.c_int, C.c_int, C.POINTER(C.c_int), C.c_double, C.c_int, _dp, _dp, _dp]
        L.hadi_write_convergence_csv.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), _dp, _dp, _dp]
        L.hadi_exact_reruns.argtypes = [C.c_void_p]
        L.hadi_exact_reruns.restype = C.c_longlong
        L.hadi_batch_exact_reruns.argtypes = [C.c_void_p]
        L.hadi_batch_exact_reruns.restype = C.c_longlong
        L.hadi_price_batch.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                       C.POINTER(Point), _dp, _dp, _dp]
        L.hadi_jacobian_batch.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                          C.POINTER(Point), C.c_double, _dp, _dp]
        L.hadi_batch_create.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                        C.POINTER(Point), C.c_int, C.c_double, C.c_int, C.c_int,
                                        C.POINTER(C.c_void_p)]
        L.hadi_batch_num_items.argtypes = [C.c_void_p]
        L.hadi_batch_kernel_info.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.hadi_batch_launch.argtypes = [C.c_void_p]
        L.hadi_batch_values_dev.argtypes = [C.c_void_p]
        L.hadi_batch_values_dev.restype = C.c_void_p
        L.hadi_batch_fetch.argtypes = [C.c_void_p, _dp]
        L.hadi_batch_elapsed_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.hadi_batch_destroy.argtypes = [C.c_void_p]
        L.hadi_batch_destroy.restype = None
        L.hadi_jacobian_assemble.argtypes = [C.c_int, _dp, C.c_double, _dp, _dp]
        L.hadi_partition.argtypes = [C.c_int, _ip, C.c_int, C.c_int, _ip, _ip]
        L.hadi_item_costs.argtypes = [C.POINTER(Numerics), C.c_int, C.POINTER(Point), C.c_int, _ip]
        L.hadi_solve5.argtypes = [_dp, _dp, _dp]
        L.hadi_lm_update.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp]
        L.hadi_calibrate.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                     C.POINTER(Point), _dp, C.POINTER(LmOptions), C.POINTER(Comm),
                                     C.POINTER(LmResult)]
        L.hadi_grid.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.hadi_bs_call.argtypes = [C.c_double] * 5
        L.hadi_bs_call.restype = C.c_double
        L.hadi_transfer_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.hadi_measure_fp64.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3
        L.hadi_batch_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.hadi_batch_create_ex.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                           C.POINTER(Point), C.c_int, _dp, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.hadi_batch_values_per_item.argtypes = [C.c_void_p]
        L.hadi_jacobian_assemble_ex.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_double, _dp, _dp]
        L.hadi_jacobian_v0_weight.argtypes = [C.c_int, C.c_double, C.c_double, _ip, _ip, _dp]
        L.hadi_jacobian_batch_ex.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                             C.POINTER(Point), C.POINTER(JacobianOptions), _dp, _dp]
        L.hadi_calibrate_ex.argtypes = [C.c_void_p, C.POINTER(Model), C.POINTER(Numerics), C.c_int,
                                        C.POINTER(Point), _dp, C.POINTER(LmOptions), C.POINTER(JacobianOptions),
                                        C.POINTER(Comm), C.POINTER(LmResult)]
        L.hadi_plan_schedule.argtypes = [C.c_int, _ip, C.c_int, C.c_double, C.c_int, _ip, _ip, _dp]
        L.hadi_bs_vega.argtypes = [C.c_double] * 5
        L.hadi_bs_vega.restype = C.c_double
        L.hadi_bs_implied_vol.argtypes = [C.c_double] * 7
        L.hadi_bs_implied_vol.restype = C.c_double
        L.hadi_bs_implied_vol_bisect.argtypes = [C.c_double] * 8
        L.hadi_bs_implied_vol_bisect.restype = C.c_double
        L.hadi_dividend_adjusted_spot.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp, _dp]
        L.hadi_dividend_adjusted_spot.restype = C.c_double
        L.hadi_market_prices.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(Point), C.c_int,
                                         _dp, _dp, _dp, _dp]
        L.hadi_implied_vols.argtypes = [C.c_double, C.c_double, C.c_int, C.POINTER(Point), _dp, _dp, C.c_double,
                                        _dp, _dp, _dp]
        L.hadi_write_calibration_csv.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                                 C.POINTER(Point), _dp, _dp, C.POINTER(Model), C.POINTER(LmResult),
                                                 C.c_double, C.c_double]
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


NCCL_ID_BYTES = 128


def nccl_unique_id():
    """128-byte NCCL id (call on rank 0, hand to every rank, then Context.comm_init)."""
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    rc = lib().hadi_nccl_unique_id(buf)
    if rc != OK:
        raise HadiError(rc, "hadi_nccl_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw


def make_model(S0, V0, r_d, r_f, kappa, eta, sigma, rho):
    return Model(S0, V0, r_d, r_f, kappa, eta, sigma, rho)


def make_points(strikes, maturities, time_steps, delta_t=None):
    """CalibrationPoint array; delta_t defaults to maturity / N as the reference computes it."""
    strikes = np.atleast_1d(np.asarray(strikes, dtype=np.float64))
    n = strikes.size
    mats = np.broadcast_to(np.asarray(maturities, dtype=np.float64), (n,))
    Ns = np.broadcast_to(np.asarray(time_steps, dtype=np.int64), (n,))
    dts = mats / Ns if delta_t is None else np.broadcast_to(np.asarray(delta_t, dtype=np.float64), (n,))
    pts = (Point * max(n, 1))()
    for k in range(n):
        pts[k] = Point(float(strikes[k]), float(mats[k]), int(Ns[k]), float(dts[k]), k)
    return pts, n


class _NumKeep:
    """Numerics struct plus the numpy arrays its pointers refer to."""

    def __init__(self, m1, m2, theta, style=EUROPEAN, payoff=CALL, scheme=DOUGLAS, divs=None, boundary=0,
                 dividend_schedule=0):
        if divs is not None and len(divs[0]) > 0:
            self.arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in divs]
            nd = self.arrs[0].size
            self.num = Numerics(m1, m2, theta, style, payoff, scheme, nd, _d(self.arrs[0]), _d(self.arrs[1]),
                                _d(self.arrs[2]), boundary, dividend_schedule)
        else:
            self.arrs = []
            self.num = Numerics(m1, m2, theta, style, payoff, scheme, 0, None, None, None, boundary, dividend_schedule)


def make_numerics(m1, m2, theta, style=EUROPEAN, payoff=CALL, scheme=DOUGLAS, divs=None, boundary=0,
                  dividend_schedule=0):
    """boundary: BC_REFERENCE_CALL (parity path) | BC_PUT; dividend_schedule: DIVIDENDS_DEVICE (parity path) |
    DIVIDENDS_ALL — the two opt-in extensions of include/hadi.h (parity unpinned)."""
    return _NumKeep(m1, m2, theta, style, payoff, scheme, divs, boundary, dividend_schedule)


def grid(m1, m2, K, S0, V0):
    s, v = np.zeros(m1 + 1), np.zeros(m2 + 1)
    rc = lib().hadi_grid(m1, m2, K, S0, V0, _d(s), _d(v))
    if rc != OK:
        raise HadiError(rc)
    return s, v


def measure_fp64(device=0):
    """(un-fused DMUL+DADD TFLOP/s, DFMA TFLOP/s, dependent-DADD latency ns) measured on `device`."""
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    rc = lib().hadi_measure_fp64(device, C.byref(a), C.byref(b), C.byref(c))
    if rc != OK:
        raise HadiError(rc)
    return a.value, b.value, c.value


def bs_call(S, K, r, vol, T):
    return lib().hadi_bs_call(S, K, r, vol, T)


def bs_vega(S, K, r, vol, T):
    return lib().hadi_bs_vega(S, K, r, vol, T)


def bs_implied_vol(S, K, r, T, v0, target, eps):
    return lib().hadi_bs_implied_vol(S, K, r, T, v0, target, eps)


def bs_implied_vol_bisect(S, K, r, T, target, eps, a, b):
    return lib().hadi_bs_implied_vol_bisect(S, K, r, T, target, eps, a, b)


def _divs(divs):
    if divs is None or len(divs[0]) == 0:
        return 0, None, None, None, []
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in divs]
    return arrs[0].size, _d(arrs[0]), _d(arrs[1]), _d(arrs[2]), arrs


def dividend_adjusted_spot(S0, T, r_d, divs):
    nd, a, b, c, keep = _divs(divs)
    return lib().hadi_dividend_adjusted_spot(S0, T, r_d, nd, a, b, c)


def market_prices(S0, r_d, vol, pts, n, divs=None):
    nd, a, b, c, keep = _divs(divs)
    out = np.zeros(max(n, 1))
    rc = lib().hadi_market_prices(S0, r_d, vol, n, pts, nd, a, b, c, _d(out))
    if rc != OK:
        raise HadiError(rc)
    return out[:n]


def implied_vols(spot, r_d, pts, n, market, fitted, eps=0.01):
    market = np.ascontiguousarray(market, dtype=np.float64)
    fitted = np.ascontiguousarray(fitted, dtype=np.float64)
    miv, fiv, dif = np.zeros(max(n, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1))
    rc = lib().hadi_implied_vols(spot, r_d, n, pts, _d(market), _d(fitted), eps, _d(miv), _d(fiv), _d(dif))
    if rc != OK:
        raise HadiError(rc)
    return miv[:n], fiv[:n], dif[:n]


def write_calibration_csv(path, fmt, spot, r_d, n_maturities, n_strikes, pts, market, fitted, initial, result,
                          total_time_s, iv_eps=0.01):
    """`result` is the dict Context.calibrate returns."""
    market = np.ascontiguousarray(market, dtype=np.float64)
    fitted = np.ascontiguousarray(fitted, dtype=np.float64)
    res = LmResult()
    for k in range(5):
        res.params[k] = result["params"][k]
    res.final_error = result["final_error"]
    res.iterations = result["iterations"]
    res.pde_solves = result["pde_solves"]
    rc = lib().hadi_write_calibration_csv(os.fsencode(path), fmt, spot, r_d, n_maturities, n_strikes, pts,
                                          _d(market), _d(fitted), C.byref(initial), C.byref(res), total_time_s,
                                          iv_eps)
    if rc != OK:
        raise HadiError(rc)


def write_convergence_csv(path, m2_sizes, prices, errors, seconds):
    m2s = np.ascontiguousarray(m2_sizes, dtype=np.int32)
    rc = lib().hadi_write_convergence_csv(str(path).encode(), m2s.size, m2s.ctypes.data_as(C.POINTER(C.c_int)),
                                          _d(np.ascontiguousarray(prices, dtype=np.float64)),
                                          _d(np.ascontiguousarray(errors, dtype=np.float64)),
                                          _d(np.ascontiguousarray(seconds, dtype=np.float64)))
    if rc != OK:
        raise HadiError(rc, "hadi_write_convergence_csv")


def solve5(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(5)
    rc = lib().hadi_solve5(_d(A), _d(b), _d(x))
    if rc != OK:
        raise HadiError(rc)
    return x


def lm_update(J, r, lam):
    J = np.ascontiguousarray(J, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64)
    delta = np.zeros(5)
    rc = lib().hadi_lm_update(r.size, _d(J), _d(r), lam, _d(delta))
    if rc != OK:
        raise HadiError(rc)
    return delta


def partition(costs, world, rank):
    costs = np.ascontiguousarray(costs, dtype=np.int32)
    b, e = C.c_int(0), C.c_int(0)
    rc = lib().hadi_partition(costs.size, costs.ctypes.data_as(_ip), world, rank, C.byref(b), C.byref(e))
    if rc != OK:
        raise HadiError(rc)
    return b.value, e.value


def item_costs(num, pts, n, mode):
    nc = ITEMS_PER_OPTION[mode]
    costs = np.zeros(max(n * nc, 1), dtype=np.int32)
    rc = lib().hadi_item_costs(C.byref(num.num), n, pts, mode, costs.ctypes.data_as(_ip))
    if rc != OK:
        raise HadiError(rc)
    return costs[:n * nc]


def jacobian_assemble(values, eps):
    values = np.ascontiguousarray(values, dtype=np.float64)
    n = values.size // 6
    J, base = np.zeros((n, 5)), np.zeros(n)
    rc = lib().hadi_jacobian_assemble(n, _d(values), eps, _d(J), _d(base))
    if rc != OK:
        raise HadiError(rc)
    return J, base


def plan_schedule(time_steps, slots, setup=1.5):
    """(segments [(item, n0, n1, hin, hout)], slot_off, heaviest load in steps) of the split schedule; ([], None, 0)
    when nothing is cut."""
    ts = np.ascontiguousarray(time_steps, dtype=np.int32)
    cap = 2 * ts.size + slots + 8
    seg = np.zeros(5 * cap, dtype=np.int32)
    off = np.zeros(slots + 1, dtype=np.int32)
    hv = C.c_double(0.0)
    rc = lib().hadi_plan_schedule(ts.size, ts.ctypes.data_as(_ip), slots, setup, cap, seg.ctypes.data_as(_ip),
                                  off.ctypes.data_as(_ip), C.byref(hv))
    if rc < 0:
        raise HadiError(rc)
    if rc == 0:
        return [], None, 0.0
    return [tuple(int(x) for x in seg[5 * q:5 * q + 5]) for q in range(rc)], off, hv.value


def jacobian_v0_weight(m2, V0, eps_v0):
    lo, hi, w = C.c_int(0), C.c_int(0), C.c_double(0.0)
    rc = lib().hadi_jacobian_v0_weight(m2, V0, eps_v0, C.byref(lo), C.byref(hi), C.byref(w))
    if rc != OK:
        raise HadiError(rc)
    return lo.value, hi.value, w.value


def jacobian_assemble_ex(values, mode, eps5, v0_weight=0.0):
    values = np.ascontiguousarray(values, dtype=np.float64)
    eps5 = np.ascontiguousarray(np.broadcast_to(np.asarray(eps5, dtype=np.float64), (5,)))
    n = values.size // (ITEMS_PER_OPTION[mode] * VALUES_PER_ITEM[mode])
    J, base = np.zeros((n, 5)), np.zeros(n)
    rc = lib().hadi_jacobian_assemble_ex(n, mode, _d(values), _d(eps5), v0_weight, _d(J), _d(base))
    if rc != OK:
        raise HadiError(rc)
    return J, base


class Context:
    """hadi_ctx: one CUDA device, one stream."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().hadi_create(C.byref(self._h), device)
        if rc != OK:
            self._h = None
            raise HadiError(rc, "hadi_create failed (no usable CUDA device?)")

    def close(self):
        if self._h:
            # prepared batches hold pool buffers of this context: release them first (a Batch destroyed after its
            # Context would hand the library a dangling pointer)
            for b in list(getattr(self, "_batches", ())):
                b.destroy()
            lib().hadi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise HadiError(rc, lib().hadi_last_error(self._h).decode())

    @property
    def kernel_launches(self):
        return lib().hadi_kernel_launches(self._h)

    # ---- in-library exchange (NCCL communicator owned by the context) ---------------------------------------
    def comm_init(self, world, rank, id_bytes):
        buf = C.create_string_buffer(bytes(id_bytes), NCCL_ID_BYTES)
        self._check(lib().hadi_comm_init(self._h, world, rank, buf))

    def comm_finalize(self):
        lib().hadi_comm_finalize(self._h)

    @property
    def comm_world(self):
        return lib().hadi_comm_world(self._h)

    def price_batch_sharded(self, model, num, pts, n):
        prices = np.zeros(max(n, 1))
        self._check(lib().hadi_price_batch_sharded(self._h, C.byref(model), C.byref(num.num), n, pts, _d(prices)))
        return prices[:n]

    def jacobian_batch_sharded(self, model, num, pts, n, mode=MODE_JACOBIAN, eps=1e-6):
        jo = make_jacobian_options(mode, eps)
        J, base = np.zeros((max(n, 1), 5)), np.zeros(max(n, 1))
        self._check(lib().hadi_jacobian_batch_sharded(self._h, C.byref(model), C.byref(num.num), n, pts, C.byref(jo),
                                                      _d(J), _d(base)))
        return J[:n], base[:n]

    def convergence_study(self, model, K, T, N, theta, scheme, m2_sizes, ref_price, repeats=20):
        """ConvergenceExporter::testWithRelatedGridSizes on the GPU: (prices, relative errors, mean seconds)."""
        m2s = np.ascontiguousarray(m2_sizes, dtype=np.int32)
        n = m2s.size
        p, e, t = np.zeros(max(n, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1))
        self._check(lib().hadi_convergence_study(self._h, C.byref(model), K, T, N, theta, scheme, n,
                                                 m2s.ctypes.data_as(C.POINTER(C.c_int)), ref_price, repeats, _d(p), _d(e),
                                                 _d(t)))
        return p[:n], e[:n], t[:n]

    @property
    def exact_reruns(self):
        """Solves this context repeated with IEEE divisions (guarded division out of range / hand-off time-out)."""
        return lib().hadi_exact_reruns(self._h)

    def transfer_bytes(self):
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._check(lib().hadi_transfer_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def price_batch(self, model, num, pts, n, want_U=False, want_lambda=False):
        P = (num.num.m1 + 1) * (num.num.m2 + 1)
        prices = np.zeros(max(n, 1))
        U = np.zeros((max(n, 1), P)) if want_U else None
        lam = np.zeros((max(n, 1), P)) if want_lambda else None
        self._check(lib().hadi_price_batch(self._h, C.byref(model), C.byref(num.num), n, pts, _d(prices), _d(U),
                                           _d(lam)))
        out = {"prices": prices[:n]}
        if want_U:
            out["U"] = U[:n]
        if want_lambda:
            out["lambda"] = lam[:n]
        return out

    def jacobian_batch(self, model, num, pts, n, eps=1e-6):
        J, base = np.zeros((max(n, 1), 5)), np.zeros(max(n, 1))
        self._check(lib().hadi_jacobian_batch(self._h, C.byref(model), C.byref(num.num), n, pts, eps, _d(J),
                                              _d(base)))
        return J[:n], base[:n]

    def jacobian_batch_ex(self, model, num, pts, n, mode, eps=1e-6):
        """Jacobian taken as `mode` says (forward / interpolated V0 column / central), eps scalar or [5]."""
        jo = make_jacobian_options(mode, eps)
        J, base = np.zeros((max(n, 1), 5)), np.zeros(max(n, 1))
        self._check(lib().hadi_jacobian_batch_ex(self._h, C.byref(model), C.byref(num.num), n, pts, C.byref(jo),
                                                 _d(J), _d(base)))
        return J[:n], base[:n]

    def batch(self, model, num, pts, n, mode=MODE_PRICE, eps=1e-6, begin=0, end=-1):
        return Batch(self, model, num, pts, n, mode, eps, begin, end)

    def calibrate(self, model, num, pts, n, market, max_iter, tol, delta_tol, lambda0=0.01, eps=1e-6, comm=None,
                  jac_mode=None, schedule=LM_SCHEDULE_REFERENCE):
        market = np.ascontiguousarray(market, dtype=np.float64)
        opt = LmOptions(max_iter, tol, delta_tol, lambda0, float(np.atleast_1d(eps)[0]))
        res = LmResult()
        if jac_mode is None and schedule == LM_SCHEDULE_REFERENCE:
            self._check(lib().hadi_calibrate(self._h, C.byref(model), C.byref(num.num), n, pts, _d(market),
                                             C.byref(opt), None if comm is None else C.byref(comm), C.byref(res)))
        else:
            jo = make_jacobian_options(MODE_JACOBIAN if jac_mode is None else jac_mode, eps, schedule)
            self._check(lib().hadi_calibrate_ex(self._h, C.byref(model), C.byref(num.num), n, pts, _d(market),
                                                C.byref(opt), C.byref(jo),
                                                None if comm is None else C.byref(comm), C.byref(res)))
        return dict(params=list(res.params), final_error=res.final_error, lam=res.lambda_,
                    delta_norm=res.delta_norm, iterations=res.iterations, converged=res.converged,
                    pde_solves=res.pde_solves, gpu_ms=res.gpu_ms, exact_reruns=res.exact_reruns)


class Batch:
    """hadi_batch: descriptors and grids resident in HBM; launch() is asynchronous."""

    def __init__(self, ctx, model, num, pts, n, mode, eps, begin, end):
        self.ctx = ctx
        self._keep = (model, num, pts)
        self._h = C.c_void_p()
        eps5 = np.ascontiguousarray(np.broadcast_to(np.asarray(eps, dtype=np.float64), (5,)))
        ctx._check(lib().hadi_batch_create_ex(ctx._h, C.byref(model), C.byref(num.num), n, pts, mode, _d(eps5),
                                              begin, end, C.byref(self._h)))
        self.n_items = lib().hadi_batch_num_items(self._h)
        self.values_per_item = lib().hadi_batch_values_per_item(self._h)
        if not hasattr(ctx, "_batches"):
            ctx._batches = weakref.WeakSet()
        ctx._batches.add(self)

    def update_model(self, model):
        """Re-aim the prepared batch at new (kappa, eta, sigma, rho, V0); S0, r_d, r_f as at creation."""
        self.ctx._check(lib().hadi_batch_update_model(self._h, C.byref(model)))

    def launch(self):
        self.ctx._check(lib().hadi_batch_launch(self._h))

    def fetch(self):
        nv = self.n_items * self.values_per_item
        vals = np.zeros(max(nv, 1))
        self.ctx._check(lib().hadi_batch_fetch(self._h, _d(vals)))
        return vals[:nv]

    @property
    def exact_reruns(self):
        return lib().hadi_batch_exact_reruns(self._h)

    @property
    def kernel_info(self):
        """(variant id, CTAs of the launch, CTAs per solve) the library planned for this batch."""
        v, g, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self.ctx._check(lib().hadi_batch_kernel_info(self._h, C.byref(v), C.byref(g), C.byref(c)))
        return v.value, g.value, c.value

    def elapsed_ms(self):
        ms = C.c_float(0.0)
        self.ctx._check(lib().hadi_batch_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def values_dev(self):
        return lib().hadi_batch_values_dev(self._h)

    def destroy(self):
        if self._h:
            lib().hadi_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
