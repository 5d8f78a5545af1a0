"""Development aid: per-stage cycle counters of the wide kernel (libhadi_timing.so, -DHADI_PHASE_TIMING)."""
import ctypes as C, importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200")
spec = importlib.util.spec_from_file_location("hadi", os.path.join(pkg, "hadi.py")); hadi = importlib.util.module_from_spec(spec); spec.loader.exec_module(hadi)
hadi.LIB_PATH = os.path.join(pkg, os.environ.get("HADI_LIB", "libhadi_timing.so"))
os.environ["HADI_FORCE_VARIANT"] = "9"
L = hadi.lib()
L.hadi_batch_prof_raw.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
ctx = hadi.Context(0)
mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
names = ["setup", "barrier", "rhs1", "a1chain", "a2chain", "rhs2", "store", "factors"]
cases = [(1, 50, 400, 200, 1), (1, 50, 400, 200, 0), (8, 50, 400, 200, 1), (1, 50, 100, 50, 0)]
if len(sys.argv) > 1:   # n,N,m1,m2,scheme ...
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for (n, N, m1, m2, scheme) in cases:
    num = hadi.make_numerics(m1, m2, 0.8, 0, 0, scheme, None)
    pts, n = hadi.make_points([100.0 + k for k in range(n)], 1.0, N)
    bt = ctx.batch(mdl, num, pts, n)
    for r in range(2):
        bt.launch(); bt.fetch()
    ms = bt.elapsed_ms()
    raw = (C.c_longlong * (8 * 148))()
    nc = L.hadi_batch_prof_raw(bt._h, raw, 148)
    a = np.array(raw[:8 * nc], dtype=np.int64).reshape(nc, 8) / N / 2   # two launches accumulate
    print("%dx%d scheme %d n=%d: %.3f ms, %.1f us/step; cycles per step (mean over %d CTAs / max):" % (m1 + 1, m2 + 1, scheme, n, ms, 1e3 * ms / N, nc))
    print("   ", {names[k]: (int(a[:, k].mean()), int(a[:, k].max())) for k in range(8)}, "sum", int(a.sum(axis=1).mean()))
    bt.destroy()
