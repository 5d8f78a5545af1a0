"""Development aid: per-phase cycle breakdown (libhadi_timing.so) and the FP64 micro-benchmark."""
import ctypes as C
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200")
spec = importlib.util.spec_from_file_location("hadi", os.path.join(pkg, "hadi.py"))
hadi = importlib.util.module_from_spec(spec)
spec.loader.exec_module(hadi)
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
if libs:
    hadi.LIB_PATH = os.path.join(pkg, libs[0])
elif os.path.exists(os.path.join(pkg, "libhadi_timing.so")) and "--notiming" not in sys.argv:
    hadi.LIB_PATH = os.path.join(pkg, "libhadi_timing.so")
print("library:", os.path.basename(hadi.LIB_PATH))
L = hadi.lib()
L.hadi_measure_fp64.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3
L.hadi_batch_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
a, b, c = C.c_double(), C.c_double(), C.c_double()
print("fp64 rc", L.hadi_measure_fp64(0, C.byref(a), C.byref(b), C.byref(c)), "unfused TF/s %.2f fma TF/s %.2f dep-DADD ns %.3f" % (a.value, b.value, c.value))
ctx = hadi.Context(0)
mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
names = ["setup+div", "a1fwd", "explicit", "a1", "a2", "project", "ringwait", "rhs2"]
for (n, N, m1, m2, style, dv) in [(296, 50, 100, 50, 1, divs), (500, 50, 100, 50, 1, divs), (148, 50, 100, 50, 1, divs), (296, 50, 100, 50, 0, None), (592, 20, 50, 25, 0, None)]:
    num = hadi.make_numerics(m1, m2, 0.8, style, 0, 0, dv)
    strikes = [70 + 60.0 * i / n for i in range(n)]
    pts, n = hadi.make_points(strikes, 1.0, N)
    bt = ctx.batch(mdl, num, pts, n)
    for r in range(3):
        bt.launch(); bt.fetch()
    ms = bt.elapsed_ms()
    cyc = (C.c_longlong * 8)()
    L.hadi_batch_phase_cycles(bt._h, cyc)
    tot = sum(cyc)
    per = {names[k]: round(cyc[k] / (n * N)) for k in range(8)}
    flops = n * N * (m1 + 1) * (m2 + 1) * (72 if style else 64)
    print(f"n={n} N={N} {m1}x{m2} style={style} div={dv is not None}: {ms:.3f} ms, {flops/ms/1e9:.2f} TFLOP/s alg; cycles per item-step: {per}")
    bt.destroy()
