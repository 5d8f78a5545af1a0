"""Development aid: A/B experiment builds (make exp NAME=...) against libhadi.so on config 2.
usage: python tools/ab_libs.py libhadi_x.so [libhadi_y.so ...]   (run on the GPU box; each library in its own process)"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200")


def run(libname):
    spec = importlib.util.spec_from_file_location("hadi", os.path.join(pkg, "hadi.py"))
    hadi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hadi)
    hadi.LIB_PATH = os.path.join(pkg, libname)
    L = hadi.lib()
    ctx = hadi.Context(0)
    mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
    divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
    names = ["setup+div", "a1fwd", "explicit", "a1", "a2", "project", "ringwait", "rhs2"]
    out = {}
    for tag, (n, N, m1, m2, style, dv) in dict(c2=(500, 50, 100, 50, 1, divs), eu=(592, 50, 100, 50, 0, None),
                                               one=(148, 50, 100, 50, 1, divs)).items():
        num = hadi.make_numerics(m1, m2, 0.8, style, 0, 0, dv)
        pts, n = hadi.make_points([70 + 0.12 * i for i in range(n)], 1.0, N)
        bt = ctx.batch(mdl, num, pts, n)
        best = 1e9
        for r in range(6):
            bt.launch()
            vals = bt.fetch().copy()
            best = min(best, bt.elapsed_ms())
        cyc = (C.c_longlong * 8)()
        L.hadi_batch_phase_cycles(bt._h, cyc)
        per = {names[k]: round(cyc[k] / (n * N)) for k in range(8) if cyc[k]}
        print(f"{libname:24s} {tag}: best {best:.3f} ms  phases {per}", flush=True)
        np.save(os.path.join(ROOT, "gpurun_out", f"ab_{libname}_{tag}.npy"), vals)
        bt.destroy()


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        run(sys.argv[2])
        sys.exit(0)
    libs = ["libhadi.so"] + sys.argv[1:]
    for lib in libs:
        subprocess.run([sys.executable, __file__, "--one", lib], check=False)
    for lib in libs[1:]:
        for tag in ("c2", "eu", "one"):
            a = np.load(os.path.join(ROOT, "gpurun_out", f"ab_libhadi.so_{tag}.npy"))
            b = np.load(os.path.join(ROOT, "gpurun_out", f"ab_{lib}_{tag}.npy"))
            print(lib, tag, "bit-equal to libhadi.so:", bool(np.array_equal(a, b)))
