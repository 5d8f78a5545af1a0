"""Development aid: the 51x26 kernel (variant 1) at different CTA sizes: the reference's benchmark protocol (European, strike
85, N = 20) and the Jacobian batch of the 10x10 LM surface.  usage: time_51x26.py lib.so ..."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.path.join(ROOT, "pde-based-heston-solver-gpu-accelerated_b200")
def run(libname):
    import importlib.util
    spec = importlib.util.spec_from_file_location("hadi", os.path.join(pkg, "hadi.py")); hadi = importlib.util.module_from_spec(spec); spec.loader.exec_module(hadi)
    hadi.LIB_PATH = os.path.join(pkg, libname)
    ctx = hadi.Context(0)
    mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
    divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
    out = []
    for (n, N, style, dv) in ((1, 20, 0, None), (100, 20, 0, None), (500, 20, 0, None), (2000, 20, 0, None), (2000, 50, 1, divs)):
        num = hadi.make_numerics(50, 25, 0.8, style, 0, 0, dv)
        pts, n = hadi.make_points([85.0 + 0.01 * k for k in range(n)], 1.0, N)
        bt = ctx.batch(mdl, num, pts, n)
        ts = []
        for r in range(5):
            bt.launch(); v = bt.fetch().copy(); ts.append(bt.elapsed_ms())
        out.append("n=%d N=%d style %d: %.4f ms (sum %.10f)" % (n, N, style, min(ts), float(v.sum())))
        bt.destroy()
    print(libname, bt.kernel_info if False else "", "; ".join(out), flush=True)
if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        run(sys.argv[2]); sys.exit(0)
    for lib in ["libhadi.so"] + sys.argv[1:]:
        subprocess.run([sys.executable, __file__, "--one", lib], check=False)
