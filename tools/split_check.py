"""Development aid: split schedule on/off — timings and bit-equality (run on the GPU box)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(tag):
    import __graft_entry__ as ge
    hadi = ge.load_hadi()
    ctx = hadi.Context(0)
    mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
    divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
    cases = dict(c2=(500, 50, 100, 50, 1, divs), eu700=(700, 20, 100, 50, 0, None), s26=(1000, 20, 50, 25, 1, divs),
                 one=(296, 50, 100, 50, 1, divs), rt=(400, 12, 64, 32, 1, divs))
    for name, (n, N, m1, m2, style, dv) in cases.items():
        num = hadi.make_numerics(m1, m2, 0.8, style, 0, 0, dv)
        Ns = [N + (i % 3) * 5 for i in range(n)] if name == "s26" else N
        pts, n = hadi.make_points([70 + 60.0 * i / n for i in range(n)], 1.0, Ns)
        bt = ctx.batch(mdl, num, pts, n)
        best = 1e9
        for r in range(5):
            bt.launch()
            vals = bt.fetch().copy()
            best = min(best, bt.elapsed_ms())
        print(f"{tag:8s} {name:6s} n={n:5d}: best {best:.3f} ms", flush=True)
        np.save(os.path.join(ROOT, "gpurun_out", f"split_{tag}_{name}.npy"), vals)
        bt.destroy()


if __name__ == "__main__":
    if len(sys.argv) == 2:
        run(sys.argv[1])
        sys.exit(0)
    for tag, env in (("split", {}), ("nosplit", {"HADI_NO_SPLIT": "1"}), ("forcebad", {"HADI_DEBUG_STOP": "-7:0"})):
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, __file__, tag], env=e, check=False)
    for name in ("c2", "eu700", "s26", "one", "rt"):
        a = np.load(os.path.join(ROOT, "gpurun_out", f"split_nosplit_{name}.npy"))
        for tag in ("split", "forcebad"):
            b = np.load(os.path.join(ROOT, "gpurun_out", f"split_{tag}_{name}.npy"))
            print(name, tag, "bit-equal to nosplit:", bool(np.array_equal(a, b)), "max abs diff", float(np.max(np.abs(a - b))))
