"""Development aid: sweep HADI_SPLIT_SETUP on config 2 (run on the GPU box)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import __graft_entry__ as ge
    hadi = ge.load_hadi()
    ctx = hadi.Context(0)
    mdl = hadi.make_model(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
    divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
    out = []
    for name, (n, N, style, dv) in dict(c2=(500, 50, 1, divs), few=(330, 50, 1, divs), eu=(1000, 50, 0, None)).items():
        num = hadi.make_numerics(100, 50, 0.8, style, 0, 0, dv)
        pts, n = hadi.make_points([70 + 60.0 * i / n for i in range(n)], 1.0, N)
        bt = ctx.batch(mdl, num, pts, n)
        ts = []
        for r in range(8):
            bt.launch()
            bt.fetch()
            ts.append(bt.elapsed_ms())
        out.append(f"{name} best {min(ts):.3f} med {sorted(ts)[4]:.3f}")
        bt.destroy()
    print(os.environ.get("HADI_SPLIT_SETUP", "default"), os.environ.get("HADI_NO_SPLIT", ""), " | ".join(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 2:
        run()
        sys.exit(0)
    for v in ("nosplit", "0.5", "1.0", "1.5", "2.0", "3.0", "4.0"):
        e = dict(os.environ)
        if v == "nosplit":
            e["HADI_NO_SPLIT"] = "1"
        else:
            e["HADI_SPLIT_SETUP"] = v
        subprocess.run([sys.executable, __file__, "x"], env=e, check=False)
