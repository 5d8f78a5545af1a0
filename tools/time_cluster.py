import os, sys
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
for var in ("6", "5", "7"):
    os.environ["HADI_FORCE_VARIANT"] = var
    for scheme, name in ((hadi.CRAIG_SNEYD, "CS"), (hadi.DOUGLAS, "DO")):
        for N in (20, 200):
            num = hadi.make_numerics(400, 200, 0.8, hadi.EUROPEAN, hadi.CALL, scheme, None)
            pts, n = hadi.make_points([100.0], 1.0, N)
            bt = ctx.batch(mdl, num, pts, n)
            ts = []
            for r in range(3):
                bt.launch(); v = bt.fetch(); ts.append(bt.elapsed_ms())
            print("variant", var, name, "N", N, "ms", min(ts), "per step us", 1e3 * min(ts) / N, "price", v[0])
            bt.destroy()
print("exact reruns:", getattr(ctx, "exact_reruns", None) if not callable(getattr(ctx, "exact_reruns", None)) else ctx.exact_reruns())
