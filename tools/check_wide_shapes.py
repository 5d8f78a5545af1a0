"""Development aid: the wide kernel against the default kernels on odd grid shapes (bit-equality of prices and grids)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.01, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
divs = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
bad = 0
for (m1, m2) in ((8, 4), (12, 12), (37, 36), (63, 31), (64, 32), (129, 65), (255, 127), (256, 128), (511, 255), (600, 100), (1000, 60), (600, 510), (760, 400), (1023, 200), (300, 299)):
    for scheme, style, dv in ((0, 1, divs), (1, 0, None), (3, 0, None)):
        for nopt in (1, 3, 11):
            num = hadi.make_numerics(m1, m2, 0.8, style, hadi.CALL, scheme, dv)
            pts, n = hadi.make_points([95.0 + 1.5 * k for k in range(nopt)], 1.0, 5)
            os.environ.pop("HADI_FORCE_VARIANT", None)
            os.environ["HADI_WIDE_MAX_ITEMS"] = "0"
            try:
                a = ctx.price_batch(mdl, num, pts, n, want_U=True)
            except hadi.HadiError as e:
                print(m1, m2, scheme, nopt, "default refuses:", e); continue
            os.environ.pop("HADI_WIDE_MAX_ITEMS")
            os.environ["HADI_FORCE_VARIANT"] = "9"
            try:
                b = ctx.price_batch(mdl, num, pts, n, want_U=True)
            except hadi.HadiError as e:
                print(m1, m2, scheme, nopt, "wide refuses:", str(e)[:80]); continue
            ok = np.array_equal(a["prices"], b["prices"]) and np.array_equal(a["U"], b["U"])
            bad += not ok
            print(m1, m2, "scheme", scheme, "style", style, "n", nopt, "equal", ok, flush=True)
print("mismatches:", bad)
