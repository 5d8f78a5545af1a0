"""Development aid: race check + kernel time of config 2 for a given library."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
hadi = ge.load_hadi()
libname = sys.argv[1]
hadi.LIB_PATH = os.path.join(os.path.dirname(hadi.LIB_PATH), libname)
BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02] * 4)
ctx = hadi.Context(0)
mdl = hadi.make_model(**BASE)
num = hadi.make_numerics(100, 50, 0.8, 1, 0, 0, DIVS)
K = [70 + 0.12 * i for i in range(500)]
pts, nn = hadi.make_points(K, 1.0, 50)
bt = ctx.batch(mdl, num, pts, nn)
ts = []
for r in range(12):
    bt.launch(); g = bt.fetch(); ts.append(bt.elapsed_ms())
print(libname, "config2 kernel ms: min %.4f median %.4f" % (min(ts[2:]), float(np.median(ts[2:]))))
num_e = hadi.make_numerics(100, 50, 0.8, 0, 0, 0, None)
bt2 = ctx.batch(mdl, num_e, pts, nn)
ts = []
for r in range(12):
    bt2.launch(); bt2.fetch(); ts.append(bt2.elapsed_ms())
print(libname, "EU 500x50 kernel ms: min %.4f median %.4f" % (min(ts[2:]), float(np.median(ts[2:]))))
