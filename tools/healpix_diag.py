import math, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
s2g = ge.load_package()
from oracle import oracle as orc
nside = 128
rng = np.random.default_rng(5)
for n in (1200, 24000):
    ang = math.sqrt(4 * math.pi / (12 * nside * nside))
    pos = rng.normal(size=(n, 3)) * 60.0
    dist = np.linalg.norm(pos, axis=1)
    hsml = dist * np.sin(ang * rng.uniform(3.0, 12.0, n))
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 1e4; w = rng.random(n) + 0.5
    a, wm = s2g.healpix_deposit(pos, hsml, m, rho, q, w, nside, s2g.WendlandC4(2), True)
    ra, rw, _ = orc.healpix_deposit(pos, hsml, m, rho, q, w, nside, "WendlandC4", 2, True)
    rel = np.abs(wm - rw) / np.maximum(np.abs(rw), 1e-300)
    rel[rw == 0] = 0
    k = np.argsort(rel)[-5:]
    print(n, "max rel", rel.max(), "quantiles", np.quantile(rel[rw > 0], [0.5, 0.99, 0.9999]))
    print("  worst pixels value/max:", rw[k] / rw.max(), rel[k])
