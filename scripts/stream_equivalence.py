"""CPU: the identity the streaming kernels rest on, demonstrated on the oracle's C restatement.
`oracle_set_fresh_statistics(1)` makes the restatement recompute S_in / Q_in after every action from the active set alone
(beta of the last FullStat, ghost term of deleted bases) instead of correcting them incrementally as the reference does
(oracle/eben_gauss.c).  Prints, per design, the largest relative difference in intercept / residual variance / weights
between the two and the number of fits whose support changed.  usage: python scripts/stream_equivalence.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import rlayer as R

lib = R.fit_lib("port")
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "inputs_bundled.npz"))
X, y = g["BASIS"].astype(np.float64), g["y"].astype(np.float64)
rng = np.random.default_rng(0)
for epis, K, N in ((True, 60, 400), (False, 481, 900), (True, 120, 900)):
    cols = np.sort(rng.choice(481, K, replace=False)) if K < 481 else np.arange(481)
    Xs, ys = X[:N, cols], y[:N]
    worst, flips, n = 0.0, 0, 0
    for lam in (2.0, 0.5, 0.1, 0.03, 0.01):
        for al in (1.0, 0.5, 0.05):
            lib.lib.oracle_set_fresh_statistics(0); a = R.eb_elastic_net_gaussian(Xs, ys, lam, al, epis, lib)
            lib.lib.oracle_set_fresh_statistics(1); b = R.eb_elastic_net_gaussian(Xs, ys, lam, al, epis, lib)
            lib.lib.oracle_set_fresh_statistics(0)
            n += 1
            ra, rb = a.raw_beta[:, 2], b.raw_beta[:, 2]
            if not np.array_equal(ra != 0, rb != 0):
                flips += 1
                continue
            d = max(abs(a.intercept[0] - b.intercept[0]) / abs(a.intercept[0]), abs(a.resid_var - b.resid_var) / a.resid_var,
                    float(np.max(np.abs(ra - rb))) / max(float(np.max(np.abs(ra))), 1e-300))
            worst = max(worst, d)
    print(f"Epis={epis} N={N} K={K}: {n} fits, support changes {flips}, largest relative difference {worst:.2e}")
