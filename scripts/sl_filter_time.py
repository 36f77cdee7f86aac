"""GPU box: timing of the single-locus prefilter (pareben_sl_filter) on bundled and synthetic genotypes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb
from oracle import rlayer as R
g = np.load("tests/golden/inputs_bundled.npz")
X, y = g["BASIS"].astype(float), g["y"]
for rep in range(2):
    t = time.time(); out = pb.SLFilter(X, y, 0.02, 0.05, "yes"); dt = time.time() - t
print(f"bundled 1000 x 481, {481*480//2} pairs: {dt*1e3:.1f} ms end to end (H2D + kernels + D2H), kept {out['main'].size} main / {out['pairs'].shape[0]} pairs")
t = time.time(); want = R.sl_filter(X, y, 0.02, 0.05, True); dtc = time.time() - t
print(f"  numpy restatement of the R script on one core: {dtc:.1f} s; same kept sets: {np.array_equal(out['main'], want[0]) and np.array_equal(out['pairs'], want[1])}")
rng = np.random.default_rng(5)
n, k = 1000, 3000
Xs = rng.choice([-1.0, 0.0, 1.0], size=(n, k), p=[0.25, 0.5, 0.25]); ys = 2 * Xs[:, 10] * Xs[:, 900] + Xs[:, 5] + rng.normal(0, 2, n)
t = time.time(); out = pb.SLFilter(Xs, ys, 0.1, 0.12, "yes"); dt = time.time() - t
print(f"synthetic 1000 x 3000, {k*(k-1)//2} pairs: {dt:.2f} s, kept {out['main'].size} main / {out['pairs'].shape[0]} pairs; planted pair found: {[11, 901] in out['pairs'].tolist()}")
