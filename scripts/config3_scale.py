"""GPU box: BASELINE config 3 at full size on a synthetic stand-in (the yeast file itself is not in the checkout):
N = 3803 rows x K = 28220 +-1 markers with linkage, 10 folds, Gaussian main effects.
usage: python scripts/config3_scale.py [grid_step]    (grid_step 1 = all 4,000 fits)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb

step = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rng = np.random.default_rng(20260103)
n, k, nf = 3803, 28220, 10
t = time.time()
X = np.empty((n, k), dtype=np.int8, order="F")
col = rng.choice(np.array([-1, 1], np.int8), n)
for j in range(k):
    if j % 40 == 0: col = rng.choice(np.array([-1, 1], np.int8), n)
    else:
        flip = rng.random(n) < 0.03
        col = np.where(flip, -col, col).astype(np.int8)
    X[:, j] = col
Xf = X.astype(np.float64, order="F")
beta = np.zeros(k); idx = rng.choice(k, 25, replace=False); beta[idx] = rng.normal(0, 0.5, 25)
y = 10 + Xf @ beta + rng.normal(0, 1.0, n)
print(f"generated {n} x {k} in {time.time()-t:.1f} s", flush=True)
folds = pb.AssignToFolds(Xf, nf)
t = time.time(); grid = pb.BuildGrid(Xf, y, nf); print(f"BuildGrid (lambda_max on device) {time.time()-t:.2f} s, 10*lambda_max = {grid['lambda'].max():.4f}", flush=True)
rows = np.arange(0, 400, step)
fold = np.tile(np.arange(1, nf + 1), rows.size); a = np.repeat(grid["alpha"][rows], nf); l = np.repeat(grid["lambda"][rows], nf)
t = time.time()
with pb.Problem(Xf, y, folds, nf, False, "gaussian") as p:
    print(f"problem_create (H2D + per-fold layouts) {time.time()-t:.2f} s", flush=True)
    t = time.time(); err, st, ns, it = p.run_fits(fold, a, l); dt = time.time() - t
    fl, ms, _ = p.counters()
print(f"config-3 scale: {fold.size} fits in {dt:.2f} s ({fold.size/dt:.1f} fits/s), kernel {ms:.0f} ms, alg {fl/1e12:.1f} TFLOP -> {fl/ms/1e9:.2f} TFLOP/s, "
      f"max active set {ns.max()}, status counts {dict(zip(*np.unique(st, return_counts=True)))}, finite {np.isfinite(err).all()}")
np.savez_compressed("gpurun_out/config3_scale.npz", err=err, st=st, ns=ns, it=it, lam=l, alpha=a, fold=fold)
