"""GPU box: Epis="yes" cross-validation at config-4 scale (K loci -> K(K+1)/2 candidates generated on the fly),
bundled BASIS/y rows, Gaussian; prints kernel time and fits/s.  usage: python scripts/epis_scale.py K n_folds grid_step"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb
g = np.load("tests/golden/inputs_bundled.npz")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 10
step = int(sys.argv[3]) if len(sys.argv) > 3 else 8
X, y = g["BASIS"][:, :K].astype(float), g["y"]
folds = pb.AssignToFolds(X, nf)
t = time.time(); grid_e = pb.BuildGrid(X, y, nf, "yes"); t_grid = time.time() - t
grid = pb.BuildGrid(X, y, nf, "no")      # the Epis lambda_max (un-normalised response, R/BuildGrid.R:26) selects nothing on this data: use the main-effect grid
rows = np.arange(0, 400, step)
fold = np.tile(np.arange(1, nf + 1), rows.size); a = np.repeat(grid["alpha"][rows], nf); l = np.repeat(grid["lambda"][rows], nf)
with pb.Problem(X, y, folds, nf, True, "gaussian") as p:
    t = time.time(); err, st, ns, it = p.run_fits(fold, a, l); dt = time.time() - t
    fl, ms, _ = p.counters()
print(f"Epis Gaussian N={X.shape[0]} K={K} Kc={K*(K+1)//2} folds={nf}: {fold.size} fits, lambda_max+grid {t_grid*1e3:.0f} ms, kernel {ms:.0f} ms, "
      f"{fold.size/dt:.1f} fits/s, alg {fl/1e12:.2f} TFLOP -> {fl/ms/1e9:.2f} TFLOP/s, maxM {ns.max()}, status {np.unique(st)}")
