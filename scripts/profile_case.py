"""Small, short case for `ncu --set full`: a 296-fit slice of the config-2 grid (2 blocks per SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pass
import numpy as np
import pareben_b200 as pb

g = np.load("tests/golden/inputs_bundled.npz")
prior = sys.argv[1] if len(sys.argv) > 1 else "binomial"
if prior == "binomial":
    X, y, nf = g["BASISbinomial"].astype(float), g["yBinomial"].astype(float), 5
else:
    X, y, nf = g["BASIS"].astype(float), g["y"], 10
folds = pb.AssignToFolds(X, nf)
grid = pb.BuildGrid(X, y, nf)
n_fits = int(sys.argv[2]) if len(sys.argv) > 2 else 296
rows = np.linspace(0, 399, n_fits // nf + 1).astype(int)
fold = np.tile(np.arange(1, nf + 1), rows.size)[:n_fits]
a = np.repeat(grid["alpha"][rows], nf)[:n_fits]; l = np.repeat(grid["lambda"][rows], nf)[:n_fits]
with pb.Problem(X, y, folds, nf, False, prior) as p:
    for rep in range(2):
        err, st, ns, it = p.run_fits(fold, a, l)
        fl, ms, _ = p.counters()
        print(f"{prior} fits={n_fits} kernel {ms:.1f} ms alg {fl/1e9:.1f} GFLOP -> {fl/ms/1e9:.3f} TFLOP/s maxM {ns.max()} status {np.unique(st)}")
