"""Gram organisation of the Gaussian fits against the per-fit cache (PAREBEN_GRAM=0): same tables, time of both.
usage: python scripts/gram_check.py [stride]   (bundled Gaussian 1000 x 481, 10 folds; every `stride`-th fit of the 4,000)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb

g = np.load("tests/golden/inputs_bundled.npz")
X, y, nf = g["BASIS"].astype(float), g["y"], 10
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 1
folds = pb.AssignToFolds(X, nf); grid = pb.BuildGrid(X, y, nf)
fold = np.tile(np.arange(1, nf + 1), 400); a = np.repeat(grid["alpha"], nf); l = np.repeat(grid["lambda"], nf)
sel = np.arange(0, fold.size, stride); fold, a, l = fold[sel], a[sel], l[sel]
res = {}
for mode in ("0", "1"):
    os.environ["PAREBEN_GRAM"] = mode
    with pb.Problem(X, y, folds, nf, False, "gaussian") as p:
        for rep in range(3):
            t0 = time.perf_counter()
            err, st, ns, it = p.run_fits(fold, a, l)
            wall = time.perf_counter() - t0
            fl, ms, nl = p.counters()
            print(f"gram={mode} rep {rep}: {fold.size} fits kernel {ms:.1f} ms (wall {wall*1e3:.1f}) launches {nl} model {fl/ms/1e9:.2f} TFLOP/s maxM {ns.max()} status {np.unique(st)}", flush=True)
        res[mode] = (err.copy(), ns.copy(), it.copy())
e0, n0, i0 = res["0"]; e1, n1, i1 = res["1"]
rel = np.abs(e1 - e0) / np.abs(e0)
print(f"gram vs cache: max rel fold-error difference {rel.max():.3e}, p99 {np.quantile(rel, 0.99):.3e}, support-size mismatches {(n0 != n1).sum()}, outer-iteration mismatches {(i0 != i1).sum()}")
