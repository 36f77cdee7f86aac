"""Binomial score contraction on the INT8 tensor cores (imma_contract.cuh) against the FP64 DMMA path (PAREBEN_NO_IMMA=1):
same tables, time of both.  Config 2: bundled BASISbinomial 500 x 481, 5 folds, the full 2,000-fit grid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb

g = np.load("tests/golden/inputs_bundled.npz")
X, y, nf = g["BASISbinomial"].astype(float), g["yBinomial"].astype(float), 5
folds = pb.AssignToFolds(X, nf); grid = pb.BuildGrid(X, y, nf)
fold = np.tile(np.arange(1, nf + 1), 400); a = np.repeat(grid["alpha"], nf); l = np.repeat(grid["lambda"], nf)
res = {}
for mode in ("dmma", "imma"):
    if mode == "dmma": os.environ["PAREBEN_NO_IMMA"] = "1"
    else: os.environ.pop("PAREBEN_NO_IMMA", None)
    with pb.Problem(X, y, folds, nf, False, "binomial") as p:
        for rep in range(3):
            err, st, ns, it = p.run_fits(fold, a, l)
            fl, ms, nl = p.counters()
            print(f"{mode} rep {rep}: {fold.size} fits kernel {ms:.1f} ms model {fl/ms/1e9:.2f} TFLOP/s maxM {ns.max()} status {np.unique(st)}", flush=True)
        res[mode] = (err.copy(), ns.copy(), it.copy())
e0, n0, i0 = res["dmma"]; e1, n1, i1 = res["imma"]
rel = np.abs(e1 - e0) / np.abs(e0)
print(f"imma vs dmma: max rel fold-error difference {rel.max():.3e}, p99 {np.quantile(rel, 0.99):.3e}, median {np.median(rel):.3e}, support-size mismatches {(n0 != n1).sum()}, outer-iteration mismatches {(i0 != i1).sum()}")
bad = np.argsort(-rel)[:5]
print("worst:", [(int(i), float(rel[i]), int(n0[i]), int(n1[i])) for i in bad])
