"""GPU box: config 2 (binomial, 2,000 fits) against the golden table; prints the fits that differ most."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb
g = np.load("tests/golden/config2_binomial.npz"); b = np.load("tests/golden/inputs_bundled.npz")
X, y = b["BASISbinomial"].astype(float), b["yBinomial"].astype(float)
err, st, ns = pb.cv_grid(X, y, g["fold_id"], 5, g["grid_alpha"], g["grid_lambda"], prior="binomial")
rel = np.abs(err - g["fold_err"]) / np.abs(g["fold_err"])
for i in np.argsort(rel.ravel())[::-1][:5]:
    r, f = divmod(int(i), 5)
    print(f"row {r} fold {f + 1} lambda {g['grid_lambda'][r]:.17g} alpha {g['grid_alpha'][r]:.17g} gpu {err[r, f]:.17g} ref {g['fold_err'][r, f]:.17g} rel {rel[r, f]:.3e} M {ns[r, f]}")
