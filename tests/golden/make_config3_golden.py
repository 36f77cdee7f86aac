"""Generates tests/golden/config3_yeast.npz: BASELINE config 3 on its REAL stand-in, at the stated size.

Run HERE (the build container), where /root/reference exists:   python tests/golden/make_config3_golden.py
`data/yeastFull.rda` is not in the checkout; the stand-in is what the reference's own timing script builds
(paper_materials/Timing Tests/test_time_Gaus.R:13-19, Real Data Analysis/SL_filter.R:5-15):
  genotype_full.txt (in genotype_full.zip: 4390 strains x 28,220 markers, +-1)  joined with  pheno_left.txt (3803 ids)
  -> BASIS 3803 x 28,220, y = phenotype column 2.  Read with pareben_b200.io (the R-free reader).
CrossValidate(BASIS, y, nFolds = 10, Epis = "no", prior = "gaussian"): grid from BuildGrid, folds from R >= 3.6
`set.seed(1); sample()`.  Expected values: the reference's own C (oracle/_ref, elasticNetLinearNeMainEff compiled
unmodified) for 40 grid rows (every lambda rank x alpha in {1, 0.5}) on fold 1 -- each fit in its own process with a
time limit, because below some lambda the reference's active set reaches basisMax = 1e7/K = 354 and it then writes
past its buffers (elasticNetLinearNeMainEff.c:605-611): such fits are recorded as NaN and the GPU test
only checks that the kernel flags them (status bit BASIS_CAP).
The matrix travels bit-packed (+1 -> 1, -1 -> 0): 13 MB raw, less compressed.
"""
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
from oracle import rlayer as R  # noqa: E402
from pareben_b200 import io as pio  # noqa: E402

D = "/root/reference/paper_materials/Timing Tests/"
LIMIT_S = 2400

WORKER = r"""
import sys, os, numpy as np
sys.path.insert(0, %r)
os.environ["OPENBLAS_NUM_THREADS"] = "1"
from oracle import rlayer as R
z = np.load(sys.argv[1]); lam = float(sys.argv[2]); alpha = float(sys.argv[3])
fit = R.eb_elastic_net_gaussian(z["Xtr"], z["ytr"], lam, alpha, False, R.fit_lib("reference"))
e = R.get_fold_error(z["Xte"], z["yte"], fit, "gaussian")
m = 0 if fit.weight[0, 0] == 0 else fit.weight.shape[0]
print("RESULT %%.17g %%d" %% (e, m))
"""


def main():
    X, y = pio.load_problem(D + "genotype_full.zip", D + "pheno_left.txt")
    assert X.shape == (3803, 28220) and set(np.unique(X)) == {-1, 1}
    Xf = X.astype(np.float64)
    n_folds = 10
    ga, gl = R.build_grid(Xf, y, n_folds, False)
    fid = R.assign_to_folds(X.shape[0], n_folds)
    rows = np.array([lr * 20 + ai for lr in range(20) for ai in (0, 10)])
    tr, te = fid != 1, fid == 1
    tmp = tempfile.mkdtemp()
    dpath = os.path.join(tmp, "fold1.npz")
    np.savez(dpath, Xtr=Xf[tr], ytr=y[tr], Xte=Xf[te], yte=y[te])

    def run(r):
        try:
            out = subprocess.run([sys.executable, "-c", WORKER % ROOT, dpath, repr(float(gl[r])), repr(float(ga[r]))],
                                 capture_output=True, text=True, timeout=LIMIT_S)
            for line in out.stdout.splitlines():
                if line.startswith("RESULT"):
                    _, e, m = line.split()
                    return float(e), int(m)
            return float("nan"), -2                        # crashed (heap overrun past basisMax)
        except subprocess.TimeoutExpired:
            return float("nan"), -3

    order = sorted(range(rows.size), key=lambda i: gl[rows[i]])          # long fits first
    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(lambda i: run(rows[i]), order))
    err = np.full(rows.size, np.nan); nsel = np.full(rows.size, -1, np.int32)
    for i, (e, m) in zip(order, res):
        err[i] = e; nsel[i] = m
        print(f"row {rows[i]:3d} lambda {gl[rows[i]]:.5g} alpha {ga[rows[i]]:.2f}: err {e:.10g} M {m}", flush=True)
    np.savez_compressed(HERE + "/config3_yeast.npz", bits=np.packbits(X == 1, axis=1), n=X.shape[0], k=X.shape[1], y=y,
                        grid_alpha=ga, grid_lambda=gl, fold_id=fid, rows=rows, fold=1, fold_err=err, n_selected=nsel,
                        lambda_max=R.get_lambda_max(Xf, y, False))
    print("file bytes", os.path.getsize(HERE + "/config3_yeast.npz"), "fits with a reference value", int(np.isfinite(err).sum()))


if __name__ == "__main__":
    main()
