"""Generates tests/golden/yeast_epi008.npz: the BASELINE config-4 stand-in with REAL data.

Run HERE (the build container), where /root/reference exists:   python tests/golden/make_yeast_golden.py
Input: the yeast genotype matrix the reference's authors used for their Epis runs,
  /root/reference/paper_materials/Real Data Analysis/Full_Test/filter_matrix_epi0.08.zip  (3844 x 201, +-1)
  /root/reference/paper_materials/Real Data Analysis/Full_Test/pheno_Zeo_residual         (3844 phenotypes)
(`data/yeastSubset.rda` itself is not in the checkout, .MISSING_LARGE_BLOBS:2.)  CrossValidate(..., nFolds = 10,
Epis = "yes", prior = "gaussian"): 20,301 candidate columns generated on the fly.  Expected outputs come from the
reference's own C (oracle/_ref, compiled unmodified) driven through the numpy restatement of the R layer, for 9 grid
rows (lambda ranks 13, 14, 15 x alpha 1, 0.5, 0.05) x 10 folds -- seconds per fit on a CPU core.  Other parts of this
grid are left out because single reference fits there run for many minutes (the Epis lambda_max is computed on the
un-normalised response, R/BuildGrid.R:26, so the grid is far off the useful range for this data).  The GPU box has no /root/reference: the -m gpu test reads only this file.
"""
import os
import sys
import zipfile
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
from oracle import rlayer as R  # noqa: E402

D = "/root/reference/paper_materials/Real Data Analysis/Full_Test/"


def _task(args):
    X, y, fid, f, lam, a = args
    e, fit = R.fit_one(X, y, fid, f, lam, a, True, "gaussian", R.fit_lib("reference"))
    return e, (0 if fit.weight[0, 0] == 0 else fit.weight.shape[0])


def main():
    lines = zipfile.ZipFile(D + "filter_matrix_epi0.08.zip").read("filter_matrix_epi0.08").decode().strip().split("\n")
    X = np.array([[float(v) for v in ln.rstrip("\t").split("\t")[1:]] for ln in lines])
    y = np.array([float(v) for v in open(D + "pheno_Zeo_residual").read().split()])
    assert X.shape == (3844, 201) and y.shape == (3844,) and set(np.unique(X)) == {-1.0, 1.0}
    n_folds = 10
    ga, gl = R.build_grid(X, y, n_folds, True)
    fid = R.assign_to_folds(X.shape[0], n_folds)
    rows = np.array([lr * 20 + ai for lr in (13, 14, 15) for ai in (0, 10, 19)])
    jobs = [(X, y, fid, f, gl[r], ga[r]) for r in rows for f in range(1, n_folds + 1)]
    with Pool(8) as pool:
        res = pool.map(_task, jobs, chunksize=1)
    err = np.array([r[0] for r in res]).reshape(rows.size, n_folds)
    nsel = np.array([r[1] for r in res], np.int32).reshape(rows.size, n_folds)
    np.savez_compressed(HERE + "/yeast_epi008.npz", X=X.astype(np.int8), y=y, grid_alpha=ga, grid_lambda=gl, fold_id=fid,
                        rows=rows, fold_err=err, n_selected=nsel, lambda_max=R.get_lambda_max(X, y, True))
    print("rows", rows.size, "fits", err.size, "n_selected range", nsel.min(), nsel.max(), "file bytes", os.path.getsize(HERE + "/yeast_epi008.npz"))


if __name__ == "__main__":
    main()
