"""Generates tests/golden/*.npz.  Run HERE (the build container), where /root/reference exists:
    python tests/golden/make_golden.py [--binomial]
Inputs come from the reference's bundled data (/root/reference/data/*.rda); expected outputs
come from the reference's own C (oracle/_ref/libeben_ref.so, compiled unmodified) driven through
the numpy restatement of the R layer (oracle/rlayer.py).  The GPU box has no /root/reference,
so the `-m gpu` tests read only these files.
"""
from __future__ import annotations

import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

from oracle import rlayer as R  # noqa: E402
from oracle.rdata import read_rda, read_rds  # noqa: E402
from oracle.rrng import RRng  # noqa: E402

REF = "/root/reference"


def _task(args):
    X, y, fid, f, lam, a, epis, prior = args
    lib = R.fit_lib("reference")
    e, fit = R.fit_one(X, y, fid, f, lam, a, epis, prior, lib)
    nsel = 0 if fit.weight[0, 0] == 0 else fit.weight.shape[0]
    return e, nsel


def grid_table(X, y, n_folds, epis, prior, rows, procs=8):
    ga, gl = R.build_grid(X, y, n_folds, epis)
    fid = R.assign_to_folds(X.shape[0], n_folds)
    jobs = [(X, y, fid, f, gl[r], ga[r], epis, prior) for r in rows for f in range(1, n_folds + 1)]
    with Pool(procs) as pool:
        res = pool.map(_task, jobs, chunksize=1)
    err = np.array([r[0] for r in res]).reshape(len(rows), n_folds)
    nsel = np.array([r[1] for r in res], dtype=np.int32).reshape(len(rows), n_folds)
    return dict(grid_alpha=ga, grid_lambda=gl, fold_id=fid, rows=np.asarray(rows), fold_err=err, n_selected=nsel)


def main():
    B = read_rda(f"{REF}/data/BASIS.rda")["BASIS"]
    y = read_rda(f"{REF}/data/y.rda")["y"].ravel()
    Bb = read_rda(f"{REF}/data/BASISbinomial.rda")["BASISbinomial"]
    yb = read_rda(f"{REF}/data/yBinomial.rda")["yBinomial"].ravel()
    np.savez_compressed(f"{HERE}/inputs_bundled.npz", BASIS=B.astype(np.int8), y=y,
                        BASISbinomial=Bb.astype(np.int8), yBinomial=yb.astype(np.int8))

    # R RNG known answers + the fold vectors the configs use
    np.savez(f"{HERE}/rrng.npz",
             sample10_rejection=np.array(RRng(1, "Rejection").sample_int(10)),
             sample10_rounding=np.array(RRng(1, "Rounding").sample_int(10)),
             folds_50_3=R.assign_to_folds(50, 3), folds_500_5=R.assign_to_folds(500, 5),
             folds_1000_10=R.assign_to_folds(1000, 10), folds_50_3_rounding=R.assign_to_folds(50, 3, "Rounding"))

    # config 1: the README example, whole grid
    X1, y1 = B[:50, :100].astype(float), y[:50]
    g = grid_table(X1, y1, 3, False, "gaussian", range(400))
    s = R.summarise(g["grid_alpha"], g["grid_lambda"], g["fold_err"])
    a, l, _ = R.select_optimum(s)
    ls = R.local_search_replay(g["grid_alpha"], g["grid_lambda"], g["fold_err"])
    np.savez_compressed(f"{HERE}/config1_gaussian.npz", alpha_optimal=a, lambda_optimal=l, summary_mse=s["MSE"],
                        summary_se=s["SE"], summary_alpha=s["alpha"], summary_lambda=s["lambda"],
                        local_cv=ls[0], local_alpha=ls[1], local_lambda=ls[2], local_full=ls[3], **g)
    print("config1", a, l)

    # bundled Gaussian 1000 x 481, 3 folds, every 23rd grid row (active sets up to ~200)
    g = grid_table(B.astype(float), y, 3, False, "gaussian", range(0, 400, 23))
    np.savez_compressed(f"{HERE}/gauss_bundled_sample.npz", **g)
    print("bundled sample max nsel", g["n_selected"].max())

    # Gaussian Epis on a 120 x 25 slice (325 candidates), every 9th row
    g = grid_table(B[:120, :25].astype(float), y[:120], 3, True, "gaussian", range(0, 400, 9))
    np.savez_compressed(f"{HERE}/gauss_epis_slice.npz", **g)

    # final-model outputs (`.C` layouts) on all rows
    fit = R.eb_elastic_net_gaussian(X1, y1, l, a, False, R.fit_lib("reference"))
    np.savez(f"{HERE}/final_gaussian_config1.npz", lam=l, alpha=a, raw_beta=fit.raw_beta, wald=fit.wald,
             intercept=fit.intercept, resid_var=fit.resid_var)

    # the published CrossValidate output (paper_materials): pins grid construction, ordering, SE, argmin
    r = read_rds(f"{REF}/paper_materials/Real Data Analysis/10000_Features/LooserSubset_10000_ParCV_5-3-2018.RDS")
    d, sm = r["Results.Detail"], r["Results.Summary"]
    np.savez_compressed(f"{HERE}/published_cv_10000.npz", detail_fold=np.asarray(d["foldId"]), detail_alpha=d["alpha"],
                        detail_lambda=d["lambda"], detail_mse=d["MSE"], summary_alpha=sm["alpha"],
                        summary_lambda=sm["lambda"], summary_se=sm["SE"], summary_mse=sm["MSE"],
                        lambda_optimal=r["lambda.optimal"], alpha_optimal=r["alpha.optimal"])

    if "--binomial" in sys.argv:
        Xb = Bb.astype(float)
        g = grid_table(Xb, yb.astype(float), 5, False, "binomial", range(400))
        s = R.summarise(g["grid_alpha"], g["grid_lambda"], g["fold_err"], "binomial")
        a, l, _ = R.select_optimum(s, "binomial")
        np.savez_compressed(f"{HERE}/config2_binomial.npz", alpha_optimal=a, lambda_optimal=l,
                            summary_likelihood=s["Likelihood"], summary_se=s["SE"], **g)
        print("config2", a, l)
        g = grid_table(Bb[::4, :20].astype(float), yb[::4].astype(float), 3, True, "binomial", range(0, 400, 9))
        np.savez_compressed(f"{HERE}/binom_epis_slice.npz", **g)


if __name__ == "__main__":
    main()
