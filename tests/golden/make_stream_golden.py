"""Generates tests/golden/stream_k2000.npz: a K = 2,000-locus slice of BASELINE config 5 (Epis = "yes": 2,001,000
candidates), the parity anchor of the streaming kernels at a candidate count the per-fit cache was never meant for.

Run HERE (the build container), where /root/reference exists:   python tests/golden/make_stream_golden.py
Data: pareben_b200.synth.config5(n = 300, k = 2000) (seeded numpy PCG64; the generator's parameters are stored in the
file, not the matrix).  Folds: R >= 3.6 `set.seed(1); sample(...)`, 3 folds.  Expected outputs come from the reference's
own C (oracle/_ref/libeben_ref.so, elasticNetLinearNeEpisEff compiled unmodified): fold errors and support sizes for
  * rows of BuildGrid's own Epis grid (lambda from the un-normalised pair scan, R/BuildGrid.R:24-27): every fit keeps
    only its initial basis -- the regime of config 5 at full size;
  * smaller lambdas, where the fits add, re-estimate and delete main and pair effects (one with ~110 of them).
Minutes of CPU per non-trivial fit (CacheBP over 2e6 candidates per outer iteration), hence the short list.
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
from oracle import rlayer as R  # noqa: E402
from pareben_b200.synth import config5  # noqa: E402

N, K, SEED, NF = 300, 2000, 20260101, 3


def _task(args):
    X, y, fid, f, lam, a = args
    e, fit = R.fit_one(X, y, fid, f, lam, a, True, "gaussian", R.fit_lib("reference"))
    return e, (0 if fit.weight[0, 0] == 0 else fit.weight.shape[0])


def main():
    d = config5(N, K, SEED)
    X = d["X"].astype(np.float64); y = d["y"]
    fid = R.assign_to_folds(N, NF)
    ga, gl = R.build_grid(X, y, NF, True)
    rows = np.array([0, 170, 399])
    # on this data the active set jumps from 1 to ~110 between lambda = 0.08 and 0.07 (alpha = 1): both sides are kept;
    # the large fit (7.5 minutes of reference C) for fold 1 only
    lam = np.concatenate([gl[rows], [1.0, 0.1, 0.08, 0.07]])
    alpha = np.concatenate([ga[rows], [0.05, 1.0, 1.0, 1.0]])
    jobs = [(X, y, fid, f, lam[i], alpha[i]) for i in range(lam.size) for f in range(1, NF + 1) if i < lam.size - 1 or f == 1]
    jobs.sort(key=lambda j: j[4])
    with Pool(8) as pool:
        res = pool.map(_task, jobs, chunksize=1)
    err = np.full((lam.size, NF), np.nan); nsel = np.full((lam.size, NF), -1, np.int32)
    for j, r in zip(jobs, res):
        i = int(np.flatnonzero((lam == j[4]) & (alpha == j[5]))[0])
        err[i, j[3] - 1] = r[0]; nsel[i, j[3] - 1] = r[1]
    np.savez_compressed(HERE + "/stream_k2000.npz", n=N, k=K, seed=SEED, n_folds=NF, fold_id=fid, lam=lam, alpha=alpha,
                        fold_err=err, n_selected=nsel, lambda_max=R.get_lambda_max(X, y, True), y_check=y[:8])
    print("fits", err.size, "n_selected", nsel.tolist(), "file bytes", os.path.getsize(HERE + "/stream_k2000.npz"))


if __name__ == "__main__":
    main()
