import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pareben_b200 as pb
from oracle import rlayer as R
def _data(seed, n, k):
    rng = np.random.default_rng(seed)
    X = rng.choice([-1.0, 0.0, 1.0], size=(n, k), p=[0.25, 0.5, 0.25])
    y = 20 + 2.5 * X[:, 1] - 2.0 * X[:, k // 2] + 2.2 * X[:, 3] * X[:, k - 2] + rng.normal(0, 1.5, n)
    return X, y
X, y = _data(3, 160, 24)
ref = R.fit_lib(R.available_kind())
for lam, al in ((1.5, 1.0), (0.6, 0.5)):
    want = R.eb_elastic_net_gaussian(X, y, lam, al, True, ref)
    print("ref   ", lam, al, want.weight[:, :3].tolist(), want.intercept, want.resid_var)
    for mode in (pb.MODE_CACHED, pb.MODE_STREAMING):
        pb.set_mode(mode)
        with pb.Problem(X, y, None, 0, True, "gaussian") as p:
            t, wald, icpt, resid, st = p.fit(al, lam)
        nz = t[t[:, 4] != 0]
        print("mode", mode, lam, al, nz[:, :3].tolist(), icpt[0], resid, "status", st)
    pb.set_mode(pb.MODE_AUTO)
    # the same rows as a 'fold': append 40 held-out rows so that training rows = the 160
    X2 = np.vstack([X, X[:40]]); y2 = np.concatenate([y, y[:40]])
    fid = np.concatenate([np.full(160, 2), np.full(40, 1)]).astype(np.int32)
    err, st, ns = pb.cv_grid(X2, y2, fid, 2, np.array([al]), np.array([lam]), epis=True)
    e, fit = R.fit_one(X2, y2, fid, 1, lam, al, True, "gaussian", ref)
    print("cv fold1 gpu", err[0, 0], "ref", e, "ns", ns[0, 0], fit.weight.shape[0], "status", st[0, 0])
