"""CPU suite, part 3: the host-side mirror of the R layer (pareben_b200/cross_validate.py, rcompat.py)
against the oracle's independent restatement, plus the world_size-2 gloo merge path."""
import math
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT, golden
from oracle import rlayer as R


def test_assign_to_folds_matches_oracle():
    import pareben_b200 as pb
    g = golden("rrng.npz")
    assert np.array_equal(pb.AssignToFolds(np.zeros((50, 2)), 3), g["folds_50_3"])
    assert np.array_equal(pb.AssignToFolds(np.zeros((500, 2)), 5), g["folds_500_5"])
    assert np.array_equal(pb.AssignToFolds(np.zeros((1000, 2)), 10), g["folds_1000_10"])
    assert np.array_equal(pb.AssignToFolds(np.zeros((50, 2)), 3, sample_kind="Rounding"), g["folds_50_3_rounding"])
    mine = np.arange(50) % 3 + 1
    assert np.array_equal(pb.AssignToFolds(np.zeros((50, 2)), 3, mine), mine)       # user foldId of length N passes through


def test_rcompat_numerics():
    from pareben_b200 import rcompat
    x = np.array([2246.40804244, 2004.15893851, 1558.65730397])
    assert rcompat.mean(x) == R.r_mean(x) and rcompat.sd(x) == R.r_sd(x)
    assert np.array_equal(rcompat.seq(1.0, 0.05, -0.05), R.r_seq(1.0, 0.05, -0.05))
    a = rcompat.seq(1.0, 0.05, -0.05)
    assert a.size == 20 and a[-1] == 0.05 and a[-2] == 1 + 18 * (-0.05)


def test_summary_and_local_search_against_oracle(monkeypatch):
    """Feed the host layer the reference's config-1 error table through a stubbed device call and
    compare every derived object with the oracle's restatement."""
    import pareben_b200 as pb
    from pareben_b200 import cross_validate as cv
    g = golden("config1_gaussian.npz")
    X = golden("inputs_bundled.npz")["BASIS"][:50, :100].astype(float)
    y = golden("inputs_bundled.npz")["y"][:50]
    grid = {"alpha": g["grid_alpha"], "lambda": g["grid_lambda"]}
    monkeypatch.setattr(cv, "_grid_and_errors", lambda *a, **k: (grid, g["fold_err"].copy(), np.zeros((400, 3), np.int32), g["n_selected"]))
    out = pb.CrossValidate(X, y, 3)
    assert out["alpha.optimal"] == float(g["alpha_optimal"]) and out["lambda.optimal"] == float(g["lambda_optimal"])
    assert np.array_equal(out["Results.Summary"]["MSE"], g["summary_mse"])
    assert np.array_equal(out["Results.Summary"]["SE"], g["summary_se"])
    assert np.array_equal(out["Results.Detail"]["foldId"][:4], [1, 2, 3, 1])
    assert np.array_equal(out["Results.Detail"]["MSE"], g["fold_err"].ravel())
    loc = pb.CrossValidate(X, y, 3, search="local")
    assert np.array_equal(loc["CrossValidation"], g["local_cv"]) and np.array_equal(loc["fullCV"], g["local_full"])
    assert loc["alpha.optimal"] == float(g["local_alpha"]) and loc["lambda.optimal"] == float(g["local_lambda"])


def test_which_min_skips_na_and_reports_status(monkeypatch):
    """R's which.min skips NA (np.argmin would return the first NaN); a fit flagged by its status word must not
    pass silently; an all-NA table stops like the reference's is.na(Mu0) check."""
    import warnings
    import pytest
    import pareben_b200 as pb
    from pareben_b200 import cross_validate as cv
    g = golden("config1_gaussian.npz")
    grid = {"alpha": g["grid_alpha"], "lambda": g["grid_lambda"]}
    err = g["fold_err"].copy()
    order = np.lexsort((grid["lambda"], grid["alpha"]))
    err[order[0], 1] = np.nan                       # first row of the sorted summary becomes NA
    st = np.zeros((400, 3), np.int32); st[order[0], 1] = pb.FIT_NONFINITE; st[5, 0] = pb.FIT_BASIS_CAP
    monkeypatch.setattr(cv, "_grid_and_errors", lambda *a, **k: (grid, err.copy(), st, g["n_selected"]))
    X = np.zeros((50, 100)); y = np.zeros(50)
    with pytest.warns(RuntimeWarning, match="2 of 1200 fits.*basis cap 1.*non-finite 1"):
        out = pb.CrossValidate(X, y, 3)
    assert out["alpha.optimal"] == float(g["alpha_optimal"]) and out["lambda.optimal"] == float(g["lambda_optimal"])
    assert np.isnan(out["Results.Summary"]["MSE"][0])
    with pytest.warns(RuntimeWarning):
        loc = pb.CrossValidate(X, y, 3, search="local")
    assert np.isfinite(loc["alpha.optimal"])
    monkeypatch.setattr(cv, "_grid_and_errors", lambda *a, **k: (grid, np.full((400, 3), np.nan), np.zeros((400, 3), np.int32), g["n_selected"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(ValueError, match="NA"):
            pb.CrossValidate(X, y, 3)


def test_cv_grid_validates_lengths(built):
    """The C side only sees pointers: mismatched lengths must raise before the ctypes call (no device needed)."""
    import pytest
    import pareben_b200 as pb
    X = np.zeros((10, 3)); y = np.arange(10.0); f = np.tile([1, 2], 5)
    with pytest.raises(ValueError, match="Target"):
        pb.cv_grid(X, y[:9], f, 2, np.ones(2), np.ones(2))
    with pytest.raises(ValueError, match="fold_id"):
        pb.cv_grid(X, y, f[:8], 2, np.ones(2), np.ones(2))
    with pytest.raises(ValueError, match="alpha"):
        pb.cv_grid(X, y, f, 2, np.ones(3), np.ones(2))
    with pytest.raises(ValueError, match="n_folds"):
        pb.cv_grid(X, y, f, 0, np.ones(2), np.ones(2))


def test_default_device_follows_local_rank(monkeypatch):
    import pareben_b200 as pb
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    assert pb.default_device() == 0
    monkeypatch.setenv("LOCAL_RANK", "3")
    assert pb.default_device() == 3


def test_two_rank_gloo_merge(tmp_path):
    """world_size 2 on CPU (gloo): each rank fills only its shard (as pareben_cv_grid does) and the
    host layer's all-reduce must rebuild the full table on both ranks."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        import pareben_b200 as pb
        from pareben_b200 import cross_validate as cv, _lib
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        lam = np.repeat(np.exp(np.linspace(1, -6, 20)), 20); alpha = np.tile(np.linspace(1, .05, 20), 20)
        truth = np.sin(np.arange(1200.0)).reshape(400, 3) + 3
        class FakeProblem:                    # stands where the device-resident problem would be: fills only this rank's shard
            def __init__(self, *a, **k): pass
            def __enter__(self): return self
            def __exit__(self, *e): pass
            def lambda_max(self): return 1.0
            def cv_grid(self, a, l, shard=0, n_shards=1):
                mine = pb.shard_plan(l, 3, shard, n_shards)
                err = np.zeros(1200); st = np.zeros(1200, np.int32); ns = np.zeros(1200, np.int32)
                err[mine] = truth.ravel()[mine]; ns[mine] = mine % 7; st[mine] = (mine % 11 == 0)
                return err.reshape(400, 3), st.reshape(400, 3), ns.reshape(400, 3)
        _lib.Problem = FakeProblem
        cv._grid_from_lambda_max = lambda lm: {{"alpha": alpha, "lambda": lam}}
        grid, err, st, ns = cv._grid_and_errors(None, None, None, 3, "no", "gaussian", 0)
        assert np.array_equal(err, truth), "merged table differs"
        assert np.array_equal(ns.ravel(), np.arange(1200) % 7) and np.array_equal(st.ravel(), (np.arange(1200) % 11 == 0).astype(np.int32))
        dist.destroy_process_group()
        print("rank", rank, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_gram_organisation_identities(bundled):
    """What the Gram organisation of the Gaussian kernels rests on, checked in numpy on the bundled design:
    (i) the candidate-cache row of basis p, sum_h (x_c[h]/s_c)(x_p[h]/s_p) as CacheBP* computes it
    (elasticNetLinearNeMainEff.c:1157-1178), is row p of one matrix per fold -- it involves neither the response nor any
    solver state; for genotype codes the dot products are exact integers, so the shared matrix differs from the reference's
    per-term-rounded sum by a few ulp;
    (ii) ||t - PHI mu||^2 = t't - 2 mu'(PHI't) + mu'(PHI'PHI)mu with PHI't = xt[used] and PHI'PHI = C[used][:, used], to
    1e-13 relative when the residual is 1 % of t't (where the outer loop stops)."""
    import pareben_b200 as pb
    X = bundled["BASIS"].astype(np.float64)
    y = bundled["y"].astype(np.float64)
    folds = pb.AssignToFolds(X, 10)
    tr = folds != 1
    Xf, t = X[tr], y[tr] - y[tr].mean()
    s = np.sqrt((Xf * Xf).sum(axis=0)); s[s == 0] = 1.0
    ints = Xf.T @ Xf
    assert np.array_equal(ints, np.rint(ints))                      # exact integer dot products
    C = ints / np.outer(s, s)
    Z = Xf / s                                                      # the reference normalises BASIS in place (:87-99)
    used = np.array([0, 17, 230, 480, 5, 99])
    for p in used:
        row = np.array([np.dot(Z[:, c], Z[:, p]) for c in range(X.shape[1])])      # CacheBP*'s loop
        assert np.max(np.abs(row - C[p])) < 1e-14                   # correlations, |.| <= 1: a few ulp (an exact 0 in C is 5e-18 there)
    # (ii): a posterior mean that leaves 1 % of the variance -- fit on a response the six columns explain almost fully
    PHI = Z[:, used]
    rng = np.random.default_rng(3)
    t2 = PHI @ rng.normal(0, 1, used.size); t2 += 0.1 * np.linalg.norm(t2) / np.sqrt(t2.size) * rng.normal(0, 1, t2.size)
    mu = np.linalg.lstsq(PHI, t2, rcond=None)[0]
    direct = np.sum((t2 - PHI @ mu) ** 2)
    xt = Z.T @ t2
    gram = t2 @ t2 - 2 * mu @ xt[used] + mu @ C[np.ix_(used, used)] @ mu
    assert direct / (t2 @ t2) < 0.02
    assert abs(gram - direct) / direct < 1e-12
    assert t.size == Xf.shape[0]
