"""GPU suite (-m gpu): the CUDA path, called through the C-ABI, against the reference's results.

Expected values are the committed golden vectors (produced by the reference's own C, see
tests/golden/make_golden.py) and, where oracle/_ref or the C restatement is loadable on the box,
live oracle fits on fresh inputs.  Tolerance: the path is FP64; reductions are re-associated
relative to OpenBLAS, so per-fit errors must agree to rel 1e-8 (north_star's bound) and the
discrete outputs (support sizes, selected alpha/lambda) must be identical.
"""
import numpy as np
import pytest

from conftest import golden
from oracle import rlayer as R

pytestmark = pytest.mark.gpu
RTOL = 1e-8


@pytest.fixture(scope="module")
def pb(built):
    import pareben_b200 as pb
    if pb.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu suite must run on the B200 box")
    return pb


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_config1_gaussian_full_grid(pb, bundled):
    g = golden("config1_gaussian.npz")
    X, y = bundled["BASIS"][:50, :100].astype(float), bundled["y"][:50]
    err, st, ns = pb.cv_grid(X, y, g["fold_id"], 3, g["grid_alpha"], g["grid_lambda"])
    assert np.all(st == 0)
    assert np.array_equal(ns, g["n_selected"]), f"{(ns != g['n_selected']).sum()} fits with a different support size"
    assert _rel(err, g["fold_err"]) < RTOL
    out = pb.CrossValidate(X, y, 3)
    assert out["alpha.optimal"] == float(g["alpha_optimal"])
    assert abs(out["lambda.optimal"] - float(g["lambda_optimal"])) <= 1e-13 * float(g["lambda_optimal"])
    assert _rel(out["Results.Summary"]["MSE"], g["summary_mse"]) < RTOL
    assert _rel(out["Results.Summary"]["SE"], g["summary_se"]) < 1e-6
    assert np.allclose(out["Results.Summary"]["alpha"], g["summary_alpha"])
    loc = pb.CrossValidate(X, y, 3, foldId=g["fold_id"], search="local")
    assert loc["alpha.optimal"] == float(g["local_alpha"])
    assert abs(loc["lambda.optimal"] - float(g["local_lambda"])) <= 1e-13 * float(g["local_lambda"])
    assert np.array_equal(loc["fullCV"][:, 1] != 0, g["local_full"][:, 1] != 0)       # same early stops
    assert np.allclose(loc["fullCV"], g["local_full"], rtol=1e-8, atol=0)


def test_grid_and_lambda_max_on_device(pb, bundled):
    for (n, k, epis) in ((50, 100, False), (120, 25, True), (1000, 481, False)):
        X, y = bundled["BASIS"][:n, :k].astype(float), bundled["y"][:n]
        want = R.get_lambda_max(X, y, epis)
        got = pb.GetLambdaMax(X, y, "yes" if epis else "no")
        assert abs(got - want) <= 1e-12 * abs(want)
        ga, gl = R.build_grid(X, y, 3, epis)
        grid = pb.BuildGrid(X, y, 3, "yes" if epis else "no")
        assert np.array_equal(grid["alpha"], ga) and np.allclose(grid["lambda"], gl, rtol=1e-12, atol=0)


def test_bundled_gaussian_rows(pb, bundled):
    g = golden("gauss_bundled_sample.npz")
    X, y = bundled["BASIS"].astype(float), bundled["y"]
    rows = g["rows"]
    err, st, ns = pb.cv_grid(X, y, g["fold_id"], 3, g["grid_alpha"][rows], g["grid_lambda"][rows])
    assert np.all(st == 0)
    assert np.array_equal(ns, g["n_selected"])
    assert _rel(err, g["fold_err"]) < RTOL


def test_gaussian_epis_slice(pb, bundled):
    g = golden("gauss_epis_slice.npz")
    X, y = bundled["BASIS"][:120, :25].astype(float), bundled["y"][:120]
    rows = g["rows"]
    err, st, ns = pb.cv_grid(X, y, g["fold_id"], 3, g["grid_alpha"][rows], g["grid_lambda"][rows], epis=True)
    assert np.array_equal(ns, g["n_selected"])
    assert _rel(err, g["fold_err"]) < RTOL


def test_final_model_layout(pb, bundled):
    g = golden("final_gaussian_config1.npz")
    X, y = bundled["BASIS"][:50, :100].astype(float), bundled["y"][:50]
    with pb.Problem(X, y, None, 0, False, "gaussian") as p:
        table, wald, icpt, resid, st = p.fit(float(g["alpha"]), float(g["lam"]))
    ref = g["raw_beta"]
    assert table.shape == ref.shape == (100, 4)
    assert np.array_equal(table[:, :2], ref[:, :2])
    assert np.array_equal(table[:, 2] != 0, ref[:, 2] != 0)
    assert np.allclose(table[:, 2:], ref[:, 2:], rtol=1e-8, atol=1e-14)
    assert abs(icpt[0] - g["intercept"][0]) <= 1e-10 * abs(g["intercept"][0])
    assert abs(resid - float(g["resid_var"])) <= 1e-8 * float(g["resid_var"])
    assert abs(wald - float(g["wald"])) <= 1e-8 * abs(float(g["wald"]))
    fit = pb.EBelasticNet_Gaussian(X, y, float(g["lam"]), float(g["alpha"]))
    assert fit["weight"].shape[1] == 6 and fit["weight"].shape[0] == int((ref[:, 2] != 0).sum())


def test_live_oracle_random_inputs(pb):
    """Fresh seeded inputs (ragged fold sizes, a duplicated column, an all-zero column) against
    whichever oracle library is present on this box."""
    lib = R.fit_lib(R.available_kind())
    rng = np.random.default_rng(11)
    n, k = 101, 40
    X = rng.choice([-1.0, 0.0, 1.0], size=(n, k), p=[0.25, 0.5, 0.25])
    X[:, 17] = X[:, 4]            # exact duplicate: ties must resolve to the lower index
    X[:, 23] = 0.0                # all-zero column: scale forced to 1
    y = 50 + 3 * X[:, 4] - 2 * X[:, 9] + rng.normal(0, 1.5, n)
    folds = R.assign_to_folds(n, 4)
    lam = np.array([3.0, 0.8, 0.2, 0.05, 0.01]); alpha = np.array([1.0, 0.7, 0.5, 0.2, 0.05])
    err, st, ns = pb.cv_grid(X, y, folds, 4, alpha, lam)
    for i in range(lam.size):
        for f in range(1, 5):
            e, fit = R.fit_one(X, y, folds, f, lam[i], alpha[i], lib=lib)
            m = 0 if fit.weight[0, 0] == 0 else fit.weight.shape[0]
            assert ns[i, f - 1] == m
            assert abs(err[i, f - 1] - e) <= RTOL * abs(e)


def test_full_size_properties(pb, bundled):
    """Size-independent properties at the bundled full size (1000 x 481, 10 folds): the table is
    bitwise reproducible, independent of batch composition/order, and shards merge exactly."""
    X, y = bundled["BASIS"].astype(float), bundled["y"]
    folds = pb.AssignToFolds(X, 10)
    grid = pb.BuildGrid(X, y, 10)
    rows = np.arange(0, 400, 16)
    a, l = grid["alpha"][rows], grid["lambda"][rows]
    e1, s1, n1 = pb.cv_grid(X, y, folds, 10, a, l)
    e2, _, _ = pb.cv_grid(X, y, folds, 10, a, l)
    assert np.array_equal(e1, e2)                                   # run-to-run bitwise
    perm = np.random.default_rng(0).permutation(rows.size)
    e3, _, _ = pb.cv_grid(X, y, folds, 10, a[perm], l[perm])
    assert np.array_equal(e3, e1[perm])                             # batch order does not matter
    merged = np.zeros_like(e1)
    for r in range(3):
        es, _, _ = pb.cv_grid(X, y, folds, 10, a, l, shard=r, n_shards=3)
        merged += es
    assert np.array_equal(merged, e1)                               # 3 shards == 1 shard, bitwise
    assert np.all(np.isfinite(e1)) and np.all(e1 > 0) and np.all(s1 == 0)
    # SSE at the largest lambda equals the intercept-only prediction error when nothing is selected
    empty = n1 == 0
    for i, f in zip(*np.nonzero(empty)):
        te = folds == f + 1
        tr = ~te
        assert abs(e1[i, f] - np.sum((y[te] - y[tr].mean()) ** 2)) <= 1e-9 * e1[i, f]


def test_edge_cases(pb):
    rng = np.random.default_rng(3)
    X = rng.normal(size=(30, 6)); y = rng.normal(size=30) + 10
    folds = np.arange(30) % 3 + 1
    with pytest.raises(pb.ParebenError):
        pb.cv_grid(X, y, np.zeros(30, np.int32), 3, np.ones(1), np.ones(1))          # fold labels out of range
    err, st, ns = pb.cv_grid(X, y, folds, 3, np.ones(1), np.array([1e6]))           # huge lambda: empty model
    assert np.all(np.isfinite(err))
    err2, _, _ = pb.cv_grid(X[:, :1], y, folds, 3, np.ones(2), np.array([1.0, 0.01]))   # K = 1
    assert err2.shape == (2, 3) and np.all(np.isfinite(err2))


def test_config2_binomial_full_grid(pb, bundled):
    """BASELINE config 2: bundled 500 x 481 logistic data, 5 folds, all 2,000 fits."""
    g = golden("config2_binomial.npz")
    X, y = bundled["BASISbinomial"].astype(float), bundled["yBinomial"].astype(float)
    err, st, ns = pb.cv_grid(X, y, g["fold_id"], 5, g["grid_alpha"], g["grid_lambda"], prior="binomial")
    assert np.all(st == 0)
    diff = ns != g["n_selected"]
    assert diff.sum() == 0, f"{diff.sum()} of 2000 fits with a different support size"
    # Binomial tolerance.  The reference's IRLS accepts/rejects Newton steps on `newTotalError >= errorLog`
    # (NEmainEff.c:1991) and stops at |g| < 1e-6: near convergence the objective change of a step (~g^2/H ~ 1e-13) is
    # below the rounding error of the objective itself (~1e-16 * 270), so the outcome -- and with it the final weights at
    # the 1e-8 level -- is decided by the ORDER in which the N log-terms are added.  This is demonstrated, not assumed:
    # tests/test_oracle.py::test_binomial_irls_accept_reject_depends_on_summation_order reverses that one sum inside the
    # reference's own algorithm and 13 of the 100 top-lambda fits move by up to 1.07e-8
    # (tests/golden/make_binomial_alt_golden.py), each with a handful of such decision points.  A parallel reduction cannot
    # add in the reference's order, so on those 100 fits the kernel is held to the demonstrated sensitivity (2e-8; measured:
    # the same fits move, by the same 5e-9 .. 1.07e-8, most of them landing exactly on the reversed-sum branch) and
    # everywhere else to 1e-9 (measured 1.6e-10).
    rel = np.abs(err - g["fold_err"]) / np.abs(g["fold_err"])
    alt = golden("config2_binomial_alt.npz")
    rows = alt["rows"]
    rel_alt = np.abs(err[rows] - alt["fold_err_reversed_sum"]) / np.abs(alt["fold_err_reversed_sum"])
    assert rel[rows].max() < 2e-8
    moved = np.abs(alt["fold_err_reversed_sum"] - g["fold_err"][rows]) / np.abs(g["fold_err"][rows]) > 1e-10
    assert np.all(rel[rows][~moved] < 1e-9), "a fit the summation order does not move must agree tightly"
    assert (np.minimum(rel[rows], rel_alt) < 1e-10).mean() > 0.95
    rest = np.ones(err.shape[0], bool); rest[rows] = False
    assert rel[rest].max() < 1e-9
    assert np.quantile(rel, 0.99) < 1e-12
    out = pb.CrossValidate(X, y, 5, prior="binomial")
    assert abs(out["alpha.optimal"] - float(g["alpha_optimal"])) < 1e-15
    assert abs(out["lambda.optimal"] - float(g["lambda_optimal"])) <= 1e-13 * float(g["lambda_optimal"])
    assert _rel(out["Results.Summary"]["Likelihood"], g["summary_likelihood"]) < 2e-8
    loc = pb.CrossValidate(X, y, 5, foldId=g["fold_id"], prior="binomial", search="local")
    want = R.local_search_replay(g["grid_alpha"], g["grid_lambda"], -g["fold_err"])
    assert loc["alpha.optimal"] == want[1] and abs(loc["lambda.optimal"] - want[2]) <= 1e-13 * want[2]
    assert np.allclose(loc["fullCV"][:, :3], want[3][:, :3], rtol=2e-8, atol=0)
    assert np.allclose(loc["fullCV"][:, 3], want[3][:, 3], rtol=1e-4, atol=0)      # SE = sd/sqrt(n): differences of near-equal numbers


def test_binomial_epis_slice(pb, bundled):
    g = golden("binom_epis_slice.npz")
    X, y = bundled["BASISbinomial"][::4, :20].astype(float), bundled["yBinomial"][::4].astype(float)
    rows = g["rows"]
    err, st, ns = pb.cv_grid(X, y, g["fold_id"], 3, g["grid_alpha"][rows], g["grid_lambda"][rows], epis=True, prior="binomial")
    assert np.array_equal(ns, g["n_selected"])
    assert np.max(np.abs(err - g["fold_err"]) / np.maximum(np.abs(g["fold_err"]), 1e-12)) < RTOL


def test_binomial_final_model_live(pb, bundled):
    lib = R.fit_lib(R.available_kind())
    X, y = bundled["BASISbinomial"][:, :120].astype(float), bundled["yBinomial"].astype(float)
    rng = np.random.default_rng(5)
    rows = rng.permutation(500)[:300]
    X, y = X[rows], y[rows]
    for lam, a in ((0.5, 1.0), (0.05, 0.5)):
        want = R.eb_elastic_net_binomial(X, y, lam, a, False, lib)
        got = pb.EBelasticNet_Binomial(X, y, lam, a)
        assert got["weight"].shape == want.weight.shape
        assert np.array_equal(got["weight"][:, :2], want.weight[:, :2])
        assert np.allclose(got["weight"][:, 2:4], want.weight[:, 2:4], rtol=5e-8, atol=1e-12)
        assert np.allclose(got["Intercept"], want.intercept, rtol=5e-8)
        assert abs(got["logLikelihood"] - want.log_likelihood) <= 1e-8 * abs(want.log_likelihood)


def test_one_call_over_two_devices(pb, bundled):
    """pareben_cv_grid(n_devices = 2): one call, two GPUs, in-library gather -- bitwise the single-GPU table."""
    if pb.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = golden("config2_binomial.npz")
    X, y = bundled["BASISbinomial"].astype(float), bundled["yBinomial"].astype(float)
    rows = np.arange(0, 400, 3)
    a, l = g["grid_alpha"][rows], g["grid_lambda"][rows]
    e1, s1, n1 = pb.cv_grid(X, y, g["fold_id"], 5, a, l, prior="binomial")
    e2, s2, n2 = pb.cv_grid(X, y, g["fold_id"], 5, a, l, prior="binomial", n_devices=2)
    e0, s0, n0 = pb.cv_grid(X, y, g["fold_id"], 5, a, l, prior="binomial", n_devices=0)       # 0 = every visible device
    assert np.array_equal(e1, e2) and np.array_equal(s1, s2) and np.array_equal(n1, n2)
    assert np.array_equal(e1, e0)
    with pytest.raises(pb.ParebenError):
        pb.cv_grid(X, y, g["fold_id"], 5, a, l, prior="binomial", n_devices=pb.device_count() + 1)


def test_cross_validate_end_to_end_single_upload(pb, bundled):
    """CrossValidate through the public mirror: BuildGrid's lambda_max and the fits share one resident problem; the result
    equals the golden table made from the reference's C (config 1)."""
    g = golden("config1_gaussian.npz")
    X, y = bundled["BASIS"][:50, :100].astype(float), bundled["y"][:50]
    out = pb.CrossValidate(X, y, 3)
    assert out["alpha.optimal"] == float(g["alpha_optimal"])
    assert abs(out["lambda.optimal"] - float(g["lambda_optimal"])) <= 1e-12 * float(g["lambda_optimal"])   # lambda_max is summed on the device
    assert np.allclose(out["Results.Summary"]["MSE"], g["summary_mse"], rtol=RTOL, atol=0)
    assert np.allclose(out["Results.Summary"]["lambda"], np.sort(np.unique(g["grid_lambda"]))[np.tile(np.arange(20), 20)], rtol=1e-12, atol=0)
    if pb.device_count() >= 2:
        out2 = pb.CrossValidate(X, y, 3, n_devices=2)
        assert np.array_equal(out2["Results.Detail"]["MSE"], out["Results.Detail"]["MSE"])
