"""Test infrastructure: CPU restatement of parEBEN's R layer + EBEN's R wrappers (numpy).

This is the ORACLE for everything in the hot path that the reference implements in R:
  BuildGrid / GetLambdaMax   /root/reference/R/BuildGrid.R:5-52
  AssignToFolds              /root/reference/R/AssignToFolds.R:6-19
  TestModel                  /root/reference/R/TestModel.R:6-39
  GetFoldError               /root/reference/R/GetModelError.R:6-59
  CrossValidate (global)     /root/reference/R/CrossValidate.R:61-117
  LocalSearch                /root/reference/R/LocalSearch.R:6-130
  EBelasticNet.Gaussian      /root/reference/EBEN_orig/R/EBelasticNet.Gaussian.R:1-101
  EBelasticNet.Binomial      /root/reference/EBEN_orig/R/EBelasticNet.Binomial.R:1-85
The per-fit arithmetic (EBEN's C) is delegated to a shared library exposing the reference's
four `.C` entry points: oracle/_ref/libeben_ref.so (the reference's own C, compiled
unmodified) or oracle/libeben_oracle.so (this repo's C restatement, symbols prefixed
`oracle_`).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may use it;
the product (pareben_b200/) never imports this package.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass

import numpy as np

from .rrng import RRng

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libeben_ref.so")
PORT_LIB = os.path.join(_HERE, "libeben_oracle.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class FitLib:
    """ctypes view of the four `.C` symbols (EBelasticNet.Gaussian.R:16-51, .Binomial.R:12-46)."""

    def __init__(self, kind: str = "reference"):
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        if kind == "reference":
            path, prefix = REF_LIB, ""
        elif kind == "port":
            path, prefix = PORT_LIB, "oracle_"
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built: run `make -C oracle`")
        self.kind = kind
        self.lib = ctypes.CDLL(path)
        self.g_main = getattr(self.lib, prefix + "elasticNetLinearNeMainEff")
        self.g_epis = getattr(self.lib, prefix + "elasticNetLinearNeEpisEff", None) if kind == "port" \
            else self.lib.elasticNetLinearNeEpisEff
        self.b_main = getattr(self.lib, prefix + "ElasticNetBinaryNEmainEff", None) if kind == "port" \
            else self.lib.ElasticNetBinaryNEmainEff
        self.b_epis = getattr(self.lib, prefix + "ElasticNetBinaryNEfull", None) if kind == "port" \
            else self.lib.ElasticNetBinaryNEfull
        for f in (self.g_main, self.g_epis):
            if f is not None:
                f.restype = None
                f.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _ip, _dp]
        for f in (self.b_main, self.b_epis):
            if f is not None:
                f.restype = None
                f.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _ip, _ip]


_LIBS: dict[str, FitLib] = {}


def fit_lib(kind: str = "reference") -> FitLib:
    if kind not in _LIBS:
        _LIBS[kind] = FitLib(kind)
    return _LIBS[kind]


def available_kind() -> str:
    """'reference' if the compiled reference is present, else the restatement."""
    return "reference" if os.path.exists(REF_LIB) else "port"


@dataclass
class Fit:
    weight: np.ndarray        # M x 6: loc1, loc2, beta, var, t, p (p left as nan: needs pt())
    intercept: np.ndarray     # [mu0] (gaussian) / [mu0, var0] (binomial)
    wald: float
    resid_var: float = float("nan")      # gaussian only
    log_likelihood: float = float("nan")  # binomial only
    raw_beta: np.ndarray | None = None


def _as_c(BASIS, Target):
    X = np.asfortranarray(np.asarray(BASIS, dtype=np.float64))   # as.double(matrix): column-major
    y = np.ascontiguousarray(np.asarray(Target, dtype=np.float64).ravel())
    return X, y


def _blup_table(result: np.ndarray, keep: np.ndarray, epis: bool, n: int) -> np.ndarray:
    ncol = result.shape[1]
    if keep.size == 0:
        blup = np.zeros((1, ncol))
    else:
        blup = result[keep, :]
    if epis:
        main = blup[blup[:, 0] == blup[:, 1], :]
        ep = blup[blup[:, 0] != blup[:, 1], :]
        main = main[np.argsort(main[:, 0], kind="stable"), :]
        ep = ep[np.argsort(ep[:, 0], kind="stable"), :]
        blup = np.vstack([main, ep])
    blup = blup[:, :4]
    t = np.abs(blup[:, 2]) / (np.sqrt(blup[:, 3]) + 1e-20)
    p = np.full_like(t, np.nan)
    try:
        from scipy.stats import t as student
        p = 2.0 * (1.0 - student.cdf(t, df=n - 1))
    except Exception:  # pragma: no cover
        pass
    return np.column_stack([blup, t, p])


def eb_elastic_net_gaussian(BASIS, Target, lam, alpha, epis: bool = False, lib: FitLib | None = None) -> Fit:
    """EBelasticNet.Gaussian (EBelasticNet.Gaussian.R:1-101)."""
    lib = lib or fit_lib(available_kind())
    X, y = _as_c(BASIS, Target)
    n, k = X.shape
    n_eff = (k + 1) * k // 2 if epis else k
    ncol = 5 if epis else 4
    beta = np.zeros(n_eff * ncol)
    wald = np.zeros(1); icpt = np.zeros(1); resid = np.zeros(1)
    fn = lib.g_epis if epis else lib.g_main
    fn(_ptr(X), _ptr(y), _ptr(np.array([float(lam)])), _ptr(np.array([float(alpha)])), _ptr(beta),
       _ptr(wald), _ptr(icpt), ctypes.byref(ctypes.c_int(n)), ctypes.byref(ctypes.c_int(k)),
       ctypes.byref(ctypes.c_int(0)), _ptr(resid))
    result = beta.reshape((n_eff, ncol), order="F")
    keep = np.nonzero(result[:, 4] != 0)[0] if epis else np.nonzero(result[:, 2] != 0)[0]
    return Fit(_blup_table(result, keep, epis, n), icpt.copy(), float(wald[0]), resid_var=float(resid[0]),
               raw_beta=result)


def eb_elastic_net_binomial(BASIS, Target, lam, alpha, epis: bool = False, lib: FitLib | None = None) -> Fit:
    """EBelasticNet.Binomial (EBelasticNet.Binomial.R:1-85)."""
    lib = lib or fit_lib(available_kind())
    X, y = _as_c(BASIS, Target)
    n, k = X.shape
    n_eff = 2 * k if epis else k
    beta = np.zeros(n_eff * 4)
    wald = np.zeros(1); icpt = np.zeros(2); logl = np.zeros(1)
    fn = lib.b_epis if epis else lib.b_main
    fn(_ptr(X), _ptr(y), _ptr(np.array([float(lam)])), _ptr(np.array([float(alpha)])), _ptr(logl), _ptr(beta),
       _ptr(wald), _ptr(icpt), ctypes.byref(ctypes.c_int(n)), ctypes.byref(ctypes.c_int(k)),
       ctypes.byref(ctypes.c_int(0)), ctypes.byref(ctypes.c_int(n_eff)))
    result = beta.reshape((n_eff, 4), order="F")
    keep = np.nonzero(result[:, 2] != 0)[0]
    return Fit(_blup_table(result, keep, epis, n), icpt.copy(), float(wald[0]), log_likelihood=float(logl[0]),
               raw_beta=result)


# ----------------------------------------------------------------------------------------------
# R numerics helpers
def _r_sum(x) -> np.longdouble:
    return np.sum(np.asarray(x, dtype=np.longdouble))


def r_mean(x) -> float:
    """R's mean(): long-double sum / n, then one refinement pass (summary.c)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    n = x.size
    s = _r_sum(x) / n
    t = _r_sum(x.astype(np.longdouble) - s)
    return float(s + t / n)


def r_sd(x) -> float:
    """sd(): sqrt(var), var = two-pass (cov.c) with n-1 denominator."""
    x = np.asarray(x, dtype=np.float64).ravel()
    n = x.size
    if n < 2:
        return float("nan")
    m = np.longdouble(r_mean(x))
    d = x.astype(np.longdouble) - m
    return float(np.sqrt(float(np.sum(d * d) / (n - 1))))


def r_seq(frm: float, to: float, by: float) -> np.ndarray:
    """seq.default(from, to, by) for doubles."""
    n = int(math.floor((to - frm) / by + 1e-10))
    x = frm + np.arange(n + 1, dtype=np.float64) * by
    return np.minimum(x, to) if by > 0 else np.maximum(x, to)


# ----------------------------------------------------------------------------------------------
def get_lambda_max(BASIS, Target, epis: bool = False) -> float:
    """GetLambdaMax (BuildGrid.R:5-32).  Note the Epis pairs use the UN-normalised response (:26)."""
    X = np.asarray(BASIS, dtype=np.float64)
    y = np.asarray(Target, dtype=np.float64).ravel()
    lam = math.log(1.1)
    centred = y - r_mean(y)
    response = centred / math.sqrt(float(_r_sum(centred * centred)))
    norms = np.sqrt(np.array([float(_r_sum(X[:, j] * X[:, j])) for j in range(X.shape[1])]))
    with np.errstate(divide="ignore", invalid="ignore"):
        cor = (X / norms).T @ response
    for c in cor:                      # NaN (all-zero column) never compares greater, as in R... R would
        if c > lam:                    # actually raise on `if (NA)`; bundled/synthetic data have no zero column.
            lam = float(c)
    if epis:
        k = X.shape[1]
        for i in range(k - 1):
            P = X[:, i:i + 1] * X[:, i + 1:]
            nrm = np.sqrt(np.sum(P * P, axis=0))
            with np.errstate(divide="ignore", invalid="ignore"):
                c = (P / nrm).T @ centred
            c = c[~np.isnan(c)]
            if c.size and c.max() > lam:
                lam = float(c.max())
    return lam


def build_grid(BASIS, Target, n_folds: int = 0, epis: bool = False):
    """BuildGrid (BuildGrid.R:34-52): returns (alpha[400], lambda[400]) in expand.grid order (alpha fastest)."""
    lam_max = get_lambda_max(BASIS, Target, epis) * 10
    lam_min = math.log(0.001 * lam_max)
    step = (math.log(lam_max) - lam_min) / 19
    Lambda = np.exp(r_seq(math.log(lam_max), lam_min, -step))
    Alpha = r_seq(1.0, 0.05, -0.05)
    grid_alpha = np.tile(Alpha, Lambda.size)
    grid_lambda = np.repeat(Lambda, Alpha.size)
    return grid_alpha, grid_lambda


def assign_to_folds(n: int, n_folds: int, sample_kind: str = "Rejection", seed: int | None = 1,
                    rng: RRng | None = None) -> np.ndarray:
    """AssignToFolds (AssignToFolds.R:6-19): set.seed(1); sample(balanced labels, N).  1-based labels."""
    rng = rng or RRng(seed, sample_kind)
    labels = list(np.tile(np.arange(1, n_folds + 1), n // n_folds))
    if n % n_folds != 0:
        labels += list(range(1, n % n_folds + 1))
    return np.array(rng.sample(labels, n), dtype=np.int32)


def get_fold_error(basis_test, target_test, fit: Fit, prior: str = "gaussian") -> float:
    """GetFoldError (GetModelError.R:6-59)."""
    Xt = np.asarray(basis_test, dtype=np.float64)
    yt = np.asarray(target_test, dtype=np.float64).ravel()
    betas = fit.weight
    M = betas.shape[0]
    mu = betas[:, 2]
    mu0 = float(fit.intercept[0])
    ntest = Xt.shape[0]
    if prior == "gaussian":
        cols = np.zeros((ntest, M))
        empty = False
        for i in range(M):
            l1, l2 = int(betas[i, 0]), int(betas[i, 1])
            if l1 != 0:
                cols[:, i] = Xt[:, l1 - 1] if l1 == l2 else Xt[:, l1 - 1] * Xt[:, l2 - 1]
            else:
                empty = True        # basisTest <- rep(0, ntest): predictions collapse to Mu0 (:21-26)
        pred = np.full(ntest, mu0) if empty else mu0 + cols @ mu
        temp = yt - pred
        return float(temp @ temp)
    if M == 1 and betas[0, 0] == 0:
        return 0.0
    cols = np.zeros((ntest, M))
    for i in range(M):
        l1, l2 = int(betas[i, 0]), int(betas[i, 1])
        cols[:, i] = Xt[:, l1 - 1] if l1 == l2 else Xt[:, l1 - 1] * Xt[:, l2 - 1]
    with np.errstate(over="ignore"):
        temp = np.exp(mu0 + cols @ mu)
    if temp.max() > 1e10:
        temp[temp > 1e10] = 1e5
    if temp.min() < 1e-10:
        temp[temp < 1e-10] = 1e-5
    return r_mean(yt * np.log(temp / (1 + temp)) + (1 - yt) * np.log(1 / (1 + temp)))


def fit_one(BASIS, Target, fold_id, fold, lam, alpha, epis=False, prior="gaussian", lib=None):
    """One (fold, alpha, lambda) task of TestModel (TestModel.R:11-36) -> (error, Fit)."""
    X = np.asarray(BASIS); y = np.asarray(Target).ravel()
    tr = fold_id != fold
    te = fold_id == fold
    if prior == "gaussian":
        fit = eb_elastic_net_gaussian(X[tr], y[tr], lam, alpha, epis, lib)
    else:
        fit = eb_elastic_net_binomial(X[tr], y[tr], lam, alpha, epis, lib)
    return get_fold_error(X[te], y[te], fit, prior), fit


def summarise(grid_alpha, grid_lambda, fold_err, prior="gaussian"):
    """group_by(alpha, lambda) %>% summarise(...)  (CrossValidate.R:72-76, 93-97).

    fold_err is [n_grid, n_folds] in grid-row order.  Returns dict of columns sorted by alpha
    then lambda ascending.  SE = sd/sqrt(max(foldId)) with the *per-fold* values."""
    n_grid, n_folds = fold_err.shape
    order = np.lexsort((grid_lambda, grid_alpha))
    se = np.array([r_sd(fold_err[r]) / math.sqrt(n_folds) for r in order])
    mean = np.array([r_mean(fold_err[r]) for r in order])
    out = {"alpha": grid_alpha[order], "lambda": grid_lambda[order], "SE": se, "row": order}
    if prior == "gaussian":
        out["MSE"] = mean
    else:
        out["Likelihood"] = -mean
    return out


def select_optimum(summary, prior="gaussian"):
    """which.min over the summary (CrossValidate.R:78-80).  For the binomial prior the reference
    indexes a non-existent column (:99) and returns empty; the intended rule
    which.min(Likelihood) is used here and the deviation is documented in DESIGN.md."""
    col = summary["MSE"] if prior == "gaussian" else summary["Likelihood"]
    idx = int(np.argmin(col))          # first minimum, like which.min
    return float(summary["alpha"][idx]), float(summary["lambda"][idx]), idx


def cross_validate_global(BASIS, Target, n_folds, epis=False, prior="gaussian", lib=None,
                          sample_kind="Rejection", rows=None):
    """CrossValidate(search="global") (CrossValidate.R:62-108) driven fit by fit through `lib`.
    `rows` restricts the grid rows evaluated (others are left NaN) for sampled comparisons."""
    X = np.asarray(BASIS); y = np.asarray(Target).ravel()
    ga, gl = build_grid(X, y, n_folds, epis)
    fold_id = assign_to_folds(X.shape[0], n_folds, sample_kind)
    err = np.full((ga.size, n_folds), np.nan)
    nsel = np.zeros((ga.size, n_folds), dtype=np.int32)
    for r in (range(ga.size) if rows is None else rows):
        for f in range(1, n_folds + 1):
            e, fit = fit_one(X, y, fold_id, f, gl[r], ga[r], epis, prior, lib)
            err[r, f - 1] = e
            nsel[r, f - 1] = 0 if fit.weight[0, 0] == 0 else fit.weight.shape[0]
    out = {"grid_alpha": ga, "grid_lambda": gl, "fold_id": fold_id, "fold_err": err, "n_selected": nsel}
    if rows is None:
        s = summarise(ga, gl, err, prior)
        a, l, idx = select_optimum(s, prior)
        out.update(summary=s, alpha_optimal=a, lambda_optimal=l)
    return out


def local_search_replay(grid_alpha, grid_lambda, fold_err):
    """LocalSearch's early-stopping walk (LocalSearch.R:51-124) replayed over a full table of
    per-fold SSEs (fits are independent of visiting order).
    Returns (CrossValidation[20x4], alpha_opt, lambda_opt, fullCV[400x4])."""
    Alpha = grid_alpha[:20]
    Lambda = grid_lambda[::20]
    n_alpha, n_step = Alpha.size, Lambda.size
    n_folds = fold_err.shape[1]
    mse_cv = np.zeros((n_step * n_alpha, 4))
    each = np.zeros((n_alpha, 4))
    step = 0
    for ia in range(n_alpha):
        sse = np.full((n_step, 2), 1e10)
        for i_s in range(n_step):
            # which.min(SSE1Alpha[1:(i_s-1),1]); for i_s==1 R's 1:0 == c(1,0) -> row 1
            upto = max(i_s, 1) if i_s != 0 else 1
            mi = int(np.argmin(sse[:upto, 0]))
            previous = sse[mi, 0] + sse[mi, 1]
            row = i_s * 20 + ia
            fe = fold_err[row]
            m, se = r_mean(fe), r_sd(fe) / math.sqrt(n_folds)
            sse[i_s] = (m, se)
            mse_cv[step] = (Alpha[ia], Lambda[i_s], m, se)
            step += 1
            if m - previous > 0:
                break
        idx = int(np.argmin(sse[:, 0]))
        each[ia] = (Alpha[ia], Lambda[idx], sse[idx, 0], sse[idx, 1])
    idx = int(np.argmin(each[:, 2]))
    return each, float(each[idx, 0]), float(each[idx, 1]), mse_cv


def sl_filter(BASIS, Target, tau_main: float = 0.02, tau_pair: float = 0.05, epis: bool = True):
    """SL_filter.R (/root/reference/paper_materials/Real Data Analysis/SL_filter.R:17-52): `scale()` the response and
    the genotype columns (centre, n-1 standard deviation), keep main effects with |ys' xs| / n > tau_main (:21); for
    k = 1..K-1 form the products of column k with columns k+1..K, scale them, keep those above tau_pair (:28-39).
    Returns (main 1-based columns, pairs as 1-based (i, j) with i < j in the script's visiting order, stats)."""
    X = np.asarray(BASIS, dtype=np.float64)
    y = np.asarray(Target, dtype=np.float64).ravel()
    n, k = X.shape

    def scale(a):
        a = a - a.mean(axis=0)
        with np.errstate(invalid="ignore", divide="ignore"):
            return a / np.sqrt((a * a).sum(axis=0) / (n - 1))

    ys = scale(y[:, None])[:, 0]
    with np.errstate(invalid="ignore"):
        sm = np.abs(ys @ scale(X)) / n
    main = np.nonzero(sm > tau_main)[0]
    pairs, sp = [], []
    if epis:
        for a in range(k - 1):
            prod = X[:, a + 1:] * X[:, [a]]
            with np.errstate(invalid="ignore"):
                st = np.abs(ys @ scale(prod)) / n
            hit = np.nonzero(st > tau_pair)[0]
            pairs += [(a + 1, a + 2 + int(b)) for b in hit]
            sp += [float(st[b]) for b in hit]
    return main + 1, np.array(pairs, dtype=np.int64).reshape(-1, 2), sm[main], np.array(sp)
