"""Synthetic genotype designs of BASELINE config 5: i.i.d. {-1, 0, 1} loci with P = (0.25, 0.5, 0.25),
linkage inside blocks of 50 (locus j copies locus j-1 on 90 % of the individuals), 10 causal main effects and 10
causal pairs with N(0, 2^2) effects, Gaussian noise of sd 10 around 100; the binomial response thresholds the
standardised linear predictor.  Bench and scale scripts only -- nothing in the solver depends on it."""
from __future__ import annotations

import numpy as np


def config5(n: int = 1000, k: int = 20000, seed: int = 20260101, block: int = 50, copy: float = 0.9, n_causal: int = 10):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.choice(np.array([-1, 0, 1], dtype=np.int8), size=(n, k), p=[0.25, 0.5, 0.25])
    keep = rng.random((n, k)) < copy
    for j in range(1, k):                       # sequential by construction: a copy of a copy is a copy
        if j % block:
            X[keep[:, j], j] = X[keep[:, j], j - 1]
    del keep
    main = rng.choice(k, n_causal, replace=False)
    pairs = rng.choice(k, (n_causal, 2), replace=False)
    b_main = rng.normal(0, 2.0, n_causal)
    b_pair = rng.normal(0, 2.0, n_causal)
    eta = X[:, main].astype(np.float64) @ b_main
    for (a, b), w in zip(pairs, b_pair):
        eta = eta + w * (X[:, a].astype(np.float64) * X[:, b])
    y = 100 + eta + rng.normal(0, 10.0, n)
    z = (eta - eta.mean()) / eta.std() * 2
    yb = (rng.random(n) < 1 / (1 + np.exp(-z))).astype(np.float64)
    return {"X": X, "y": y, "y_binomial": yb, "main": main, "pairs": pairs, "b_main": b_main, "b_pair": b_pair}
