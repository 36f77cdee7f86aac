"""R semantics the host layer must reproduce without R: set.seed()/sample() (Mersenne-Twister,
both sample.kind variants), seq(), mean(), sd().  Needed because AssignToFolds
(/root/reference/R/AssignToFolds.R:9-16) seeds R's RNG inside every task, so fold membership
is part of the reference's observable behaviour."""
from __future__ import annotations

import math

import numpy as np


class RRandom:
    """R's default generator after set.seed(seed)."""

    def __init__(self, seed: int = 1, sample_kind: str = "Rejection"):
        if sample_kind not in ("Rejection", "Rounding"):
            raise ValueError("sample_kind must be 'Rejection' (R >= 3.6) or 'Rounding' (R < 3.6)")
        self.kind = sample_kind
        x = seed & 0xFFFFFFFF
        for _ in range(50):                       # initial scrambling
            x = (69069 * x + 1) & 0xFFFFFFFF
        words = []
        for _ in range(625):                      # mti followed by the 624-word state
            x = (69069 * x + 1) & 0xFFFFFFFF
            words.append(x)
        self.state = words[1:]
        self.idx = 624                            # forces a reload on first use

    def _next32(self) -> int:
        st = self.state
        if self.idx >= 624:
            for i in range(624):
                y = (st[i] & 0x80000000) | (st[(i + 1) % 624] & 0x7FFFFFFF)
                st[i] = st[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF * (y & 1))
            self.idx = 0
        y = st[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def runif(self) -> float:
        u = self._next32() * 2.3283064365386963e-10
        lo = 2.328306437080797e-10 / 2
        return lo if u <= 0.0 else (1.0 - lo if 1.0 - u <= 0.0 else u)

    def _index(self, n: int) -> int:
        if self.kind == "Rounding":
            return int(n * self.runif())
        bits = int(math.ceil(math.log2(n))) if n > 0 else 0
        while True:
            v, got = 0, 0
            while got <= bits:
                v = (v << 16) + int(self.runif() * 65536)
                got += 16
            v &= (1 << bits) - 1
            if v < n:
                return v

    def sample(self, values) -> list:
        """sample(values, length(values)) without replacement."""
        pool = list(range(len(values)))
        n = len(pool)
        picked = []
        while len(picked) < len(values):
            j = self._index(n)
            picked.append(values[pool[j]])
            n -= 1
            pool[j] = pool[n]
        return picked


def seq(frm: float, to: float, by: float) -> np.ndarray:
    n = int((to - frm) / by + 1e-10)
    x = frm + np.arange(n + 1) * by
    return np.minimum(x, to) if by > 0 else np.maximum(x, to)


def mean(x) -> float:
    x = np.asarray(x, dtype=np.longdouble).ravel()
    m = x.sum() / x.size
    return float(m + (x - m).sum() / x.size)


def sd(x) -> float:
    x = np.asarray(x, dtype=np.longdouble).ravel()
    if x.size < 2:
        return float("nan")
    m = np.longdouble(mean(x))
    return float(np.sqrt(float(((x - m) ** 2).sum() / (x.size - 1))))
