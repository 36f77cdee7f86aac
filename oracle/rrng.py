"""Test infrastructure: R's default RNG (Mersenne-Twister + set.seed scrambling) and sample().

AssignToFolds (/root/reference/R/AssignToFolds.R:9-16) does `set.seed(1)` followed by
`sample(labels, N)`, so fold membership -- and with it every CV number -- is a function of R's
RNG.  This module restates, from R's published algorithm (src/main/RNG.c, src/main/random.c,
src/main/unique.c in the R sources; R itself is not present in this container):
  * set.seed(seed): 50 warm-up steps of the LCG 69069*s+1, then 625 more to fill
    (mti, mt[0..623]); mti is then forced to 624.
  * unif_rand(): MT19937 tempering * 2.3283064365386963e-10, clamped into the open (0,1).
  * sample.int(n, k) without replacement: partial Fisher-Yates  j=index(n); y=x[j]; x[j]=x[--n]
    with index(n) = floor(n*u)              for sample.kind="Rounding"  (R < 3.6.0)
         index(n) = rejection-sampled bits  for sample.kind="Rejection" (R >= 3.6.0, default).
Known answers (from R's documentation of the 3.6.0 change, re-checked by the survey):
  set.seed(1); sample(10) == 9 4 7 1 2 5 3 10 6 8 (Rejection) / 3 4 5 7 2 8 9 6 10 1 (Rounding).
"""
from __future__ import annotations

import math

_N, _M = 624, 397
_UPPER, _LOWER = 0x80000000, 0x7FFFFFFF
_I2_32M1 = 2.328306437080797e-10


class RRng:
    def __init__(self, seed: int = 1, sample_kind: str = "Rejection"):
        if sample_kind not in ("Rejection", "Rounding"):
            raise ValueError(sample_kind)
        self.sample_kind = sample_kind
        self.set_seed(seed)

    def set_seed(self, seed: int) -> None:
        s = seed & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        state = []
        for _ in range(_N + 1):
            s = (69069 * s + 1) & 0xFFFFFFFF
            state.append(s)
        self.mt = state[1:]
        self.mti = _N          # FixupSeeds: dummy[0] = 624

    def _genrand(self) -> int:
        mt = self.mt
        if self.mti >= _N:
            for kk in range(_N):
                y = (mt[kk] & _UPPER) | (mt[(kk + 1) % _N] & _LOWER)
                mt[kk] = mt[(kk + _M) % _N] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def unif_rand(self) -> float:
        v = self._genrand() * 2.3283064365386963e-10
        if v <= 0.0:
            return 0.5 * _I2_32M1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * _I2_32M1
        return v

    def _rbits(self, bits: int) -> float:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return float(v)

    def unif_index(self, dn: float) -> int:
        if self.sample_kind == "Rounding":
            return int(math.floor(dn * self.unif_rand()))
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dn > dv:
                return int(dv)

    def sample_int(self, n: int, k: int | None = None) -> list[int]:
        """sample.int(n, k), replace = FALSE; 1-based results."""
        k = n if k is None else k
        x = list(range(n))
        out = []
        for _ in range(k):
            j = self.unif_index(float(n))
            out.append(x[j] + 1)
            n -= 1
            x[j] = x[n]
        return out

    def sample(self, values, k: int | None = None) -> list:
        values = list(values)
        return [values[i - 1] for i in self.sample_int(len(values), k)]
