"""R's default random number stream, restated for the tests (test infrastructure): Mersenne-Twister
seeded the way set.seed() does it (R: src/main/RNG.c, RNG_Init: 50 rounds of the LCG 69069 x + 1, then
one round per state word; mti = 624 forces a fresh block) and normals by inversion
(snorm.c INVERSION: u = unif_rand(); u = (int)(2^27 u) + unif_rand(); qnorm(u / 2^27)).

Why: the reference's own test tests/testthat/test_approx.R:38-53 pins the Gaussian approximation of an
svm model whose data are `set.seed(123); rnorm(5)`.  With this file those inputs can be regenerated
without R, so that golden vector pins the oracle (tests/test_oracle.py).  qnorm here is
scipy.special.ndtri (Cephes) rather than R's AS241: they agree to ~1e-16, far below the 1e-9 the
golden comparison uses.
"""
import numpy as np
from scipy.special import ndtri


class RStream:
    def __init__(self, seed):
        s = np.uint32(seed)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(69069) * s + np.uint32(1)
            key = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                key[j] = s
        self.bg = np.random.MT19937()
        st = self.bg.state
        st["state"]["key"] = key[1:].copy()  # dummy[0] is mti, the 624 state words follow
        st["state"]["pos"] = 624             # FixupSeeds: mti = N, regenerate on first use
        self.bg.state = st

    def unif_rand(self):
        x = float(self.bg.random_raw()) * 2.3283064365386963e-10  # [0, 1)
        if x <= 0.0:
            return 0.5 * 2.328306437080797e-10
        if 1.0 - x <= 0.0:
            return 1.0 - 0.5 * 2.328306437080797e-10
        return x

    def norm_rand(self):
        big = 134217728.0  # 2^27
        u = self.unif_rand()
        u = float(int(big * u)) + self.unif_rand()
        return float(ndtri(u / big))

    def rnorm(self, n, mean=0.0, sd=1.0):
        return np.array([mean + sd * self.norm_rand() for _ in range(n)])

    def runif(self, n):
        return np.array([self.unif_rand() for _ in range(n)])
