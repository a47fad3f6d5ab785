"""Pure-Python big-int restatement of the Sirius hot-path arithmetic.

TEST INFRASTRUCTURE ONLY. Nothing under ``sirius_b200/`` may import this module; it exists so
that ``tests/`` can pin the C oracle (``oracle/sirius_oracle.c``) and the CUDA path against
something whose correctness is visible by inspection (Python ints, affine group law, O(n^2) DFT).

Reference anchors (paths relative to the reference tree):
  * field/curve facts: halo2curves bn256 / grumpkin as re-exported at src/lib.rs:24-27
    (Fr = bn256 scalar = grumpkin base, Fq = bn256 base = grumpkin scalar, both a = 0,
    bn256 b = 3, grumpkin b = -17), 4x64-bit little-endian Montgomery limbs, R = 2^256.
  * fft semantics: src/fft.rs:12-27 (omega, ifft divisor), :61-115 (best_fft), :160-228 wrappers.
  * golden vectors: src/fft.rs:242-251, src/polynomial/lagrange.rs:118-126.
"""
from __future__ import annotations

import numpy as np

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R256 = 1 << 256

# field ids shared with the C oracle and the CUDA library
FIELD_FR = 0
FIELD_FQ = 1
# curve ids: the curve's BASE field is FQ for bn256 and FR for grumpkin
CURVE_BN256 = 0
CURVE_GRUMPKIN = 1

MODULUS = {FIELD_FR: FR, FIELD_FQ: FQ}
CURVE_BASE = {CURVE_BN256: FQ, CURVE_GRUMPKIN: FR}
CURVE_SCALAR = {CURVE_BN256: FR, CURVE_GRUMPKIN: FQ}
CURVE_B = {CURVE_BN256: 3, CURVE_GRUMPKIN: (-17) % FR}
CURVE_GEN = {CURVE_BN256: (1, 2), CURVE_GRUMPKIN: (1, 0)}  # grumpkin y filled below

FR_S = 28
FR_GENERATOR = 7
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (FR - 1) >> FR_S, FR)
FR_ROOT_OF_UNITY_INV = pow(FR_ROOT_OF_UNITY, -1, FR)
FR_TWO_INV = pow(2, -1, FR)
# halo2curves bn256::Fr::ZETA (cube root of unity).  Not pinned by any reference test:
# the C ABI always takes zeta from the caller.
FR_ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23


def _sqrt_mod(a: int, p: int) -> int:
    """Tonelli-Shanks."""
    a %= p
    if a == 0:
        return 0
    assert pow(a, (p - 1) // 2, p) == 1, "not a square"
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return r


# grumpkin generator: halo2curves uses (1, sqrt(1 - 17)); either root generates the (prime order) group.
_gy = _sqrt_mod((1 + CURVE_B[CURVE_GRUMPKIN]) % FR, FR)
CURVE_GEN[CURVE_GRUMPKIN] = (1, min(_gy, FR - _gy))

# ------------------------------------------------------------------ limb <-> int conversions


def to_mont_limbs(vals, modulus: int) -> np.ndarray:
    """ints (canonical) -> uint64 [n,4] array of Montgomery-form little-endian limbs."""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(vals):
        m = (v % modulus) * R256 % modulus
        out[i, 0] = m & mask
        out[i, 1] = (m >> 64) & mask
        out[i, 2] = (m >> 128) & mask
        out[i, 3] = (m >> 192) & mask
    return out


def from_mont_limbs(arr: np.ndarray, modulus: int) -> list[int]:
    """uint64 [...,4] Montgomery limbs -> canonical Python ints."""
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(R256, -1, modulus)
    out = []
    for row in a:
        m = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
        out.append(m * rinv % modulus)
    return out


def points_to_limbs(points, curve: int) -> np.ndarray:
    """[(x,y) | None] -> uint64 [n,8] (x limbs, y limbs), identity encoded as (0,0)
    (src/commitment.rs:43-45)."""
    p = CURVE_BASE[curve]
    flat = []
    for pt in points:
        if pt is None:
            flat += [0, 0]
        else:
            flat += [pt[0], pt[1]]
    return to_mont_limbs(flat, p).reshape(len(points), 8)


def limbs_to_points(arr: np.ndarray, curve: int):
    p = CURVE_BASE[curve]
    vals = from_mont_limbs(np.asarray(arr).reshape(-1, 4), p)
    pts = []
    for i in range(0, len(vals), 2):
        x, y = vals[i], vals[i + 1]
        pts.append(None if (x == 0 and y == 0) else (x, y))
    return pts


# ------------------------------------------------------------------ affine group law (a = 0)


def ec_add(P, Q, curve: int):
    p = CURVE_BASE[curve]
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    y3 = (lam * (x1 - x3) - y1) % p
    return (x3, y3)


def ec_neg(P, curve: int):
    if P is None:
        return None
    return (P[0], (-P[1]) % CURVE_BASE[curve])


def ec_mul(k: int, P, curve: int):
    k %= CURVE_SCALAR[curve]
    acc = None
    add = P
    while k:
        if k & 1:
            acc = ec_add(acc, add, curve)
        add = ec_add(add, add, curve)
        k >>= 1
    return acc


def is_on_curve(P, curve: int) -> bool:
    if P is None:
        return True
    p = CURVE_BASE[curve]
    x, y = P
    return (y * y - x * x * x - CURVE_B[curve]) % p == 0


def msm_naive(scalars, points, curve: int):
    """sum_i s_i * P_i by double-and-add -- the mathematical definition that
    src/commitment.rs:81-90 (`best_multiexp(...).to_affine()`) must equal."""
    acc = None
    for s, P in zip(scalars, points):
        acc = ec_add(acc, ec_mul(s, P, curve), curve)
    return acc


def running_bases(n: int, curve: int):
    """G_i = [i+1]G by running addition (BASELINE.md section 3 synthetic bases)."""
    G = CURVE_GEN[curve]
    out, cur = [], None
    for _ in range(n):
        cur = ec_add(cur, G, curve)
        out.append(cur)
    return out


# ------------------------------------------------------------------ FFT (definition-level)


def omega_for(k: int, inverse: bool = False) -> int:
    """src/fft.rs:12-23: ROOT_OF_UNITY (or its inverse) squared (S-k) times."""
    assert k <= FR_S
    w = FR_ROOT_OF_UNITY_INV if inverse else FR_ROOT_OF_UNITY
    for _ in range(k, FR_S):
        w = w * w % FR
    return w


def dft_naive(a, omega: int, modulus: int = FR):
    """out[i] = sum_j a[j] * omega^(i*j): what best_fft (src/fft.rs:61-115) computes, natural order."""
    n = len(a)
    out = []
    for i in range(n):
        wi = pow(omega, i, modulus)
        acc, w = 0, 1
        for j in range(n):
            acc = (acc + a[j] * w) % modulus
            w = w * wi % modulus
        out.append(acc)
    return out


def fft(a):
    k = len(a).bit_length() - 1
    return dft_naive(a, omega_for(k, False))


def ifft(a):
    k = len(a).bit_length() - 1
    div = pow(FR_TWO_INV, k, FR)  # src/fft.rs:25-27
    return [x * div % FR for x in dft_naive(a, omega_for(k, True))]


def distribute_powers_zeta(a, zeta: int, into_coset: bool):
    """src/fft.rs:207-228."""
    zinv = zeta * zeta % FR
    powers = [zeta, zinv] if into_coset else [zinv, zeta]
    out = list(a)
    for idx in range(len(out)):
        i = idx % 3
        if i:
            out[idx] = out[idx] * powers[i - 1] % FR
    return out


def coset_fft(a, zeta: int = FR_ZETA):
    return fft(distribute_powers_zeta(a, zeta, True))


def coset_ifft(a, zeta: int = FR_ZETA):
    return distribute_powers_zeta(ifft(a), zeta, False)


# ------------------------------------------------------------------ Lagrange helpers (src/polynomial/lagrange.rs)


def iter_cyclic_subgroup(log_n: int):
    """lagrange.rs:22-26: 1, w, w^2, ... with w of order 2^log_n."""
    w = omega_for(log_n, False)
    cur = 1
    for _ in range(1 << log_n):
        yield cur
        cur = cur * w % FR


def eval_lagrange_polys(log_n: int, X: int):
    """lagrange.rs:50-74: L_i(X) over the 2^log_n cyclic subgroup, with the 0/0 case -> 1 (:67-68)."""
    n = 1 << log_n
    ninv = pow(n, -1, FR)
    xn1 = (pow(X, n, FR) - 1) % FR
    out = []
    for wi in iter_cyclic_subgroup(log_n):
        den = (X - wi) % FR
        if den == 0:
            out.append(1)
        else:
            out.append(wi * ninv % FR * xn1 % FR * pow(den, -1, FR) % FR)
    return out


# ------------------------------------------------------------------ deterministic synthetic inputs


class Xoshiro256ss:
    """xoshiro256** (SURVEY section 8d generator), seeded through splitmix64."""

    MASK = (1 << 64) - 1

    def __init__(self, seed: int):
        s = []
        x = seed & self.MASK
        for _ in range(4):
            x = (x + 0x9E3779B97F4A7C15) & self.MASK
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.MASK
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.MASK
            s.append(z ^ (z >> 31))
        self.s = s

    @staticmethod
    def _rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & Xoshiro256ss.MASK

    def next(self) -> int:
        s = self.s
        result = (self._rotl((s[1] * 5) & self.MASK, 7) * 9) & self.MASK
        t = (s[1] << 17) & self.MASK
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 45)
        return result

    def field(self, modulus: int) -> int:
        while True:
            v = self.next() | (self.next() << 64) | (self.next() << 128) | (self.next() << 192)
            v &= (1 << 254) - 1
            if v < modulus:
                return v
