"""Curve constants of the bn256 / grumpkin cycle the host side needs (halo2curves values)."""
from __future__ import annotations

import numpy as np

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
CURVE_BN256, CURVE_GRUMPKIN = 0, 1
BASE_FIELD = {CURVE_BN256: FQ, CURVE_GRUMPKIN: FR}
SCALAR_FIELD = {CURVE_BN256: FR, CURVE_GRUMPKIN: FQ}
FIELD_ID_OF_SCALAR = {CURVE_BN256: 0, CURVE_GRUMPKIN: 1}  # SB_FIELD_FR / SB_FIELD_FQ


def _sqrt(a: int, p: int) -> int:
    a %= p
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


def generator(curve: int):
    """bn256 G1: (1, 2) on y^2 = x^3 + 3; grumpkin: (1, sqrt(-16)) on y^2 = x^3 - 17 (smaller root)."""
    if curve == CURVE_BN256:
        return (1, 2)
    y = _sqrt((1 - 17) % FR, FR)
    return (1, min(y, FR - y))


def generator_limbs(curve: int) -> np.ndarray:
    p = BASE_FIELD[curve]
    out = np.zeros(8, dtype=np.uint64)
    for k, v in enumerate(generator(curve)):
        m = v * (1 << 256) % p
        for i in range(4):
            out[4 * k + i] = (m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF
    return out
