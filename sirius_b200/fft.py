"""Mirror of reference src/fft.rs over bn256 Fr (the only field of the cycle with a 2-adic subgroup).

    get_omega_or_inv(k, is_inverse)   src/fft.rs:12-23
    get_ifft_divisor(k)               src/fft.rs:25-27
    best_fft(a, omega, log_n)         src/fft.rs:61-115
    fft / ifft / coset_fft / coset_ifft   src/fft.rs:160-198

Arrays are uint64 [n,4] Montgomery limbs; functions transform IN PLACE like the Rust `&mut [F]` versions and
also return the array.  The field constants below are the halo2curves bn256::Fr associated constants the Rust
shim passes through the C ABI (ROOT_OF_UNITY is pinned by the reference's fft known-answer test,
src/fft.rs:242-251; ZETA is not pinned by any reference test).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FR_S = 28
FR_ROOT_OF_UNITY = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
FR_ROOT_OF_UNITY_INV = pow(FR_ROOT_OF_UNITY, -1, FR_MODULUS)
FR_TWO_INV = 0x183227397098D014DC2822DB40C0AC2E9419F4243CDCB848A1F0FAC9F8000001
FR_ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23


def fr_to_limbs(v: int) -> np.ndarray:
    m = (v % FR_MODULUS) * (1 << 256) % FR_MODULUS
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def get_omega_or_inv(k: int, is_inverse: bool) -> int:
    assert k <= FR_S, f"k={k} should no larger than F::S={FR_S}"
    w = FR_ROOT_OF_UNITY_INV if is_inverse else FR_ROOT_OF_UNITY
    for _ in range(k, FR_S):
        w = w * w % FR_MODULUS
    return w


def get_ifft_divisor(k: int) -> int:
    return pow(FR_TWO_INV, k, FR_MODULUS)


def _arr(a) -> np.ndarray:
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]):
        raise TypeError("expected a C-contiguous uint64 array of Montgomery limbs (in-place transform)")
    return a


def _log2_len(a: np.ndarray) -> int:
    n = a.size // 4
    assert n and (n & (n - 1)) == 0, "a.len().is_power_of_two()"
    return n.bit_length() - 1


def best_fft(a: np.ndarray, omega: int, log_n: int, scale: int | None = None) -> np.ndarray:
    a = _arr(a)
    assert a.size // 4 == 1 << log_n
    w = fr_to_limbs(omega)
    sc = fr_to_limbs(scale) if scale is not None else None
    rc = _lib.load().sb_ntt(
        _lib.FIELD_FR, a.ctypes.data_as(_lib.u64p), log_n, w.ctypes.data_as(_lib.u64p), sc.ctypes.data_as(_lib.u64p) if sc is not None else None
    )
    _lib.check(rc)
    return a


def fft(a: np.ndarray) -> np.ndarray:
    k = _log2_len(a)
    return best_fft(a, get_omega_or_inv(k, False), k)


def ifft(a: np.ndarray) -> np.ndarray:
    k = _log2_len(a)
    return best_fft(a, get_omega_or_inv(k, True), k, scale=get_ifft_divisor(k))


def _coset(a: np.ndarray, z: int, z2: int) -> np.ndarray:
    a = _arr(a)
    zl, z2l = fr_to_limbs(z), fr_to_limbs(z2)
    rc = _lib.load().sb_coset_scale(_lib.FIELD_FR, a.ctypes.data_as(_lib.u64p), a.size // 4, zl.ctypes.data_as(_lib.u64p), z2l.ctypes.data_as(_lib.u64p))
    _lib.check(rc)
    return a


def coset_fft(a: np.ndarray, zeta: int = FR_ZETA) -> np.ndarray:
    _coset(a, zeta, zeta * zeta % FR_MODULUS)
    return fft(a)


def coset_ifft(a: np.ndarray, zeta: int = FR_ZETA) -> np.ndarray:
    """Returns the coefficient array (the Rust version wraps it in UnivariatePoly)."""
    ifft(a)
    return _coset(a, zeta * zeta % FR_MODULUS, zeta)


def ntt_device(d_a: int, log_n: int, omega: int, scale: int | None = None, stream: int = 0) -> None:
    w = fr_to_limbs(omega)
    sc = fr_to_limbs(scale) if scale is not None else None
    rc = _lib.load().sb_ntt_device(
        _lib.FIELD_FR, ctypes.c_void_p(d_a), log_n, w.ctypes.data_as(_lib.u64p), sc.ctypes.data_as(_lib.u64p) if sc is not None else None, ctypes.c_void_p(stream or None)
    )
    _lib.check(rc)
