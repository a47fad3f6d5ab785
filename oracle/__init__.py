"""ctypes front-end for the CPU parity oracle (oracle/sirius_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (sirius_b200/) never imports this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import pyref  # noqa: F401  (re-export for tests)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsirius_oracle.so")
_lib = None

_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, OpenMP).  Building the checker is not using it."""
    src = os.path.join(_HERE, "sirius_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []) + ["all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def num_threads() -> int:
    return int(lib().so_num_threads())


def msm(curve: int, scalars: np.ndarray, bases: np.ndarray, threads: int = 0) -> np.ndarray:
    """CommitmentKey::commit restated (src/commitment.rs:81-90): returns uint64[8] affine (x,y)."""
    s, b = _c(scalars).reshape(-1, 4), _c(bases).reshape(-1, 8)
    n = s.shape[0]
    assert b.shape[0] >= n
    out = np.zeros(8, dtype=np.uint64)
    rc = lib().so_msm(ctypes.c_int(curve), _p(s), _p(b), ctypes.c_size_t(n), ctypes.c_int(threads), _p(out))
    assert rc == 0
    return out


def msm_naive(curve: int, scalars: np.ndarray, bases: np.ndarray) -> np.ndarray:
    s, b = _c(scalars).reshape(-1, 4), _c(bases).reshape(-1, 8)
    out = np.zeros(8, dtype=np.uint64)
    rc = lib().so_msm_naive(ctypes.c_int(curve), _p(s), _p(b), ctypes.c_size_t(s.shape[0]), _p(out))
    assert rc == 0
    return out


def point_add(curve: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    lib().so_point_add(ctypes.c_int(curve), _p(_c(a).reshape(8)), _p(_c(b).reshape(8)), _p(out))
    return out


def is_on_curve(curve: int, xy: np.ndarray) -> bool:
    return bool(lib().so_is_on_curve(ctypes.c_int(curve), _p(_c(xy).reshape(8))))


def running_bases(curve: int, n: int) -> np.ndarray:
    """G_i = [i+1]G, uint64 [n,8] Montgomery affine."""
    gen = pyref.points_to_limbs([pyref.CURVE_GEN[curve]], curve).reshape(8)
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().so_running_bases(ctypes.c_int(curve), _p(gen), ctypes.c_size_t(n), _p(out))
    return out


def random_field(field: int, seed: int, n: int) -> np.ndarray:
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().so_random_field(ctypes.c_int(field), ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_size_t(n), _p(out))
    return out


def best_fft(field: int, a: np.ndarray, omega: np.ndarray, threads: int = 1) -> np.ndarray:
    """src/fft.rs:61-115 on a copy of `a` (uint64 [n,4]); returns the transformed copy."""
    a = _c(a).reshape(-1, 4).copy()
    n = a.shape[0]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().so_best_fft(ctypes.c_int(field), _p(a), ctypes.c_uint32(log_n), _p(_c(omega).reshape(4)), ctypes.c_int(threads))
    assert rc == 0
    return a


def scale(field: int, a: np.ndarray, s: np.ndarray) -> np.ndarray:
    a = _c(a).reshape(-1, 4).copy()
    lib().so_scale(ctypes.c_int(field), _p(a), ctypes.c_size_t(a.shape[0]), _p(_c(s).reshape(4)))
    return a


def coset_scale(field: int, a: np.ndarray, z: np.ndarray, z2: np.ndarray) -> np.ndarray:
    a = _c(a).reshape(-1, 4).copy()
    lib().so_coset_scale(ctypes.c_int(field), _p(a), ctypes.c_size_t(a.shape[0]), _p(_c(z).reshape(4)), _p(_c(z2).reshape(4)))
    return a


def _fr1(v: int) -> np.ndarray:
    return pyref.to_mont_limbs([v], pyref.FR).reshape(4)


def fft(a: np.ndarray, threads: int = 1) -> np.ndarray:
    """fft::fft, src/fft.rs:160-165 (bn256 Fr)."""
    k = int(np.asarray(a).reshape(-1, 4).shape[0]).bit_length() - 1
    return best_fft(pyref.FIELD_FR, a, _fr1(pyref.omega_for(k, False)), threads)


def ifft(a: np.ndarray, threads: int = 1) -> np.ndarray:
    """fft::ifft, src/fft.rs:168-182."""
    k = int(np.asarray(a).reshape(-1, 4).shape[0]).bit_length() - 1
    out = best_fft(pyref.FIELD_FR, a, _fr1(pyref.omega_for(k, True)), threads)
    return scale(pyref.FIELD_FR, out, _fr1(pow(pyref.FR_TWO_INV, k, pyref.FR)))


def coset_fft(a: np.ndarray, zeta: int = pyref.FR_ZETA, threads: int = 1) -> np.ndarray:
    """fft::coset_fft, src/fft.rs:186-190."""
    z, z2 = _fr1(zeta), _fr1(zeta * zeta % pyref.FR)
    return fft(coset_scale(pyref.FIELD_FR, a, z, z2), threads)


def coset_ifft(a: np.ndarray, zeta: int = pyref.FR_ZETA, threads: int = 1) -> np.ndarray:
    """fft::coset_ifft, src/fft.rs:194-198."""
    z, z2 = _fr1(zeta), _fr1(zeta * zeta % pyref.FR)
    return coset_scale(pyref.FIELD_FR, ifft(a, threads), z2, z)


def field_binop(name: str, field: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a, b = _c(a).reshape(-1, 4), _c(b).reshape(-1, 4)
    out = np.zeros_like(a)
    getattr(lib(), f"so_field_{name}")(ctypes.c_int(field), _p(a), _p(b), _p(out), ctypes.c_size_t(a.shape[0]))
    return out


def field_inv(field: int, a: np.ndarray) -> np.ndarray:
    a = _c(a).reshape(-1, 4)
    out = np.zeros_like(a)
    lib().so_field_inv(ctypes.c_int(field), _p(a), _p(out), ctypes.c_size_t(a.shape[0]))
    return out
