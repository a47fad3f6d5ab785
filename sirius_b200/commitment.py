"""Mirror of reference src/commitment.rs: `CommitmentKey` with a device-resident key.

    CommitmentKey(curve, ck)      ~ CommitmentKey { ck: Box<[C]> }        (src/commitment.rs:29-32)
    .commit(v) -> affine point    ~ CommitmentKey::commit                 (src/commitment.rs:81-90)
    TooLongInput                  ~ Error::TooLongInput{input_len,limit}  (src/commitment.rs:24-27)
    len(), is_empty()             ~ src/commitment.rs:47-53
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class TooLongInput(ValueError):
    """commitment::Error::TooLongInput."""

    def __init__(self, input_len: int, limit: int):
        super().__init__(f"Can't commit too long input: input len: {input_len}, but limit is {limit}")
        self.input_len = input_len
        self.limit = limit


def _as_u64(a, width: int) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.uint64)
    return arr.reshape(-1, width)


class CommitmentKey:
    def __init__(self, curve: int, ck, window_bits: int = 0):
        lib = _lib.load()
        self.curve = int(curve)
        self._host = _as_u64(ck, 8)
        self._h = ctypes.c_void_p()
        _lib.check(
            lib.sb_ck_register(self.curve, self._host.ctypes.data_as(_lib.u64p), self._host.shape[0], int(window_bits), ctypes.byref(self._h))
        )

    @classmethod
    def from_device(cls, curve: int, d_ptr: int, n: int, window_bits: int = 0, stream: int = 0) -> "CommitmentKey":
        lib = _lib.load()
        self = cls.__new__(cls)
        self.curve = int(curve)
        self._host = None
        self._h = ctypes.c_void_p()
        _lib.check(lib.sb_ck_register_device(self.curve, ctypes.c_void_p(d_ptr), n, int(window_bits), ctypes.c_void_p(stream), ctypes.byref(self._h)))
        return self

    # -- Rust API mirror -------------------------------------------------------------------------
    def len(self) -> int:
        return int(_lib.load().sb_ck_len(self._h))

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def add_window(self, window_bits: int, stream: int = 0) -> None:
        """Register a further window width; commits pick the cheapest registered width per call."""
        _lib.check(_lib.load().sb_ck_add_window(self._h, int(window_bits), ctypes.c_void_p(stream or None)))

    @property
    def window_bits(self) -> int:
        return int(_lib.load().sb_ck_window_bits(self._h))

    @staticmethod
    def default_value() -> np.ndarray:
        """C::identity(), encoded (0,0)."""
        return np.zeros(8, dtype=np.uint64)

    # -- key cache files (src/commitment.rs:99-170): the raw memory of [C], 64 bytes per point ---------------
    def save_to_file(self, file_path: str) -> None:
        """`CommitmentKey::save_to_file`: the generators' bytes as they sit in memory."""
        if self._host is None:
            raise ValueError("this key was registered from device memory; no host copy to dump")
        with open(file_path, "wb") as f:
            f.write(self._host.tobytes())

    @classmethod
    def load_from_file(cls, curve: int, file_path: str, k: int, window_bits: int = 0) -> "CommitmentKey":
        """`CommitmentKey::load_from_file`: exactly 2^k points are read (`read_exact`)."""
        want = (1 << k) * 64
        with open(file_path, "rb") as f:
            buf = f.read(want)
        if len(buf) != want:
            raise EOFError("failed to fill whole buffer")  # io::ErrorKind::UnexpectedEof
        return cls(curve, np.frombuffer(buf, dtype=np.uint64).reshape(-1, 8).copy(), window_bits)

    @classmethod
    def load_or_setup_cache(cls, curve: int, cache_folder: str, label: str, k: int, setup=None) -> "CommitmentKey":
        """`CommitmentKey::load_or_setup_cache`: {cache_folder}/{label}/{k}.bin, validated point by point on the
        device.  Key generation (`setup`: Shake256 + hash_to_curve of the un-vendored halo2curves) is
        not part of the GPU hot path: pass `setup(k, label) -> uint64[2^k, 8]` to create a missing file."""
        import os

        path = os.path.join(cache_folder, label, f"{k}.bin")
        if os.path.exists(path):
            key = cls.load_from_file(curve, path, k)
            bad = np.zeros(1, dtype=np.uint64)
            _lib.check(_lib.load().sb_points_on_curve(curve, key._host.ctypes.data_as(_lib.u64p), key._host.shape[0], bad.ctypes.data_as(_lib.u64p)))
            if int(bad[0]):
                key.close()
                raise ValueError("Wrong file in cache, some ptr out of curve")  # io::ErrorKind::InvalidData
            return key
        if setup is None:
            raise NotImplementedError("CommitmentKey::setup needs halo2curves' hash_to_curve (un-vendored); supply `setup`")
        key = cls(curve, setup(k, label))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        key.save_to_file(path)
        return key

    def commit(self, v) -> np.ndarray:
        """sum_i v[i] * ck[i] as an affine point (uint64[8]); raises TooLongInput like the Rust Err."""
        s = _as_u64(v, 4)
        n = s.shape[0]
        if n > self.len():
            raise TooLongInput(n, self.len())
        out = np.zeros(8, dtype=np.uint64)
        rc = _lib.load().sb_msm(self._h, s.ctypes.data_as(_lib.u64p), n, out.ctypes.data_as(_lib.u64p))
        _lib.check(rc)
        return out

    def commit_batch(self, vs) -> np.ndarray:
        """[commit(v) for v in vs] in one device pipeline (all vectors the same length) -> uint64 [len(vs), 8]."""
        arrs = [_as_u64(v, 4) for v in vs]
        if not arrs:
            return np.zeros((0, 8), dtype=np.uint64)
        n = arrs[0].shape[0]
        assert all(a.shape[0] == n for a in arrs), "commit_batch: vectors must have equal length"
        if n > self.len():
            raise TooLongInput(n, self.len())
        ptrs = (_lib.u64p * len(arrs))(*[a.ctypes.data_as(_lib.u64p) for a in arrs])
        out = np.zeros((len(arrs), 8), dtype=np.uint64)
        _lib.check(_lib.load().sb_msm_batch(self._h, ptrs, n, len(arrs), out.ctypes.data_as(_lib.u64p)))
        return out

    def commit_batch_device(self, d_scalars: int, n: int, stride: int, batch: int, d_out_xy: int = 0, d_out_xyzz: int = 0, stream: int = 0) -> None:
        if n > self.len():
            raise TooLongInput(n, self.len())
        rc = _lib.load().sb_msm_batch_device(
            self._h, ctypes.c_void_p(d_scalars), n, stride, batch, ctypes.c_void_p(d_out_xy or None), ctypes.c_void_p(d_out_xyzz or None), ctypes.c_void_p(stream or None)
        )
        _lib.check(rc)

    def commit_device(self, d_scalars: int, n: int, d_out_xy: int = 0, d_out_xyzz: int = 0, stream: int = 0) -> None:
        """Scalars already in HBM; enqueues on `stream` and returns without synchronising."""
        if n > self.len():
            raise TooLongInput(n, self.len())
        rc = _lib.load().sb_msm_device(
            self._h, ctypes.c_void_p(d_scalars), n, ctypes.c_void_p(d_out_xy or None), ctypes.c_void_p(d_out_xyzz or None), ctypes.c_void_p(stream or None)
        )
        _lib.check(rc)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().sb_ck_release(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def smallest_power(n: int, K: int) -> int:
    """`smallest_power` inside `setup_smallest_key` (src/commitment.rs:177-180): the smallest w with 2^w >= n * 2^K,
    computed as the reference does it, `((n * 2^K) as f64).log2().ceil() as usize` (n = 0 gives -inf -> 0)."""
    import math

    v = float(n * (1 << K))
    return 0 if v == 0.0 else max(0, math.ceil(math.log2(v)))


def smallest_key_log2(k_table_size: int, num_advice_columns: int, num_lookups: int, num_selectors: int, num_fixed_columns: int) -> int:
    """The `k` that `setup_smallest_key` passes to `CommitmentKey::setup` (src/commitment.rs:182-185): the key must
    cover a witness round (advice + 5 columns per lookup) and the selector + fixed columns."""
    p1 = smallest_power(num_advice_columns + 5 * num_lookups, k_table_size)
    p2 = smallest_power(num_selectors + num_fixed_columns, k_table_size)
    return max(p1, p2)


def setup_smallest_key(curve: int, k_table_size: int, num_advice_columns: int, num_lookups: int, num_selectors: int, num_fixed_columns: int,
                       tag: str, setup) -> CommitmentKey:
    """`setup_smallest_key(k_table_size, cs, tag)` (src/commitment.rs:172-186) with the ConstraintSystem's four counts
    spelled out.  `setup(k, tag) -> uint64[2^k, 8]` stands in for `CommitmentKey::setup`, whose hash_to_curve lives in
    the un-vendored halo2curves."""
    k = smallest_key_log2(k_table_size, num_advice_columns, num_lookups, num_selectors, num_fixed_columns)
    return CommitmentKey(curve, setup(k, tag))

