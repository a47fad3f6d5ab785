"""Mirror of the SPS lookup columns, the permutation decider and the witness assembly that sit between two
commits of the prover step; every computation is a call through the C ABI.

    lookup::Arguments (compress_from, to_expressions, evaluate_coefficient_1)     src/plonk/lookup.rs:72-343
    ArgumentCoefficient1::evaluate_coefficient_2                                 src/plonk/lookup.rs:345-370
    PlonkStructure::is_sat_log_derivative                                         src/plonk/mod.rs:363-397
    PlonkStructure::run_sps_protocol_{1,2,3}                                      src/plonk/mod.rs:431-660
    sparse::matrix_multiply + is_sat_permutation's mismatch count                 src/polynomial/sparse.rs:7-20,
                                                                                  src/nifs/sangria/mod.rs:385-453
    util::{concatenate_with_padding, batch_invert_assigned}                      src/util/mod.rs:120-153,214-218

Cells are uint64 [n,4] Montgomery arrays.  The reference's layout quirk is kept: the SPS protocol writes
concat(ls, ts, ms) / concat(hs, gs) while eval_advice_var and is_sat_log_derivative read the lookup columns
interleaved per lookup -- the two agree for one lookup argument only (see oracle/lookup_ref.py, Q1).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .commitment import CommitmentKey
from .polynomial import Expression, GraphEvaluator, compress_expression


def _cells(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_lib.u64p)


# ------------------------------------------------------------------------------------------ util/mod.rs


def concatenate_with_padding(vs: Sequence[np.ndarray], pad_size: int) -> np.ndarray:
    """util/mod.rs:214-218 on the host (pure data movement; pads, never truncates)."""
    vs = [_cells(v) for v in vs]
    total = sum(max(v.shape[0], pad_size) for v in vs)
    out = np.zeros((total, 4), dtype=np.uint64)
    pos = 0
    for v in vs:
        out[pos : pos + v.shape[0]] = v
        pos += max(v.shape[0], pad_size)
    return out


def concatenate_with_padding_device(vs: Sequence[np.ndarray], pad_size: int, d_out: int, out_capacity: int, stream: int = 0) -> int:
    """The same, written straight into a device round vector (sb_concat_pad_device); returns the cell count."""
    lib = _lib.load()
    vs = [_cells(v) for v in vs]
    ptrs = (_lib.u64p * max(1, len(vs)))(*[_p(v) for v in vs])
    lens = (ctypes.c_size_t * max(1, len(vs)))(*[v.shape[0] for v in vs])
    got = ctypes.c_size_t(0)
    _lib.check(lib.sb_concat_pad_device(ptrs, lens, len(vs), pad_size, ctypes.c_void_p(d_out), out_capacity, ctypes.byref(got), ctypes.c_void_p(stream or None)))
    return int(got.value)


def batch_invert_assigned(field: int, numerators: np.ndarray, denominators: np.ndarray) -> np.ndarray:
    """util/mod.rs:128-153 for one column: numerator * denominator^-1 with a zero denominator inverting to zero.
    Callers pass denominator 1 for Assigned::Trivial / Assigned::Zero (numerator() is then the value / zero)."""
    lib = _lib.load()
    num, den = _cells(numerators), _cells(denominators)
    assert num.shape == den.shape
    out = np.zeros_like(num)
    _lib.check(lib.sb_scaled_inverse(field, _p(den), None, _p(num), _p(out), num.shape[0]))
    return out


# ------------------------------------------------------------------------------------------ polynomial/sparse.rs


class SparseMatrix:
    """Vec<(row, col, value)> of an N x N matrix (sparse.rs:5) resident on the device."""

    def __init__(self, field: int, entries: Sequence[Tuple[int, int, np.ndarray]], N: int):
        lib = _lib.load()
        rows = np.array([e[0] for e in entries], dtype=np.uint64)
        cols = np.array([e[1] for e in entries], dtype=np.uint64)
        vals = _cells(np.stack([np.asarray(e[2], dtype=np.uint64).reshape(4) for e in entries])) if entries else np.zeros((0, 4), dtype=np.uint64)
        self._h = ctypes.c_void_p()
        self.N = N
        _lib.check(lib.sb_sparse_register(field, _p(rows) if len(entries) else None, _p(cols) if len(entries) else None,
                                          _p(vals) if len(entries) else None, len(entries), N, ctypes.byref(self._h)))

    def mismatch_count(self, Z: np.ndarray) -> int:
        """#{row : (P*Z)[row] != Z[row]}"""
        lib = _lib.load()
        Z = _cells(Z)
        got = ctypes.c_uint64(0)
        _lib.check(lib.sb_sparse_mismatch(self._h, _p(Z), Z.shape[0], ctypes.cast(ctypes.byref(got), _lib.u64p)))
        return int(got.value)

    def close(self):
        if self._h.value:
            _lib.load().sb_sparse_release(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PermCheckFail(Exception):
    """VerifyError::PermCheckFail { mismatch_count } (src/nifs/sangria/mod.rs:448-452)"""

    def __init__(self, mismatch_count: int):
        super().__init__(f"permutation check fail: mismatch_count {mismatch_count}")
        self.mismatch_count = mismatch_count


def is_sat_permutation(P: SparseMatrix, instances_flat: np.ndarray, W0: np.ndarray, k: int, num_advice: int) -> None:
    """src/nifs/sangria/mod.rs:420-452: Z = instances ++ W[0][.. 2^k * num_advice]; P*Z must equal Z."""
    Z = np.concatenate([_cells(instances_flat), _cells(W0)[: (1 << k) * num_advice]])
    bad = P.mismatch_count(Z)
    if bad:
        raise PermCheckFail(bad)


# ------------------------------------------------------------------------------------------ plonk/lookup.rs


@dataclass
class ArgumentCoefficient2:
    hs: List[np.ndarray]
    gs: List[np.ndarray]


@dataclass
class ArgumentCoefficient1:
    field: int
    ls: List[np.ndarray]
    ts: List[np.ndarray]
    ms: List[np.ndarray]

    def evaluate_coefficient_2(self, r) -> ArgumentCoefficient2:
        """lookup.rs:351-365: h = 1/(l+r), g = m/(t+r), zero where the denominator is zero"""
        lib = _lib.load()
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        hs, gs = [], []
        for l, t, m in zip(self.ls, self.ts, self.ms):
            assert l.shape == t.shape == m.shape, "zip_eq"
            h, g = np.zeros_like(l), np.zeros_like(t)
            _lib.check(lib.sb_lookup_inverses(self.field, _p(l), _p(t), _p(m), _p(r), l.shape[0], _p(h), _p(g)))
            hs.append(h)
            gs.append(g)
        return ArgumentCoefficient2(hs, gs)


class Arguments:
    """lookup.rs:72-83"""

    def __init__(self, lookup_polys: List[Expression], table_polys: List[Expression], has_vector_lookup: bool):
        self.lookup_polys, self.table_polys, self.has_vector_lookup = lookup_polys, table_polys, has_vector_lookup
        self._progs = None

    @staticmethod
    def compress_from(inputs: Sequence[Sequence[Expression]], tables: Sequence[Sequence[Expression]]) -> Optional["Arguments"]:
        """lookup.rs:87-131 on already converted expressions (one list of input / table expressions per argument)."""
        max_len = max((len(a) for a in inputs), default=0)
        if max_len == 0:
            return None
        comp = lambda es: compress_expression(list(es), 0) if len(es) > 1 else es[0]
        return Arguments([comp(a) for a in inputs], [comp(t) for t in tables], max_len > 1)

    def num_lookups(self) -> int:
        return len(self.lookup_polys)

    def vanishing_lookup_polys(self, num_selectors: int, num_fixed: int, num_advice: int) -> List[Expression]:
        """lookup.rs:141-170"""
        off = num_selectors + num_fixed + num_advice
        ls = [L - Expression.Polynomial(off + i * 5) for i, L in enumerate(self.lookup_polys)]
        ts = [T - Expression.Polynomial(off + i * 5 + 1) for i, T in enumerate(self.table_polys)]
        return ls + ts

    def log_derivative_expr(self, num_selectors: int, num_fixed: int, num_advice: int, lookup_index: int, challenge_index: int):
        """lookup.rs:178-200"""
        r = Expression.Challenge(challenge_index)
        off = num_selectors + num_fixed + num_advice
        l, t, m, h, g = [Expression.Polynomial(off + lookup_index * 5 + i) for i in range(5)]
        return h * (l + r) - Expression.Constant(1), g * (t + r) - m

    def log_derivative_lhs_and_rhs(self, num_selectors: int, num_fixed: int, num_advice: int) -> List[Expression]:
        """lookup.rs:203-212"""
        ci = 1 if self.has_vector_lookup else 0
        out: List[Expression] = []
        for i in range(self.num_lookups()):
            out.extend(self.log_derivative_expr(num_selectors, num_fixed, num_advice, i, ci))
        return out

    def to_expressions(self, num_selectors: int, num_fixed: int, num_advice: int) -> List[Expression]:
        """lookup.rs:133-137"""
        return self.vanishing_lookup_polys(num_selectors, num_fixed, num_advice) + self.log_derivative_lhs_and_rhs(num_selectors, num_fixed, num_advice)

    def _programs(self, S):
        from .sangria import Program

        if self._progs is None:
            self._progs = (
                [Program(S.field, GraphEvaluator.new(p, S.modulus)) for p in self.lookup_polys],
                [Program(S.field, GraphEvaluator.new(p, S.modulus)) for p in self.table_polys],
            )
        return self._progs

    def evaluate_coefficient_1(self, S, advice: Sequence[np.ndarray], r) -> ArgumentCoefficient1:
        """lookup.rs:314-336.  `advice` are the per-column vectors of the LookupEvalDomain (eval.rs:84-131)."""
        lib = _lib.load()
        n = 1 << S.k
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(1, 4)
        adv = [_cells(a) for a in advice]
        assert all(a.shape[0] == n for a in adv), "advice columns must hold 2^k rows"
        W = np.concatenate(adv) if adv else np.zeros((0, 4), dtype=np.uint64)
        ptrs = (_lib.u64p * 1)(_p(W))
        lens = (ctypes.c_size_t * 1)(W.shape[0])
        lp, tp = self._programs(S)

        def run(prog):
            out = np.zeros((n, 4), dtype=np.uint64)
            # LookupEvalDomain: one witness round holding the advice columns, no lookup variables, challenges = [r]
            _lib.check(lib.sb_expr_eval(prog._h, S._cols, len(adv), 0, ptrs, lens, 1, None, None, 0, _p(r), 1, _p(out)))
            return out

        ls = [run(p) for p in lp]
        ts = [run(p) for p in tp]
        ms = []
        for l, t in zip(ls, ts):
            m = np.zeros_like(t)
            _lib.check(lib.sb_lookup_multiplicity(S.field, _p(l), l.shape[0], _p(t), t.shape[0], _p(m)))
            ms.append(m)
        return ArgumentCoefficient1(S.field, ls, ts, ms)

    def close(self):
        if self._progs is not None:
            for p in self._progs[0] + self._progs[1]:
                p.close()
            self._progs = None


def is_sat_log_derivative(S, W: Sequence[np.ndarray]) -> bool:
    """plonk/mod.rs:363-397: sum_i h_i == sum_i g_i per lookup, h at column positions 0,2,.. and g at 1,3,..
    of the last witness round."""
    lib = _lib.load()
    args = getattr(S, "lookup_arguments", None)
    nl = S.num_lookups
    if nl == 0:
        return True
    n = 1 << S.k
    Wr = _cells(W[2] if (args is not None and args.has_vector_lookup) else W[1])
    out = np.zeros(4, dtype=np.uint64)
    for i in range(nl):
        h = np.ascontiguousarray(Wr[(2 * i) * n : (2 * i + 1) * n])
        g = np.ascontiguousarray(Wr[(2 * i + 1) * n : (2 * i + 2) * n])
        _lib.check(lib.sb_sum_diff(S.field, _p(h), _p(g), n, _p(out)))
        if out.any():
            return False
    return True


def run_sps_protocol(S, advice: Sequence[np.ndarray], ck: CommitmentKey, challenge: Callable[[int, np.ndarray], np.ndarray]):
    """run_sps_protocol_{1,2,3} (plonk/mod.rs:431-660) minus the random oracle: `challenge(round, commitment)`
    stands in for absorb_point + squeeze (host-side Poseidon stays with the caller).
    -> (W rounds, W_commitments, challenges)."""
    n = 1 << S.k
    args: Optional[Arguments] = getattr(S, "lookup_arguments", None)
    zero = np.zeros(4, dtype=np.uint64)
    if args is None:
        W1 = concatenate_with_padding(advice, n)
        return [W1], [ck.commit(W1)], []
    if not args.has_vector_lookup:
        c1 = args.evaluate_coefficient_1(S, advice, zero)
        W1 = np.concatenate([concatenate_with_padding(advice, n), concatenate_with_padding(c1.ls + c1.ts + c1.ms, n)])
        C1 = ck.commit(W1)
        r1 = challenge(0, C1)
        c2 = c1.evaluate_coefficient_2(r1)
        W2 = concatenate_with_padding(c2.hs + c2.gs, n)
        C2 = ck.commit(W2)
        r2 = challenge(1, C2)
        return [W1, W2], [C1, C2], [r1, r2]
    W1 = concatenate_with_padding(advice, n)
    C1 = ck.commit(W1)
    r1 = challenge(0, C1)
    c1 = args.evaluate_coefficient_1(S, advice, r1)
    W2 = concatenate_with_padding(c1.ls + c1.ts + c1.ms, n)
    C2 = ck.commit(W2)
    r2 = challenge(1, C2)
    c2 = c1.evaluate_coefficient_2(r2)
    W3 = concatenate_with_padding(c2.hs + c2.gs, n)
    C3 = ck.commit(W3)
    r3 = challenge(2, C3)
    return [W1, W2, W3], [C1, C2, C3], [r1, r2, r3]
