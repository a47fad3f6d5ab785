"""Mirror of the Sangria NIFS prover steps that sit on the hot path (reference src/nifs/sangria).

    PlonkStructure (the fields the hot path reads)      src/plonk/mod.rs:127-193
    VanillaFS.commit_cross_terms(ck, S, U1, W1, U2, W2) src/nifs/sangria/mod.rs:102-158
    RelaxedPlonkWitness.fold(W2, cross_terms, r)        src/nifs/sangria/accumulator.rs:363-404

Witness round vectors are uint64 [len,4] Montgomery arrays, column-major as `concatenate_with_padding` lays
them out (src/util/mod.rs:214-218).  Challenges / u / r are uint64[4] Montgomery limbs or lists thereof.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .commitment import CommitmentKey
from .polynomial import (
    OP_MUL,
    CompressedGates,
    GraphEvaluator,
)


def _to_mont(vals: Sequence[int], modulus: int) -> np.ndarray:
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % modulus) * (1 << 256) % modulus
        for k in range(4):
            out[i, k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


class Program:
    """A GraphEvaluator uploaded through sb_expr_compile."""

    def __init__(self, field: int, ev: GraphEvaluator):
        lib = _lib.load()
        calcs = (_lib.sb_calc * max(1, len(ev.calculations)))()
        for i, (op, a, b, target) in enumerate(ev.calculations):
            c = calcs[i]
            c.opcode, c.a_kind, c.a_index, c.a_rot = op, a[0], a[1], a[2]
            if b is not None and op <= OP_MUL:
                c.b_kind, c.b_index, c.b_rot = b
            c.target = target
        consts = _to_mont(ev.constants, ev.modulus)
        rots = np.array(ev.rotations if ev.rotations else [0], dtype=np.int32)
        self.field = field
        self._h = ctypes.c_void_p()
        _lib.check(
            lib.sb_expr_compile(
                field, ctypes.cast(calcs, ctypes.c_void_p), len(ev.calculations), consts.ctypes.data_as(_lib.u64p), len(ev.constants),
                rots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(ev.rotations), ctypes.byref(self._h),
            )
        )

    @property
    def num_slots(self) -> int:
        return int(_lib.load().sb_expr_num_slots(self._h))

    def close(self):
        if self._h.value:
            _lib.load().sb_expr_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class PlonkStructure:
    """Fields of reference PlonkStructure read by commit_cross_terms / fold (src/plonk/mod.rs:127-193)."""

    field: int
    modulus: int
    k: int
    selectors: List[np.ndarray]          # Vec<Vec<bool>> as uint8 [2^k]
    fixed_columns: List[np.ndarray]      # Vec<Vec<F>>   as uint64 [2^k,4]
    num_advice_columns: int
    num_lookups: int
    custom_gates_lookup_compressed: CompressedGates
    gates: Optional[List] = None         # S.gates: the individual gate Expressions (Protogalaxy evaluates them one by one)
    lookup_arguments: Optional[object] = None  # lookup::Arguments (sirius_b200.lookup.Arguments), src/plonk/mod.rs:156

    def __post_init__(self):
        lib = _lib.load()
        self.selectors = [np.ascontiguousarray(s, dtype=np.uint8) for s in self.selectors]
        self.fixed_columns = [np.ascontiguousarray(f, dtype=np.uint64).reshape(-1, 4) for f in self.fixed_columns]
        u8p = ctypes.POINTER(ctypes.c_uint8)
        sel = (u8p * max(1, len(self.selectors)))(*[s.ctypes.data_as(u8p) for s in self.selectors])
        fx = (_lib.u64p * max(1, len(self.fixed_columns)))(*[f.ctypes.data_as(_lib.u64p) for f in self.fixed_columns])
        self._cols = ctypes.c_void_p()
        _lib.check(lib.sb_columns_register(self.field, self.k, sel, len(self.selectors), fx, len(self.fixed_columns), ctypes.byref(self._cols)))
        # GraphEvaluator::new(homogeneous) -- what the Rust shim compiles once per structure
        self._hom_prog = Program(self.field, GraphEvaluator.new(self.custom_gates_lookup_compressed.homogeneous, self.modulus))

    def is_sat(self, ck: CommitmentKey, U_challenges, U_W_commitments, W: Sequence[np.ndarray]) -> None:
        """PlonkStructure::is_sat (src/plonk/mod.rs:304-361) without the host-side sps_verify (random oracle): the
        compressed gate expression vanishes on every row, the log-derivative sums agree and the witness rounds re-open."""
        if getattr(self, "_compressed_prog", None) is None:
            self._compressed_prog = Program(self.field, GraphEvaluator.new(self.custom_gates_lookup_compressed.compressed, self.modulus))
        got = evaluate_rows(self, self._compressed_prog, W, np.asarray(U_challenges, dtype=np.uint64).reshape(-1, 4))
        mismatch = int(np.count_nonzero(np.any(got != 0, axis=1)))
        if mismatch:
            raise EvaluationMismatch(mismatch, 1 << self.k)
        from .lookup import is_sat_log_derivative

        if not is_sat_log_derivative(self, W):
            raise LogDerivativeNotSat()
        bad = sum(1 for Ci, Wi in zip(U_W_commitments, W) if not np.array_equal(ck.commit(Wi), np.asarray(Ci, dtype=np.uint64).reshape(8)))
        if bad:
            raise CommitmentMismatch(bad)

    def gate_programs(self):
        """GraphEvaluator::new(gate) per gate (get_evaluate_witness_fn, src/plonk/mod.rs:697-701), compiled once."""
        if getattr(self, "_gate_progs", None) is None:
            assert self.gates, "PlonkStructure.gates not set"
            self._gate_progs = [Program(self.field, GraphEvaluator.new(g, self.modulus)) for g in self.gates]
        return self._gate_progs

    @property
    def degree(self) -> int:
        """number of cross terms = grouped().len() - 1 (src/plonk/mod.rs:399)"""
        return self.custom_gates_lookup_compressed.degree

    def close(self):
        if getattr(self, "_cols", None) is not None and self._cols.value:
            _lib.load().sb_columns_release(self._cols)
            self._cols = ctypes.c_void_p()
        if getattr(self, "_hom_prog", None) is not None:
            self._hom_prog.close()
        for gp in getattr(self, "_gate_progs", None) or []:
            gp.close()
        if getattr(self, "_compressed_prog", None) is not None:
            self._compressed_prog.close()


def _rounds(W: Sequence[np.ndarray]):
    arrs = [np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4) for w in W]
    ptrs = (_lib.u64p * len(arrs))(*[a.ctypes.data_as(_lib.u64p) for a in arrs])
    lens = (ctypes.c_size_t * len(arrs))(*[a.shape[0] for a in arrs])
    return arrs, ptrs, lens


class EvaluationMismatch(Exception):
    """plonk::Error::EvaluationMismatch { mismatch_count, total_row } (src/plonk/mod.rs, used by the deciders)."""

    def __init__(self, mismatch_count: int, total_row: int):
        super().__init__(f"(Relaxed) plonk relation not satisfied: mismatch_count {mismatch_count}, total_row {total_row}")
        self.mismatch_count, self.total_row = mismatch_count, total_row


class CommitmentMismatch(Exception):
    """plonk::Error::CommitmentMismatch { mismatch_count }"""

    def __init__(self, mismatch_count: int):
        super().__init__(f"commitment of witness mismatch: {mismatch_count}")
        self.mismatch_count = mismatch_count


class LogDerivativeNotSat(Exception):
    """plonk::Error::LogDerivativeNotSat (src/plonk/mod.rs:348-350, src/nifs/sangria/mod.rs:378-380)"""


class ECommitmentMismatch(Exception):
    """VerifyError::ECommitmentMismatch (src/nifs/sangria/mod.rs:319-320)"""


def evaluate_rows(S: "PlonkStructure", prog: Program, W: Sequence[np.ndarray], challenges: np.ndarray, W2: Optional[Sequence[np.ndarray]] = None) -> np.ndarray:
    """GraphEvaluator::evaluate for every row (sb_expr_eval): uint64 [2^k,4]."""
    lib = _lib.load()
    a1, p1, l1 = _rounds(W)
    ch = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((1 << S.k, 4), dtype=np.uint64)
    if W2 is not None:
        a2, p2, l2 = _rounds(W2)
        _lib.check(lib.sb_expr_eval(prog._h, S._cols, S.num_advice_columns, S.num_lookups, p1, l1, len(a1), p2, l2, len(a2),
                                    ch.ctypes.data_as(_lib.u64p), ch.shape[0], out.ctypes.data_as(_lib.u64p)))
    else:
        _lib.check(lib.sb_expr_eval(prog._h, S._cols, S.num_advice_columns, S.num_lookups, p1, l1, len(a1), None, None, 0,
                                    ch.ctypes.data_as(_lib.u64p), ch.shape[0], out.ctypes.data_as(_lib.u64p)))
    return out


class VanillaFS:
    @staticmethod
    def is_sat_accumulation(S: "PlonkStructure", U_challenges, U_u, W: Sequence[np.ndarray], E: np.ndarray) -> None:
        """src/nifs/sangria/mod.rs:334-383: homogeneous gate polynomial on (W, challenges ++ [u]) must equal E row by
        row, then the log-derivative sum check of the lookup arguments."""
        ch = np.concatenate([np.asarray(U_challenges, dtype=np.uint64).reshape(-1, 4), np.asarray(U_u, dtype=np.uint64).reshape(1, 4)])
        got = evaluate_rows(S, S._hom_prog, W, ch)
        mismatch = int(np.count_nonzero(np.any(got != np.asarray(E, dtype=np.uint64).reshape(-1, 4), axis=1)))
        if mismatch:
            raise EvaluationMismatch(mismatch, 1 << S.k)
        from .lookup import is_sat_log_derivative

        if not is_sat_log_derivative(S, W):
            raise LogDerivativeNotSat()

    @staticmethod
    def is_sat_witness_commit(ck: CommitmentKey, W_commitments, W: Sequence[np.ndarray], E: np.ndarray, E_commitment) -> None:
        """src/nifs/sangria/mod.rs:455-474: every W round and E re-open to their commitments."""
        bad = sum(1 for Ci, Wi in zip(W_commitments, W) if not np.array_equal(ck.commit(Wi), np.asarray(Ci, dtype=np.uint64).reshape(8)))
        if bad:
            raise CommitmentMismatch(bad)
        if not np.array_equal(ck.commit(E), np.asarray(E_commitment, dtype=np.uint64).reshape(8)):
            raise ECommitmentMismatch()

    @staticmethod
    def commit_cross_terms(ck: CommitmentKey, S: PlonkStructure, U1_challenges, U1_u, W1: Sequence[np.ndarray], U2_challenges,
                           W2: Sequence[np.ndarray]):
        """-> (cross_terms: list of uint64 [2^k,4], cross_term_commits: uint64 [d,8]).
        challenges = U1.challenges ++ [U1.u] and U2.challenges ++ [1] (src/nifs/sangria/mod.rs:113-118)."""
        lib = _lib.load()
        one = _to_mont([1], S.modulus)
        c1 = np.concatenate([np.asarray(U1_challenges, dtype=np.uint64).reshape(-1, 4), np.asarray(U1_u, dtype=np.uint64).reshape(1, 4)])
        c2 = np.concatenate([np.asarray(U2_challenges, dtype=np.uint64).reshape(-1, 4), one])
        c1, c2 = np.ascontiguousarray(c1), np.ascontiguousarray(c2)
        a1, p1, l1 = _rounds(W1)
        a2, p2, l2 = _rounds(W2)
        n = 1 << S.k
        d = S.degree
        T = [np.zeros((n, 4), dtype=np.uint64) for _ in range(d)]
        outp = (_lib.u64p * d)(*[t.ctypes.data_as(_lib.u64p) for t in T])
        _lib.check(
            lib.sb_cross_terms(
                S._hom_prog._h, d, S._cols, S.num_advice_columns, S.num_lookups, p1, l1, len(a1), p2, l2, len(a2),
                c1.ctypes.data_as(_lib.u64p), c2.ctypes.data_as(_lib.u64p), c1.shape[0], outp,
            )
        )
        commits = ck.commit_batch(T)  # cross_terms.iter().map(|v| ck.commit(v)) (:151-154)
        return T, commits


class RelaxedPlonkWitness:
    """W: round vectors, E: error vector (src/nifs/sangria/accumulator.rs:273-276, 488)."""

    def __init__(self, field: int, W: Sequence[np.ndarray], E: np.ndarray):
        self.field = field
        self.W = [np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4) for w in W]
        self.E = np.ascontiguousarray(E, dtype=np.uint64).reshape(-1, 4)

    def fold(self, W2: Sequence[np.ndarray], cross_terms: Sequence[np.ndarray], r) -> "RelaxedPlonkWitness":
        lib = _lib.load()
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        Wn = []
        for w1, w2 in zip(self.W, W2):
            w2 = np.ascontiguousarray(w2, dtype=np.uint64).reshape(-1, 4)
            assert w1.shape == w2.shape, "zip_eq"
            out = np.zeros_like(w1)
            _lib.check(lib.sb_axpy_fold(self.field, w1.ctypes.data_as(_lib.u64p), w2.ctypes.data_as(_lib.u64p), r.ctypes.data_as(_lib.u64p), out.ctypes.data_as(_lib.u64p), w1.shape[0]))
            Wn.append(out)
        Ts = [np.ascontiguousarray(t, dtype=np.uint64).reshape(-1, 4) for t in cross_terms]
        ptrs = (_lib.u64p * max(1, len(Ts)))(*[t.ctypes.data_as(_lib.u64p) for t in Ts])
        En = np.zeros_like(self.E)
        _lib.check(lib.sb_error_fold(self.field, self.E.ctypes.data_as(_lib.u64p), ptrs, len(Ts), r.ctypes.data_as(_lib.u64p), En.ctypes.data_as(_lib.u64p), self.E.shape[0]))
        return RelaxedPlonkWitness(self.field, Wn, En)
