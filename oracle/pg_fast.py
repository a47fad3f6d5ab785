"""Protogalaxy prover polynomials at full table sizes: the same reference algorithm as oracle/pg_ref.py (src/nifs/protogalaxy,
citations there), with the per-row gate evaluation done by the C GraphEvaluator interpreter (sirius_oracle.c
so_graph_evaluate, graph_evaluator.rs:91-149,361-388) and the binary tree `left + right * c_h` (poly/mod.rs:100-185,
330-413; mod.rs:586-639) done level by level on whole arrays with the C field operations.

TEST INFRASTRUCTURE ONLY.  pg_ref.py is the literal element-by-element restatement; tests/test_oracle_pg.py pins this
module to it at small k (both row modes) so that the GPU parity tests and bench.py's `--workload cyclefold_poseidon`
can check k >= 12 in seconds.

Witness / column arrays are uint64 [.., 4] Montgomery limbs; betas / delta / challenges / results are Python ints.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import numpy as np

import oracle
from oracle import expr_ref as E
from oracle import pyref as R

M = R.FR
F = R.FIELD_FR
_u64p = ctypes.POINTER(ctypes.c_uint64)


def _mont(v: int) -> np.ndarray:
    return R.to_mont_limbs([v % M], M).reshape(1, 4)


def lincomb(arrs: Sequence[np.ndarray], coefs: Sequence[int]) -> np.ndarray:
    """sum_j coefs[j] * arrs[j], cell by cell (FoldedWitness::new, folded_witness.rs:66-143; fold_witness, mod.rs:176-210)."""
    arrs = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4) for a in arrs]
    coef = R.to_mont_limbs([c % M for c in coefs], M)
    out = np.zeros_like(arrs[0])
    ptrs = (_u64p * len(arrs))(*[a.ctypes.data_as(_u64p) for a in arrs])
    rc = oracle.lib().so_lincomb(ctypes.c_int(F), ptrs, coef.ctypes.data_as(_u64p), ctypes.c_size_t(len(arrs)), ctypes.c_size_t(arrs[0].shape[0]), out.ctypes.data_as(_u64p))
    assert rc == 0
    return out


def tree(leaves: np.ndarray, mult: Sequence[int]) -> int:
    """Root of the perfect binary tree over the leaves with node(h) = left + right * mult[h] (so_beta_tree)."""
    v = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, 4)
    log_n = v.shape[0].bit_length() - 1
    assert 1 << log_n == v.shape[0] and len(mult) >= log_n
    m = R.to_mont_limbs([c % M for c in mult[:log_n]], M) if log_n else np.zeros((1, 4), dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    rc = oracle.lib().so_beta_tree(ctypes.c_int(F), v.ctypes.data_as(_u64p), ctypes.c_uint32(log_n), m.ctypes.data_as(_u64p), out.ctypes.data_as(_u64p))
    assert rc == 0
    return R.from_mont_limbs(out.reshape(1, 4), M)[0]


class Structure:
    """The slice of PlonkStructure Protogalaxy reads (single witness round): k, selector / fixed columns, the gates."""

    def __init__(self, k: int, selectors: Sequence[np.ndarray], fixed: Sequence[np.ndarray], num_advice: int, gates: Sequence):
        self.k, self.selectors, self.fixed, self.num_advice, self.gates = k, list(selectors), list(fixed), num_advice, list(gates)
        self.evs = [E.GraphEvaluator(g, M) for g in gates]

    def count(self) -> int:
        cnt, p = (1 << self.k) * len(self.gates), 1
        while p < cnt:
            p <<= 1
        return p

    def betas_count(self) -> int:
        return self.count().bit_length() - 1


def leaves(S: Structure, W: np.ndarray, challenges: Sequence[int], row_mode: str, threads: int = 0) -> np.ndarray:
    """get_evaluate_witness_fn for every index (plonk/mod.rs:683-718): leaf[g * 2^k + row] = gate_g(row'), row' = row
    ("correct") or `index & 2^k` wrapped = 0 ("compat"); zero padding up to the next power of two."""
    n = 1 << S.k
    W = np.ascontiguousarray(W, dtype=np.uint64).reshape(-1, 4)
    adv = [W[i * n:(i + 1) * n] for i in range(S.num_advice)]
    ch = R.to_mont_limbs([c % M for c in challenges], M) if len(challenges) else np.zeros((0, 4), dtype=np.uint64)
    out = np.zeros((S.count(), 4), dtype=np.uint64)
    for g, ev in enumerate(S.evs):
        vals = E.c_graph_evaluate(F, ev, S.selectors, S.fixed, adv, ch, S.k, threads=threads)
        if row_mode == "compat":
            vals = np.tile(vals[0:1], (n, 1))
        out[g * n:(g + 1) * n] = vals
    return out


def evaluate_e(S: Structure, W, challenges, betas, row_mode="compat", threads=0) -> int:
    return tree(leaves(S, W, challenges, row_mode, threads), list(betas)[: S.betas_count()])


def compute_F_evals(S: Structure, betas, delta, W, challenges, row_mode="compat", threads=0) -> List[int]:
    """F on the points of the order-`fft_points_count_F` subgroup (before the ifft)."""
    t = S.betas_count()
    betas = list(betas)[:t]
    deltas = [delta % M]
    for _ in range(t - 1):
        deltas.append(deltas[-1] * deltas[-1] % M)
    npts = 1
    while npts < t + 1:
        npts <<= 1
    Xs = list(R.iter_cyclic_subgroup(npts.bit_length() - 1))
    lv = leaves(S, W, challenges, row_mode, threads)
    return [tree(lv, [(b + X * d) % M for b, d in zip(betas, deltas)]) for X in Xs]


def compute_F(S: Structure, betas, delta, W, challenges, row_mode="compat", threads=0) -> List[int]:
    return R.ifft(compute_F_evals(S, betas, delta, W, challenges, row_mode, threads))


def g_points(S: Structure, traces_len: int, max_degree: int):
    p = 1
    while p < traces_len * max_degree + 1:
        p <<= 1
    lagrange_domain = (traces_len + 1).bit_length() - 1
    return list(R.iter_cyclic_subgroup(p.bit_length() - 1))[:p], lagrange_domain


def compute_G_evals(S: Structure, max_degree: int, betas_stroke, acc_W, acc_ch, traces_W, traces_ch, row_mode="compat", threads=0) -> List[int]:
    t = S.betas_count()
    bs = list(betas_stroke)[:t]
    points, ld = g_points(S, len(traces_W), max_degree)
    all_w = [acc_W] + list(traces_W)
    all_c = [list(acc_ch)] + [list(c) for c in traces_ch]
    out = []
    for X in points:
        L = R.eval_lagrange_polys(ld, X)[: len(all_w)]
        Wx = lincomb(all_w, L)
        Cx = [sum(L[j] * all_c[j][i] for j in range(len(all_c))) % M for i in range(len(all_c[0]))]
        out.append(tree(leaves(S, Wx, Cx, row_mode, threads), bs))
    return out


def compute_G(S: Structure, max_degree: int, betas_stroke, acc_W, acc_ch, traces_W, traces_ch, row_mode="compat", threads=0) -> List[int]:
    return R.ifft(compute_G_evals(S, max_degree, betas_stroke, acc_W, acc_ch, traces_W, traces_ch, row_mode, threads))


def fold_witness(acc_W, incoming_Ws, lagrange_for_gamma) -> np.ndarray:
    return lincomb([acc_W] + list(incoming_Ws), list(lagrange_for_gamma)[: 1 + len(incoming_Ws)])
