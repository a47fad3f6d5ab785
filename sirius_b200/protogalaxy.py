"""Mirror of the Protogalaxy prover polynomials (reference src/nifs/protogalaxy), bn256 Fr.

    PolyContext                     poly/mod.rs:205-269   (fft_log_domain_size_K keeps the reference's point-count-as-log)
    compute_F / compute_G           poly/mod.rs:68-203 / 308-425
    PolyChallenges.iter_beta_stroke poly/mod.rs:427-462
    compute_K / compute_K_from_G    poly/mod.rs:464-509
    evaluate_e_from_trace           src/nifs/protogalaxy/mod.rs:571-640
    ProtoGalaxy.fold_witness        src/nifs/protogalaxy/mod.rs:176-210
    lagrange::{iter_cyclic_subgroup, iter_eval_lagrange_poly_for_cyclic_group, eval_vanish_polynomial}
                                    src/polynomial/lagrange.rs:22-85
    UnivariatePoly::eval            src/polynomial/univariate.rs:67-75

The O(2^k) work (leaf evaluation on Lagrange-blended witnesses, the beta tree, the witness fold, the (i)FFTs) runs
on the GPU through the C ABI; the O(#points) scalar glue (Lagrange values, K from G on 2^K points) is host integer
arithmetic exactly as it stays host Rust in the integration.  Witnesses: one round, uint64 [A*2^k,4] Montgomery.

`row_mode`: ROW_COMPAT reproduces the reference (`index & total_row`); ROW_CORRECT is `index % total_row`.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import numpy as np

from . import _lib, fft
from .polynomial import Expression, QueryIndexContext
from .sangria import PlonkStructure, _to_mont

M = fft.FR_MODULUS
ROW_COMPAT, ROW_CORRECT = 0, 1


# ---------------------------------------------------------------------------------------------- lagrange (host ints)
def iter_cyclic_subgroup(log_n: int) -> List[int]:
    w = fft.get_omega_or_inv(log_n, False)
    out, cur = [], 1
    for _ in range(1 << log_n):
        out.append(cur)
        cur = cur * w % M
    return out


def eval_lagrange_polys(X: int, log_n: int) -> List[int]:
    n = 1 << log_n
    ninv = pow(n, -1, M)
    xn1 = (pow(X, n, M) - 1) % M
    out = []
    for v in iter_cyclic_subgroup(log_n):
        den = (X - v) % M
        out.append(1 if (xn1 == 0 and den == 0) else v * ninv % M * (xn1 * pow(den, -1, M) % M) % M)
    return out


def eval_vanish_polynomial(degree: int, point: int) -> int:
    return (pow(point, degree, M) - 1) % M


def poly_eval(coeffs: Sequence[int], x: int) -> int:
    acc, p = 0, 1
    for c in coeffs:
        acc = (acc + p * c) % M
        p = p * x % M
    return acc


def _from_mont(arr: np.ndarray) -> List[int]:
    rinv = pow(1 << 256, -1, M)
    out = []
    for row in np.asarray(arr, dtype=np.uint64).reshape(-1, 4):
        v = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
        out.append(v * rinv % M)
    return out


def _degree(e: Expression, ctx: QueryIndexContext) -> int:
    k = e.kind
    if k == "const":
        return 0
    if k == "poly":
        return 1 if e.a >= ctx.num_selectors + ctx.num_fixed else 0
    if k == "chal":
        return 1
    if k in ("neg", "scaled"):
        return _degree(e.a, ctx)
    if k == "sum":
        return max(_degree(e.a, ctx), _degree(e.b, ctx))
    return _degree(e.a, ctx) + _degree(e.b, ctx)


def _next_pow2(v: int) -> int:
    p = 1
    while p < v:
        p <<= 1
    return p


class PolyContext:
    def __init__(self, S: PlonkStructure, traces_len: int):
        self.S = S
        self.count_of_evaluation_with_padding = _next_pow2((1 << S.k) * len(S.gates))
        self.instances_to_fold = traces_len + 1
        assert self.instances_to_fold & (self.instances_to_fold - 1) == 0, "instances_to_fold.is_power_of_two()"
        ctx = QueryIndexContext(len(S.selectors), len(S.fixed_columns), S.num_advice_columns, 0, S.num_lookups)
        max_degree = max([_degree(g, ctx) for g in S.gates] or [0])
        self.fft_points_count_G = _next_pow2(traces_len * max_degree + 1)

    def betas_count(self) -> int:
        return self.count_of_evaluation_with_padding.bit_length() - 1

    def fft_points_count_F(self) -> int:
        return _next_pow2(self.betas_count() + 1)

    def fft_log_domain_size_G(self) -> int:
        return self.fft_points_count_G.bit_length() - 1

    def lagrange_domain(self) -> int:
        return self.instances_to_fold.bit_length() - 1

    def fft_log_domain_size_K(self) -> int:
        return _next_pow2(max(self.fft_points_count_G + 1 - self.instances_to_fold, 0))

    def fft_points_count_K(self) -> int:
        return 1 << self.fft_log_domain_size_K()


# ---------------------------------------------------------------------------------------------- device pieces
def _pg_tree(S: PlonkStructure, traces_W: Sequence[np.ndarray], coef: List[List[int]], challenges: List[List[int]], multipliers: List[List[int]],
             point_blend: List[int], log_leaves: int, row_mode: int) -> List[int]:
    lib = _lib.load()
    progs = S.gate_programs()
    gate_arr = (ctypes.c_void_p * len(progs))(*[p._h for p in progs])
    Ws = [np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4) for w in traces_W]
    wp = (_lib.u64p * len(Ws))(*[w.ctypes.data_as(_lib.u64p) for w in Ws])
    nb, npnt = len(coef), len(multipliers)
    coef_l = _to_mont([c for row in coef for c in row], M)
    nch = len(challenges[0]) if challenges and challenges[0] else 0
    ch_l = _to_mont([c for row in challenges for c in row], M) if nch else np.zeros((1, 4), dtype=np.uint64)
    mul_l = _to_mont([c for row in multipliers for c in row], M) if log_leaves else np.zeros((1, 4), dtype=np.uint64)
    pb = (ctypes.c_uint32 * npnt)(*point_blend)
    out = np.zeros((npnt, 4), dtype=np.uint64)
    _lib.check(
        lib.sb_pg_tree(gate_arr, len(progs), S._cols, S.num_advice_columns, wp, len(Ws), coef_l.ctypes.data_as(_lib.u64p), ch_l.ctypes.data_as(_lib.u64p),
                       nch, nb, row_mode, log_leaves, mul_l.ctypes.data_as(_lib.u64p), npnt, pb, out.ctypes.data_as(_lib.u64p))
    )
    return _from_mont(out)


def _ifft_ints(vals: List[int]) -> List[int]:
    a = _to_mont(vals, M)
    fft.ifft(a)
    return _from_mont(a)


def compute_F(ctx: PolyContext, betas: Sequence[int], delta: int, trace_W: np.ndarray, trace_challenges: Sequence[int], row_mode: int = ROW_COMPAT) -> List[int]:
    t = ctx.betas_count()
    betas = list(betas)[:t]
    assert len(betas) == t
    deltas = [delta % M]
    for _ in range(t - 1):
        deltas.append(deltas[-1] * deltas[-1] % M)
    Xs = iter_cyclic_subgroup(ctx.fft_points_count_F().bit_length() - 1)
    mult = [[(b + X * d) % M for b, d in zip(betas, deltas)] for X in Xs]
    evals = _pg_tree(ctx.S, [trace_W], [[1]], [list(trace_challenges)], mult, [0] * len(Xs), t, row_mode)
    return _ifft_ints(evals)


def beta_stroke(betas: Sequence[int], alpha: int, delta: int) -> List[int]:
    out, d = [], delta % M
    for b in betas:
        out.append((b + alpha * d) % M)
        d = d * d % M
    return out


def compute_G(ctx: PolyContext, betas_stroke: Sequence[int], acc_W: np.ndarray, acc_challenges: Sequence[int], traces_W: Sequence[np.ndarray],
              traces_challenges: Sequence[Sequence[int]], row_mode: int = ROW_COMPAT) -> List[int]:
    if not traces_W:
        raise ValueError("You can't fold 0 traces")
    t = ctx.betas_count()
    bs = list(betas_stroke)[:t]
    assert len(bs) == t
    points = iter_cyclic_subgroup(ctx.fft_log_domain_size_G())[: ctx.fft_points_count_G]
    Ls = [eval_lagrange_polys(X, ctx.lagrange_domain()) for X in points]
    all_ch = [list(acc_challenges)] + [list(c) for c in traces_challenges]
    J = len(all_ch)
    assert all(len(L) == J for L in Ls), "zip_eq"
    folded_ch = [[sum(L[j] * all_ch[j][i] for j in range(J)) % M for i in range(len(all_ch[0]))] for L in Ls]
    evals = _pg_tree(ctx.S, [acc_W] + list(traces_W), Ls, folded_ch, [bs] * len(points), list(range(len(points))), t, row_mode)
    return _ifft_ints(evals)


def _batch_inverse(vals: List[int]) -> List[int]:
    """Montgomery's trick over Python ints (all values non-zero): one modular inversion for the whole list."""
    pref, acc = [], 1
    for v in vals:
        pref.append(acc)
        acc = acc * v % M
    inv = pow(acc, -1, M)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * pref[i] % M
        inv = inv * vals[i] % M
    return out


def compute_K_from_G(ctx: PolyContext, poly_G: Sequence[int], poly_F_in_alpha: int, zeta: int = fft.FR_ZETA) -> List[int]:
    """poly/mod.rs:475-509: K on the coset zeta * H, K(X) = (G(X) - F(alpha) * L0(X)) / Z(X), then coset_ifft.  Same values as the
    point-by-point form (L0 from iter_eval_lagrange_poly_for_cyclic_group incl. its 0/0 -> 1 case, lagrange.rs:50-74; Z from
    eval_vanish_polynomial, :83-85); the 2 * #points inversions are shared (host glue, O(#points))."""
    n = ctx.instances_to_fold
    log_n = ctx.lagrange_domain()
    ninv = pow(1 << log_n, -1, M)
    Xs = [zeta * w % M for w in iter_cyclic_subgroup(ctx.fft_log_domain_size_K())]
    xn1 = [(pow(X, 1 << log_n, M) - 1) % M for X in Xs]      # X^(2^log_n) - 1: numerator of every Lagrange value
    Z = [(pow(X, n, M) - 1) % M for X in Xs]
    den = [(X - 1) % M for X in Xs]                            # first subgroup element is 1
    if any(d == 0 for d in den) or any(z == 0 for z in Z):
        vals = []                                              # degenerate point: fall back to the literal form
        for X in Xs:
            g = poly_eval(poly_G, X)
            L0 = eval_lagrange_polys(X, log_n)[0]
            vals.append((g - poly_F_in_alpha * L0) * pow(eval_vanish_polynomial(n, X), -1, M) % M)
    else:
        inv = _batch_inverse(den + Z)
        m = len(Xs)
        vals = [(poly_eval(poly_G, X) - poly_F_in_alpha * (ninv * x1 % M * inv[i] % M)) * inv[m + i] % M for i, (X, x1) in enumerate(zip(Xs, xn1))]
    a = _to_mont(vals, M)
    fft.coset_ifft(a, zeta)
    return _from_mont(a)


def compute_K(ctx, poly_F_in_alpha, betas_stroke, acc_W, acc_challenges, traces_W, traces_challenges, row_mode: int = ROW_COMPAT):
    return compute_K_from_G(ctx, compute_G(ctx, betas_stroke, acc_W, acc_challenges, traces_W, traces_challenges, row_mode), poly_F_in_alpha)


def evaluate_e_from_trace(S: PlonkStructure, trace_W: np.ndarray, trace_challenges: Sequence[int], betas: Sequence[int], row_mode: int = ROW_COMPAT) -> int:
    t = _next_pow2((1 << S.k) * len(S.gates)).bit_length() - 1
    return _pg_tree(S, [trace_W], [[1]], [list(trace_challenges)], [list(betas)[:t]], [0], t, row_mode)[0]


def fold_witness(acc_W: np.ndarray, incoming_Ws: Sequence[np.ndarray], lagrange_for_gamma: Sequence[int]) -> np.ndarray:
    lib = _lib.load()
    Ws = [np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4) for w in [acc_W] + list(incoming_Ws)]
    assert all(w.shape == Ws[0].shape for w in Ws), "zip_eq"
    coef = _to_mont(list(lagrange_for_gamma)[: len(Ws)], M)
    ptrs = (_lib.u64p * len(Ws))(*[w.ctypes.data_as(_lib.u64p) for w in Ws])
    out = np.zeros_like(Ws[0])
    _lib.check(lib.sb_lincomb(_lib.FIELD_FR, ptrs, coef.ctypes.data_as(_lib.u64p), len(Ws), Ws[0].shape[0], out.ctypes.data_as(_lib.u64p)))
    return out
