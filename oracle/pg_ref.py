"""Literal Python restatement of the Protogalaxy prover polynomials (reference src/nifs/protogalaxy).

TEST INFRASTRUCTURE ONLY.  Values are canonical Python ints over bn256 Fr (the only field with the FFT domain).

    get_evaluate_witness_fn                src/plonk/mod.rs:683-718   (incl. `index & total_row`)
    compute_F                              src/nifs/protogalaxy/poly/mod.rs:68-203
    PolyContext                            poly/mod.rs:205-269          (incl. fft_log_domain_size_K)
    FoldedWitness::new / fold_*            poly/folded_witness.rs:20-180
    compute_G, BetaStrokeIter, compute_K*  poly/mod.rs:308-509
    evaluate_e_from_trace                  src/nifs/protogalaxy/mod.rs:571-640
    ProtoGalaxy::fold_witness              src/nifs/protogalaxy/mod.rs:176-210
    lagrange helpers                       src/polynomial/lagrange.rs:22-85 (oracle/pyref.py)
    UnivariatePoly::eval                   src/polynomial/univariate.rs:67-75

`row_mode`: "compat" reproduces the reference bit for bit (every leaf of a gate block evaluates row 0),
"correct" uses `index % total_row`; the reference's own tests only compare the function with itself, so both
modes are restated and the CUDA path is checked in both.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

from . import expr_ref as E
from . import pyref as R

M = R.FR


def _tree_reduce(items: list, reducer: Callable):
    """itertools::Itertools::tree_reduce for a power-of-two number of items == a perfect binary tree."""
    assert items and (len(items) & (len(items) - 1)) == 0
    level = items
    while len(level) > 1:
        level = [reducer(level[i], level[i + 1]) for i in range(0, len(level), 2)]
    return level[0]


class PGStructure:
    """The slice of PlonkStructure Protogalaxy reads: individual gates, columns, k."""

    def __init__(self, k, selectors, fixed, num_advice, num_lookups, gates, num_challenges=0):
        self.k, self.selectors, self.fixed = k, selectors, fixed
        self.num_advice, self.num_lookups, self.gates = num_advice, num_lookups, gates
        self.num_challenges = num_challenges
        self.ctx = E.Ctx(len(selectors), len(fixed), num_advice, num_challenges, num_lookups)

    def count_of_evaluation_with_padding(self):
        cnt = (1 << self.k) * len(self.gates)
        p = 1
        while p < cnt:
            p <<= 1
        return p


def gate_degree(e, ctx: E.Ctx) -> int:
    """Expression::degree (expression.rs:431-450)."""
    t = e[0]
    if t == "C":
        return 0
    if t == "P":
        return 1 if ctx.is_fold_var(e[1]) else 0
    if t == "H":
        return 1
    if t in ("N", "X"):
        return gate_degree(e[1], ctx)
    if t == "S":
        return max(gate_degree(e[1], ctx), gate_degree(e[2], ctx))
    return gate_degree(e[1], ctx) + gate_degree(e[2], ctx)


class PolyContext:
    def __init__(self, S: PGStructure, traces_len: int):
        self.S = S
        self.count = S.count_of_evaluation_with_padding()
        self.instances_to_fold = traces_len + 1
        assert self.instances_to_fold & (self.instances_to_fold - 1) == 0
        max_degree = max([gate_degree(g, S.ctx) for g in S.gates] or [0])
        v = traces_len * max_degree + 1
        p = 1
        while p < v:
            p <<= 1
        self.fft_points_count_G = p

    def betas_count(self):
        return self.count.bit_length() - 1

    def fft_points_count_F(self):
        v, p = self.betas_count() + 1, 1
        while p < v:
            p <<= 1
        return p

    def fft_log_domain_size_G(self):
        return self.fft_points_count_G.bit_length() - 1

    def lagrange_domain(self):
        return self.instances_to_fold.bit_length() - 1

    def fft_log_domain_size_K(self):
        # poly/mod.rs:263-268: a POINT COUNT used as a log
        v = max(self.fft_points_count_G + 1 - self.instances_to_fold, 0)
        p = 1
        while p < v:
            p <<= 1
        return p


def evaluate_witness_fn(S: PGStructure, witness: Sequence[Sequence[int]], challenges: Sequence[int], row_mode: str):
    """get_evaluate_witness_fn (plonk/mod.rs:683-718): i -> gate_{i / 2^k}(row)."""
    n = 1 << S.k
    evs = [E.GraphEvaluator(g, M) for g in S.gates]
    nsel, nfix = len(S.selectors), len(S.fixed)

    def col(row, index):
        if index < nsel:
            return 1 if S.selectors[index][row] else 0
        if index < nsel + nfix:
            return S.fixed[index - nsel][row]
        return E.eval_advice_var(witness, S.num_advice, S.num_lookups, n, row, index - nsel - nfix) if len(witness) > 1 else witness[0][(index - nsel - nfix) * n + row]

    limit = len(evs) * n

    def f(i):
        if i >= limit:
            return 0
        g = i // n
        row = (i & n) if row_mode == "compat" else (i % n)
        # get_rotation_idx wraps row == 2^k back into range (graph_evaluator.rs:51-53)
        return evs[g].evaluate(lambda r, idx: col(r % n, idx), challenges, row, n)

    return f


def compute_F(ctx: PolyContext, betas: Sequence[int], delta: int, witness, challenges, row_mode="compat") -> List[int]:
    t = ctx.betas_count()
    betas = list(betas)[:t]
    assert len(betas) == t
    deltas = [delta]
    for _ in range(t - 1):
        deltas.append(deltas[-1] * deltas[-1] % M)
    npts = ctx.fft_points_count_F()
    Xs = list(R.iter_cyclic_subgroup(npts.bit_length() - 1))
    cp = [[(b + X * d) % M for b, d in zip(betas, deltas)] for X in Xs]
    f = evaluate_witness_fn(ctx.S, witness, challenges, row_mode)

    def reducer(l, r):
        if l[0] == "leaf":
            return ("calc", [(l[1] + r[1] * c[0]) % M for c in cp], 1)
        h = l[2]
        assert r[2] == h
        return ("calc", [(a + b * c[h]) % M for a, b, c in zip(l[1], r[1], cp)], h + 1)

    node = _tree_reduce([("leaf", f(i)) for i in range(ctx.count)], reducer)
    return R.ifft(node[1])


def lagrange_all(X: int, log_n: int) -> List[int]:
    return R.eval_lagrange_polys(log_n, X)


def folded_witnesses(points, lagrange_domain, acc_w, acc_ch, traces_w, traces_ch):
    """FoldedWitness::new: for each X in points, sum_j L_j(X) * (witness_j, challenges_j)."""
    out = []
    all_w = [acc_w] + list(traces_w)
    all_c = [acc_ch] + list(traces_ch)
    for X in points:
        L = lagrange_all(X, lagrange_domain)[: len(all_w)] if len(all_w) <= (1 << lagrange_domain) else None
        W = [[sum(L[j] * all_w[j][r][i] for j in range(len(all_w))) % M for i in range(len(acc_w[r]))] for r in range(len(acc_w))]
        C = [sum(L[j] * all_c[j][i] for j in range(len(all_c))) % M for i in range(len(acc_ch))]
        out.append((W, C))
    return out


def compute_G(ctx: PolyContext, betas_stroke: Sequence[int], acc_w, acc_ch, traces_w, traces_ch, row_mode="compat") -> List[int]:
    t = ctx.betas_count()
    bs = list(betas_stroke)[:t]
    points = list(R.iter_cyclic_subgroup(ctx.fft_log_domain_size_G()))[: ctx.fft_points_count_G]
    folded = folded_witnesses(points, ctx.lagrange_domain(), acc_w, acc_ch, traces_w, traces_ch)
    fns = [evaluate_witness_fn(ctx.S, W, C, row_mode) for W, C in folded]

    def reducer(l, r):
        h = l[1]
        assert r[1] == h
        return ([(a + b * bs[h]) % M for a, b in zip(l[0], r[0])], h + 1)

    node = _tree_reduce([([fn(i) for fn in fns], 0) for i in range(ctx.count)], reducer)
    return R.ifft(node[0])


def beta_stroke(betas, alpha, delta):
    out, d = [], delta
    for b in betas:
        out.append((b + alpha * d) % M)
        d = d * d % M
    return out


def poly_eval(coeffs, x):
    """UnivariatePoly::eval (univariate.rs:67-75)."""
    acc, p = 0, 1
    for c in coeffs:
        acc = (acc + p * c) % M
        p = p * x % M
    return acc


def compute_K_from_G(ctx: PolyContext, poly_G, F_in_alpha, zeta=R.FR_ZETA):
    logK = ctx.fft_log_domain_size_K()
    vals = []
    for w in R.iter_cyclic_subgroup(logK):
        X = zeta * w % M
        g = poly_eval(poly_G, X)
        L0 = lagrange_all(X, ctx.lagrange_domain())[0]
        Z = (pow(X, ctx.instances_to_fold, M) - 1) % M
        vals.append((g - F_in_alpha * L0) * pow(Z, -1, M) % M)
    return R.coset_ifft(vals, zeta)


def evaluate_e(S: PGStructure, witness, challenges, betas, row_mode="compat") -> int:
    cnt = S.count_of_evaluation_with_padding()
    f = evaluate_witness_fn(S, witness, challenges, row_mode)

    def reducer(l, r):
        assert l[1] == r[1]
        return ((l[0] + r[0] * betas[r[1]]) % M, l[1] + 1)

    return _tree_reduce([(f(i), 0) for i in range(cnt)], reducer)[0]


def fold_witness(acc_w, incoming_ws, lagrange_for_gamma):
    """ProtoGalaxy::fold_witness (mod.rs:176-210)."""
    l0 = lagrange_for_gamma[0]
    out = [[w * l0 % M for w in r] for r in acc_w]
    for w, ln in zip(incoming_ws, lagrange_for_gamma[1:]):
        for ro, ri in zip(out, w):
            assert len(ro) == len(ri)
            for i in range(len(ro)):
                ro[i] = (ro[i] + ri[i] * ln) % M
    return out
