"""CPU restatement of the reference's SPS lookup columns, permutation decider and witness assembly
(SURVEY 8f-3 / 8f-4).  TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing under sirius_b200/).

Follows, literally and on Python integers (canonical field values, not Montgomery limbs):

    lookup::Arguments::{compress_from, vanishing_lookup_polys, log_derivative_expr,
                        log_derivative_lhs_and_rhs, evaluate_ls, evaluate_ts, evaluate_m,
                        evaluate_h_g, evaluate_coefficient_1}            src/plonk/lookup.rs:87-343
    ArgumentCoefficient1::evaluate_coefficient_2                        src/plonk/lookup.rs:351-365
    LookupEvalDomain::eval_advice_var                                    src/plonk/eval.rs:112-131
    PlonkStructure::is_sat_log_derivative                                src/plonk/mod.rs:363-397
    PlonkStructure::run_sps_protocol_{2,3} (layout of the round vectors) src/plonk/mod.rs:501-660
    sparse::matrix_multiply                                              src/polynomial/sparse.rs:7-20
    VanillaFS::is_sat_permutation (Z assembly + mismatch count)          src/nifs/sangria/mod.rs:385-453
    util::{concatenate_with_padding, batch_invert_assigned}              src/util/mod.rs:120-153,214-218

Pinned by: the reference's own `concatenate_with_padding` unit tests (src/util/mod.rs:235-290, restated in
tests/test_oracle_lookup.py) and the structural invariant the reference's end-to-end tests rely on -- a witness
produced by the SPS protocol satisfies the compressed gate+lookup relation on every row and the log-derivative
sum check.  No known-answer vector for l/t/m/h/g exists in the reference.

Reference quirks reproduced on purpose:
  Q1  run_sps_protocol_* lay the lookup columns out as concat(ls, ts, ms) / concat(hs, gs) (mod.rs:523,551,619,637)
      while eval_advice_var reads them interleaved per lookup (eval.rs:178-194) and is_sat_log_derivative gathers
      h at even and g at odd column positions (mod.rs:378-383).  The three agree only for ONE lookup argument.
  Q2  evaluate_m assigns the whole multiplicity to the first row of a repeated table value (lookup.rs:283-297).
  Q3  concatenate_with_padding pads but never truncates (pad_using, util/mod.rs:216).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

from . import expr_ref as X


# ------------------------------------------------------------------------------------------ util/mod.rs


def concatenate_with_padding(vs: Sequence[Sequence[int]], pad_size: int) -> List[int]:
    """util/mod.rs:214-218"""
    out: List[int] = []
    for v in vs:
        out.extend(v)
        if len(v) < pad_size:
            out.extend([0] * (pad_size - len(v)))
    return out


def batch_invert_assigned(assigned: Sequence[Sequence[tuple]], m: int) -> List[List[int]]:
    """util/mod.rs:128-153.  Assigned values are ("zero",), ("trivial", x) or ("rational", num, den)
    (halo2 `Assigned`): numerator() * denominator()^-1, a zero denominator inverting to zero (BatchInvert skips it)."""
    out = []
    for col in assigned:
        res = []
        for a in col:
            if a[0] == "zero":
                res.append(0)
            elif a[0] == "trivial":
                res.append(a[1] % m)
            else:
                den = a[2] % m
                res.append(a[1] * (pow(den, m - 2, m) if den else 0) % m)
        out.append(res)
    return out


# ------------------------------------------------------------------------------------------ polynomial/sparse.rs


def matrix_multiply(P: Sequence[Tuple[int, int, int]], Z: Sequence[int], m: int) -> List[int]:
    """sparse.rs:7-20"""
    res = [0] * len(Z)
    for row, col, value in P:
        if col < len(Z):
            res[row] = (res[row] + value * Z[col]) % m
        else:
            raise RuntimeError("invalid matrix multiply")
    return res


def permutation_mismatch_count(P, instances_flat: Sequence[int], W0: Sequence[int], k: int, num_advice: int, m: int) -> int:
    """is_sat_permutation (nifs/sangria/mod.rs:420-452): Z = instances ++ W[0][.. 2^k * num_advice]; #{y != z}."""
    Z = list(instances_flat) + list(W0[: (1 << k) * num_advice])
    Y = matrix_multiply(P, Z, m)
    return sum(1 for y, z in zip(Y, Z) if (y - z) % m != 0)


# ------------------------------------------------------------------------------------------ plonk/lookup.rs


class Arguments:
    """lookup.rs:72-83.  Built from per-argument lists of (already converted) input / table expressions."""

    def __init__(self, inputs: Sequence[Sequence[tuple]], tables: Sequence[Sequence[tuple]]):
        # compress_from (lookup.rs:87-131); compress_halo2_expression == compress_expression on converted
        # expressions (plonk/util.rs:12-33), challenge_index 0
        max_len = max((len(a) for a in inputs), default=0)
        if max_len == 0:
            raise ValueError("compress_from -> None")
        self.has_vector_lookup = max_len > 1
        self.lookup_polys = [self._compress(a) for a in inputs]
        self.table_polys = [self._compress(t) for t in tables]

    @staticmethod
    def _compress(exprs):
        if len(exprs) > 1:
            return X.compress_expression(list(exprs), 0)
        return exprs[0]

    def num_lookups(self) -> int:
        return len(self.lookup_polys)

    def vanishing_lookup_polys(self, num_selectors, num_fixed, num_advice):
        """lookup.rs:141-170"""
        off = num_selectors + num_fixed + num_advice
        ls = [X.Sub(L, X.Poly(off + i * 5)) for i, L in enumerate(self.lookup_polys)]
        ts = [X.Sub(T, X.Poly(off + i * 5 + 1)) for i, T in enumerate(self.table_polys)]
        return ls + ts

    def log_derivative_expr(self, num_selectors, num_fixed, num_advice, lookup_index, challenge_index):
        """lookup.rs:178-200"""
        r = X.Chal(challenge_index)
        off = num_selectors + num_fixed + num_advice
        l, t, mm, h, g = [X.Poly(off + lookup_index * 5 + i) for i in range(5)]
        lhs = X.Sub(X.Mul(h, X.Sum(l, r)), X.Const(1))
        rhs = X.Sub(X.Mul(g, X.Sum(t, r)), mm)
        return lhs, rhs

    def log_derivative_lhs_and_rhs(self, num_selectors, num_fixed, num_advice):
        """lookup.rs:203-212"""
        ci = 1 if self.has_vector_lookup else 0
        out = []
        for i in range(self.num_lookups()):
            out.extend(self.log_derivative_expr(num_selectors, num_fixed, num_advice, i, ci))
        return out

    def to_expressions(self, num_selectors, num_fixed, num_advice):
        """lookup.rs:133-137"""
        return self.vanishing_lookup_polys(num_selectors, num_fixed, num_advice) + self.log_derivative_lhs_and_rhs(
            num_selectors, num_fixed, num_advice
        )

    # -- evaluate_ls / evaluate_ts (lookup.rs:216-268) over a LookupEvalDomain (eval.rs:84-131)
    def _evaluate(self, polys, k, selectors, fixed, advice, r, m):
        n = 1 << k
        nsel, nfix = len(selectors), len(fixed)

        def eval_column_var(row, index):
            if index < nsel:
                return 1 if selectors[index][row] else 0
            if index < nsel + nfix:
                return fixed[index - nsel][row]
            a = index - nsel - nfix
            if a >= len(advice):
                raise IndexError("ColumnVariableIndexOutOfBoundary")
            return advice[a][row]

        out = []
        for p in polys:
            ev = X.GraphEvaluator(p, m)
            out.append([ev.evaluate(eval_column_var, [r], row, n) for row in range(n)])
        return out

    @staticmethod
    def evaluate_m(l: Sequence[int], t: Sequence[int]) -> List[int]:
        """lookup.rs:270-298"""
        processed = set()
        counts = {}
        for v in l:
            counts[v] = counts.get(v, 0) + 1
        out = []
        for v in t:
            if v in processed:
                out.append(0)
            else:
                processed.add(v)
                out.append(counts.get(v, 0))
        return out

    @staticmethod
    def evaluate_h_g(l, t, r, mm, m):
        """lookup.rs:300-312"""
        inv0 = lambda x: pow(x % m, m - 2, m) if x % m else 0
        h = [inv0(li + r) for li in l]
        assert len(t) == len(mm)
        g = [mi * inv0(ti + r) % m for ti, mi in zip(t, mm)]
        return h, g

    def evaluate_coefficient_1(self, k, selectors, fixed, advice, r, m):
        """lookup.rs:314-336 -> (ls, ts, ms)"""
        ls = self._evaluate(self.lookup_polys, k, selectors, fixed, advice, r, m)
        ts = self._evaluate(self.table_polys, k, selectors, fixed, advice, r, m)
        ms = [self.evaluate_m(l, t) for l, t in zip(ls, ts)]
        return ls, ts, ms

    @staticmethod
    def evaluate_coefficient_2(ls, ts, ms, r, m):
        """lookup.rs:351-365 -> (hs, gs)"""
        hs, gs = [], []
        for l, t, mm in zip(ls, ts, ms):
            h, g = Arguments.evaluate_h_g(l, t, r, mm, m)
            hs.append(h)
            gs.append(g)
        return hs, gs


def is_sat_log_derivative(W: Sequence[Sequence[int]], k: int, num_lookups: int, has_vector_lookup: bool, m: int) -> bool:
    """plonk/mod.rs:363-397 (h gathered at column positions 0,2,4.. and g at 1,3,5.., see Q1)."""
    n = 1 << k

    def gather(Wr, start):
        return [Wr[idx * n : idx * n + n] for idx in range(start, start + 2 * num_lookups, 2)]

    def check(hs, gs):
        return all(sum((hi - gi) for hi, gi in zip(h, g)) % m == 0 for h, g in zip(hs, gs))

    if has_vector_lookup:
        return check(gather(W[2], 0), gather(W[2], 1))
    if num_lookups > 0:
        return check(gather(W[1], 0), gather(W[1], 1))
    return True


def run_sps_protocol(args: Optional[Arguments], k, selectors, fixed, advice, m, challenge: Callable[[int, List[int]], int]):
    """The witness-round layout of run_sps_protocol_{0/1,2,3} (plonk/mod.rs:431-660).  `challenge(round, W_round)`
    stands in for commit -> absorb -> squeeze (the random oracle is host-side and out of scope here).
    Returns (W rounds, challenges)."""
    n = 1 << k
    if args is None:
        return [concatenate_with_padding(advice, n)], []
    if not args.has_vector_lookup:
        ls, ts, ms = args.evaluate_coefficient_1(k, selectors, fixed, advice, 0, m)   # r = F::ZERO (mod.rs:516)
        W1 = concatenate_with_padding(advice, n) + concatenate_with_padding(ls + ts + ms, n)
        r1 = challenge(0, W1)
        hs, gs = args.evaluate_coefficient_2(ls, ts, ms, r1, m)
        W2 = concatenate_with_padding(hs + gs, n)
        r2 = challenge(1, W2)
        return [W1, W2], [r1, r2]
    W1 = concatenate_with_padding(advice, n)
    r1 = challenge(0, W1)
    ls, ts, ms = args.evaluate_coefficient_1(k, selectors, fixed, advice, r1, m)
    W2 = concatenate_with_padding(ls + ts + ms, n)
    r2 = challenge(1, W2)
    hs, gs = args.evaluate_coefficient_2(ls, ts, ms, r2, m)
    W3 = concatenate_with_padding(hs + gs, n)
    r3 = challenge(2, W3)
    return [W1, W2, W3], [r1, r2, r3]
