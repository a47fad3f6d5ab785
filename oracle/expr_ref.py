"""Literal Python restatement of the reference's expression machinery and Sangria NIFS prover steps.

TEST INFRASTRUCTURE ONLY (see oracle/sirius_oracle.c header).  Nothing here is imported by sirius_b200/.

What is restated, with the reference lines it follows:
    Expression, Display/visualize           src/polynomial/expression.rs:112-120, 262-303
    Expression::homogeneous                 src/polynomial/expression.rs:356-429 (+ challenge_in_degree :503-515)
    GroupedPoly::{new, add, sub, mul, neg}  src/polynomial/grouped_poly.rs:58-110, 140-290
    compress_expression                     src/plonk/util.rs:35-55
    CompressedGates                         src/plonk/mod.rs:68-121
    GraphEvaluator::{new, evaluate}         src/polynomial/graph_evaluator.rs:91-150, 196-388
    PlonkEvalDomain column addressing       src/plonk/eval.rs:57-69, 153-228
    VanillaFS::commit_cross_terms (eval)    src/nifs/sangria/mod.rs:110-147
    RelaxedPlonkWitness::fold               src/nifs/sangria/accumulator.rs:363-404
    MainGate gate polynomial                src/main_gate.rs:535-583

Pinned by the reference's string known-answer tests: grouped_poly.rs:294-461 (`mul`, `creation`),
main_gate.rs:892-926 (`test_main_gate_expr`, `test_main_gate_cross_term`) -- see tests/test_oracle_expr.py.
All values are canonical Python ints mod the field.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import pyref

# ------------------------------------------------------------------------------------------------ Expression
# tuples: ("C", v) ("P", index, rot) ("H", index) ("N", a) ("S", a, b) ("M", a, b) ("X", a, k)


def Const(v):
    return ("C", v)


def Poly(index, rot=0):
    return ("P", index, rot)


def Chal(index):
    return ("H", index)


def Neg(a):
    return ("N", a)


def Sum(a, b):
    return ("S", a, b)


def Sub(a, b):  # expression.rs:498  a - b = Sum(a, Negated(b))
    return ("S", a, ("N", b))


def Mul(a, b):
    return ("M", a, b)


def Scaled(a, k):
    return ("X", a, k)


def _hex(v: int) -> str:
    # `{:?}` of a field element is 0x + 64 hex digits; trim_leading_zeros (src/util/mod.rs:185-189) strips zeros
    return "0x" + format(v, "x").lstrip("0") if v else "0x"


def visualize(e) -> str:
    t = e[0]
    if t == "C":
        return _hex(e[1])
    if t == "P":
        rot = e[2]
        r = "" if rot == 0 else (f"[{rot}]" if rot < 0 else f"[+{rot}]")
        return f"Z_{e[1]}{r}"
    if t == "H":
        return f"r_{e[1]}"
    if t == "N":
        return f"-{visualize(e[1])}"
    if t == "S":
        if e[2][0] == "N":
            return f"{visualize(e[1])} - {visualize(e[2][1])}"
        return f"{visualize(e[1])} + {visualize(e[2])}"
    if t == "M":
        l = f"({visualize(e[1])})" if e[1][0] == "S" else visualize(e[1])
        r = f"({visualize(e[2])})" if e[2][0] == "S" else visualize(e[2])
        return f"{l} * {r}"
    if t == "X":
        return f'"{_hex(e[2])}" * {visualize(e[1])}'
    raise ValueError(t)


class Ctx:
    """QueryIndexContext (expression.rs:38-70)."""

    def __init__(self, num_selectors=0, num_fixed=0, num_advice=0, num_challenges=0, num_lookups=0):
        self.num_selectors, self.num_fixed, self.num_advice = num_selectors, num_fixed, num_advice
        self.num_challenges, self.num_lookups = num_challenges, num_lookups

    def num_fold_vars(self):
        return self.num_advice + 5 * self.num_lookups

    def is_fold_var(self, index):
        return index >= self.num_selectors + self.num_fixed

    def copy(self):
        return Ctx(self.num_selectors, self.num_fixed, self.num_advice, self.num_challenges, self.num_lookups)


def challenges_of(e, out: set):
    t = e[0]
    if t == "H":
        out.add(e[1])
    elif t in ("N", "X"):
        challenges_of(e[1], out)
    elif t in ("S", "M"):
        challenges_of(e[1], out)
        challenges_of(e[2], out)
    return out


def num_challenges(e) -> int:
    return len(challenges_of(e, set()))


def challenge_in_degree(idx, degree):
    ch = Chal(idx)
    res = ch
    for _ in range(2, degree + 1):
        res = Mul(res, ch)
    return res


def homogeneous(e, ctx: Ctx):
    new = ctx.num_challenges
    t = e[0]
    if t == "C":
        return e, 0
    if t == "P":
        return e, (1 if ctx.is_fold_var(e[1]) else 0)
    if t == "H":
        return e, 1
    if t == "N":
        x, d = homogeneous(e[1], ctx)
        return Neg(x), d
    if t == "S":
        (l, ld), (r, rd) = homogeneous(e[1], ctx), homogeneous(e[2], ctx)
        if ld > rd:
            return Sum(l, Mul(r, challenge_in_degree(new, ld - rd))), ld
        if ld < rd:
            return Sum(Mul(l, challenge_in_degree(new, rd - ld)), r), rd
        return Sum(l, r), ld
    if t == "M":
        (l, ld), (r, rd) = homogeneous(e[1], ctx), homogeneous(e[2], ctx)
        return Mul(l, r), ld + rd
    if t == "X":
        x, d = homogeneous(e[1], ctx)
        return Scaled(x, e[2]), d
    raise ValueError(t)


def compress_expression(exprs, challenge_index):
    y = Chal(challenge_index)
    if len(exprs) > 1:
        acc = Const(0)
        for ex in exprs:
            acc = Sum(ex, Mul(acc, y))
        return acc
    return exprs[0] if exprs else Const(0)


# ------------------------------------------------------------------------------------------------ GroupedPoly
# terms: List[Optional[expr]] indexed by degree


def gp_new(e, ctx: Ctx) -> List[Optional[tuple]]:
    t = e[0]
    if t == "C":
        return [e]
    if t == "P":
        terms = [e]
        if ctx.is_fold_var(e[1]):  # Advice and Lookup both shift by num_fold_vars (expression.rs:62-68)
            terms.append(Poly(e[1] + ctx.num_fold_vars(), e[2]))
        return terms
    if t == "H":
        return [e, Chal(e[1] + ctx.num_challenges)]
    if t == "N":
        return gp_neg(gp_new(e[1], ctx))
    if t == "S":
        return gp_add(gp_new(e[1], ctx), gp_new(e[2], ctx))
    if t == "M":
        return gp_mul(gp_new(e[1], ctx), gp_new(e[2], ctx))
    if t == "X":
        return [None if x is None else Mul(Const(e[2]), x) for x in gp_new(e[1], ctx)]
    raise ValueError(t)


def gp_neg(a):
    return [None if x is None else Neg(x) for x in a]


def _gp_zip(a, b, rhs_map, both):
    out = []
    for i in range(max(len(a), len(b))):
        l = a[i] if i < len(a) else None
        r = b[i] if i < len(b) else None
        if l is not None and r is not None:
            out.append(both(l, rhs_map(r)))
        elif r is not None:
            out.append(rhs_map(r))
        elif l is not None:
            out.append(l)
        else:
            out.append(None)
    return out


def gp_add(a, b):
    return _gp_zip(a, b, lambda r: r, lambda l, r: Sum(l, r))


def gp_sub(a, b):
    return _gp_zip(a, b, lambda r: Neg(r), lambda l, r: Sum(l, r))


def gp_mul(a, b):
    # grouped_poly.rs:216-268: the longer operand is `lhs` (ties: other, self)
    if len(a) <= len(b):
        lhs, rhs = b, a
    else:
        lhs, rhs = a, b
    res: List[Optional[tuple]] = []
    rhs_terms = [(d, x) for d, x in enumerate(rhs) if x is not None][::-1]
    for ld, lx in [(d, x) for d, x in enumerate(lhs) if x is not None][::-1]:
        for rd, rx in rhs_terms:
            deg = ld + rd
            ex = Mul(lx, rx)
            if deg >= len(res):
                res.extend([None] * (deg + 1 - len(res)))
            res[deg] = ex if res[deg] is None else Sum(res[deg], ex)
    return res


class CompressedGates:
    def __init__(self, gates, ctx: Ctx):
        ctx = ctx.copy()
        self.compressed = compress_expression(gates, ctx.num_challenges)
        ctx.num_challenges = num_challenges(self.compressed)
        self.homogeneous, self.degree = homogeneous(self.compressed, ctx)
        ctx.num_challenges = num_challenges(self.homogeneous)
        self.ctx = ctx
        self._grouped = None

    def grouped(self):
        if self._grouped is None:
            self._grouped = gp_new(self.homogeneous, self.ctx)
        return self._grouped


# ------------------------------------------------------------------------------------------------ GraphEvaluator

VS_CONSTANT, VS_INTERMEDIATE, VS_FIXED, VS_POLY, VS_CHALLENGE = range(5)
OP_ADD, OP_SUB, OP_MUL, OP_SQUARE, OP_DOUBLE, OP_NEGATE, OP_HORNER, OP_STORE = range(8)


class GraphEvaluator:
    def __init__(self, expr, modulus: int):
        self.m = modulus
        self.constants = [0, 1, 2]
        self.rotations: List[int] = []
        self.calcs: List[Tuple[int, tuple, Optional[tuple], int]] = []
        self._index: Dict[tuple, int] = {}
        self.num_intermediates = 0
        vs = self._expr(expr)
        self._calc(OP_STORE, vs, None)

    def _rot(self, r):
        if r in self.rotations:
            return self.rotations.index(r)
        self.rotations.append(r)
        return len(self.rotations) - 1

    def _const(self, c):
        c %= self.m
        if c in self.constants:
            return (VS_CONSTANT, self.constants.index(c), 0)
        self.constants.append(c)
        return (VS_CONSTANT, len(self.constants) - 1, 0)

    def _calc(self, op, a, b):
        key = (op, a, b)
        if key in self._index:  # `find` of an equal calculation (graph_evaluator.rs:241-258)
            return (VS_INTERMEDIATE, self._index[key], 0)
        target = self.num_intermediates
        self.calcs.append((op, a, b, target))
        self._index[key] = target
        self.num_intermediates += 1
        return (VS_INTERMEDIATE, target, 0)

    def _expr(self, e):
        Z, O, T2 = (VS_CONSTANT, 0, 0), (VS_CONSTANT, 1, 0), (VS_CONSTANT, 2, 0)
        t = e[0]
        if t == "C":
            return self._const(e[1])
        if t == "P":
            return self._calc(OP_STORE, (VS_POLY, e[1], self._rot(e[2])), None)
        if t == "H":
            return self._calc(OP_STORE, (VS_CHALLENGE, e[1], 0), None)
        if t == "N":
            if e[1][0] == "C":
                return self._const(-e[1][1])
            ra = self._expr(e[1])
            return ra if ra == Z else self._calc(OP_NEGATE, ra, None)
        if t == "S":
            if e[2][0] == "N":
                ra, rb = self._expr(e[1]), self._expr(e[2][1])
                if ra == Z:
                    return self._calc(OP_NEGATE, rb, None)
                if rb == Z:
                    return ra
                return self._calc(OP_SUB, ra, rb)
            ra, rb = self._expr(e[1]), self._expr(e[2])
            return self._calc(OP_ADD, ra, rb) if ra <= rb else self._calc(OP_ADD, rb, ra)
        if t == "M":
            ra, rb = self._expr(e[1]), self._expr(e[2])
            if ra == Z or rb == Z:
                return Z
            if ra == O:
                return rb
            if rb == O:
                return ra
            if ra == T2:
                return self._calc(OP_DOUBLE, rb, None)
            if rb == T2:
                return self._calc(OP_DOUBLE, ra, None)
            if ra == rb:
                return self._calc(OP_SQUARE, ra, None)
            return self._calc(OP_MUL, ra, rb) if ra <= rb else self._calc(OP_MUL, rb, ra)
        if t == "X":
            f = e[2] % self.m
            if f == 0:
                return Z
            if f == 1:
                return self._expr(e[1])
            cst = self._const(f)
            ra = self._expr(e[1])
            return self._calc(OP_MUL, ra, cst)
        raise ValueError(t)

    # -- literal evaluate (graph_evaluator.rs:361-388) on Python ints, `getter(row, index)` = eval_column_var
    def evaluate(self, eval_column_var, challenges: Sequence[int], row: int, num_rows: int) -> int:
        m = self.m
        rots = [(row + r) % num_rows for r in self.rotations]
        inter = [0] * self.num_intermediates

        def val(vs):
            k, i, r = vs
            if k == VS_CONSTANT:
                return self.constants[i]
            if k == VS_INTERMEDIATE:
                return inter[i]
            if k == VS_POLY:
                return eval_column_var(rots[r], i)
            if k == VS_CHALLENGE:
                return challenges[i]
            raise ValueError(k)

        for op, a, b, target in self.calcs:
            if op == OP_ADD:
                v = (val(a) + val(b)) % m
            elif op == OP_SUB:
                v = (val(a) - val(b)) % m
            elif op == OP_MUL:
                v = val(a) * val(b) % m
            elif op == OP_SQUARE:
                v = val(a) ** 2 % m
            elif op == OP_DOUBLE:
                v = 2 * val(a) % m
            elif op == OP_NEGATE:
                v = (-val(a)) % m
            else:
                v = val(a)
            inter[target] = v
        return inter[self.calcs[-1][3]] if self.calcs else 0

    def to_arrays(self):
        """(calcs int32 [n,8], constants uint64 [c,4] Montgomery, rotations int32) for the C interpreter."""
        arr = np.zeros((len(self.calcs), 8), dtype=np.int32)
        for i, (op, a, b, target) in enumerate(self.calcs):
            bb = b if b is not None else (0, 0, 0)
            arr[i] = [op, a[0], a[1], a[2], bb[0], bb[1], bb[2], target]
        return arr, pyref.to_mont_limbs(self.constants, self.m), np.array(self.rotations, dtype=np.int32)


# ------------------------------------------------------------------------------------------------ gates


def main_gate_expression(T, fixed_base, advice_base, num_selectors, num_fixed_total):
    """main_gate.rs:535-583 through Expression::from_halo2_expr (expression.rs:305-340)."""
    fx = lambda j: Poly(num_selectors + fixed_base + j)  # noqa: E731
    ad = lambda j: Poly(num_selectors + num_fixed_total + advice_base + j)  # noqa: E731
    state = [ad(i) for i in range(T)]
    inp, out = ad(T), ad(T + 1)
    q_1 = [fx(i) for i in range(T)]
    q_5 = [fx(T + i) for i in range(T)]
    q_m = [fx(2 * T), fx(2 * T + 1)]
    q_i, q_o, rc = fx(2 * T + 2), fx(2 * T + 3), fx(2 * T + 4)

    def pow_5(v):
        v2 = Mul(v, v)
        return Mul(Mul(v2, v2), v)

    init = Sum(Sum(Sum(Mul(Mul(q_m[0], state[0]), state[1]), Mul(q_i, inp)), rc), Mul(q_o, out))
    if T >= 4:
        init = Sum(Mul(Mul(q_m[1], state[2]), state[3]), init)
    acc = init
    for s, q1, q5 in zip(state, q_1, q_5):
        acc = Sum(acc, Sum(Mul(q1, s), Mul(q5, pow_5(s))))
    return acc


# ------------------------------------------------------------------------------------------------ Sangria prover steps


def eval_advice_var(Ws: Sequence[Sequence[int]], num_advice, num_lookup, row_size, row, index):
    """PlonkEvalDomain index_map (eval.rs:170-204) for ONE instance's round vectors."""
    if index < num_advice:
        i, j = 0, index
    else:
        li, sub = divmod(index - num_advice, 5)
        first = sub < 3
        if not first:
            sub -= 3
        nw = len(Ws)
        if nw == 2:
            i, j = (0, num_advice + li * 3 + sub) if first else (1, li * 2 + sub)
        elif nw == 3:
            i, j = (1, li * 3 + sub) if first else (2, li * 2 + sub)
        else:
            raise IndexError("InvalidWitnessIndex")
    return Ws[i][j * row_size + row]


class Structure:
    """The slice of PlonkStructure (src/plonk/mod.rs:127-193) the cross-term computation reads."""

    def __init__(self, k, selectors, fixed, num_advice, num_lookups, gates: CompressedGates, modulus):
        self.k, self.selectors, self.fixed = k, selectors, fixed
        self.num_advice, self.num_lookups, self.gates, self.m = num_advice, num_lookups, gates, modulus


def commit_cross_terms_eval(S: Structure, U1_challenges, U1_u, W1, U2_challenges, W2) -> List[List[int]]:
    """The evaluation span of VanillaFS::commit_cross_terms (src/nifs/sangria/mod.rs:110-147), literally:
    one GraphEvaluator per degree-grouped expression, rows 0..2^k, None -> zeros."""
    n = 1 << S.k
    challenges = list(U1_challenges) + [U1_u] + list(U2_challenges) + [1]
    nsel, nfix = len(S.selectors), len(S.fixed)
    max_width = S.num_advice + 5 * S.num_lookups

    def eval_column_var(row, index):
        if index < nsel:
            return 1 if S.selectors[index][row] else 0
        if index < nsel + nfix:
            return S.fixed[index - nsel][row]
        a = index - nsel - nfix
        if a < max_width:
            return eval_advice_var(W1, S.num_advice, S.num_lookups, n, row, a)
        return eval_advice_var(W2, S.num_advice, S.num_lookups, n, row, a - max_width)

    out = []
    for ex in S.gates.grouped()[1:]:  # iter_from_first
        if ex is None:
            out.append([0] * n)
            continue
        ev = GraphEvaluator(ex, S.m)
        out.append([ev.evaluate(eval_column_var, challenges, row, n) for row in range(n)])
    return out


def fold_witness(m, W1, E1, W2, cross_terms, r):
    """RelaxedPlonkWitness::fold (accumulator.rs:363-404)."""
    W = [[(a + r * b) % m for a, b in zip(v1, v2)] for v1, v2 in zip(W1, W2)]
    powers = []
    cur = r % m
    for _ in cross_terms:
        powers.append(cur)
        cur = cur * r % m
    E = []
    for i, ei in enumerate(E1):
        acc = ei
        for tk, pw in zip(cross_terms, powers):
            acc = (acc + pw * tk[i]) % m
        E.append(acc)
    return W, E


# ------------------------------------------------------------------------------------------------ C interpreter bridge


def c_graph_evaluate(field: int, ev: GraphEvaluator, selectors: np.ndarray, fixed: np.ndarray, adv_cols: Sequence[np.ndarray],
                     challenges: np.ndarray, log_rows: int, threads: int = 0) -> np.ndarray:
    """Evaluate `ev` for every row with the C interpreter (oracle/sirius_oracle.c so_graph_evaluate).
    adv_cols[i] is the uint64 [n,4] column bound to fold-variable index i (both instances concatenated)."""
    from . import lib

    calcs, consts, rots = ev.to_arrays()
    n = 1 << log_rows
    u64p = ctypes.POINTER(ctypes.c_uint64)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    sel = [np.ascontiguousarray(s, dtype=np.uint8) for s in selectors]
    fx = [np.ascontiguousarray(f, dtype=np.uint64) for f in fixed]
    ad = [np.ascontiguousarray(a, dtype=np.uint64) for a in adv_cols]
    sel_ptrs = (u8p * max(1, len(sel)))(*[s.ctypes.data_as(u8p) for s in sel])
    fx_ptrs = (u64p * max(1, len(fx)))(*[f.ctypes.data_as(u64p) for f in fx])
    ad_ptrs = (u64p * max(1, len(ad)))(*[a.ctypes.data_as(u64p) for a in ad])
    ch = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((n, 4), dtype=np.uint64)
    rc = lib().so_graph_evaluate(
        ctypes.c_int(field), calcs.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_size_t(len(calcs)),
        consts.ctypes.data_as(u64p), ctypes.c_size_t(len(consts)), rots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_size_t(len(rots)),
        sel_ptrs, ctypes.c_size_t(len(sel)), fx_ptrs, ctypes.c_size_t(len(fx)), ad_ptrs, ctypes.c_size_t(len(ad)),
        ch.ctypes.data_as(u64p), ctypes.c_size_t(ch.shape[0]), ctypes.c_uint32(log_rows), ctypes.c_int(threads), out.ctypes.data_as(u64p))
    if rc != 0:
        raise IndexError(f"so_graph_evaluate error {rc} (EvalError)")
    return out


def tiny_gate_expression():
    """src/ivc/cyclefold/support_circuit/tiny_gate.rs:55-81 through Expression::from_halo2_expr: 1 selector, fixed
    [mul, sum0, sum1, rc], advice [state0, state1, output]."""
    s = Poly(0)
    mul, sum0, sum1, rc = (Poly(1 + j) for j in range(4))
    state0, state1, output = (Poly(5 + j) for j in range(3))
    return Mul(s, Sub(Sum(Sum(Sum(Mul(Mul(state0, state1), mul), Mul(state0, sum0)), Mul(state1, sum1)), rc), output))
