"""Host-side mirror of the reference's expression IR for the part the GPU path consumes.

In the real integration the Rust crate keeps building `Expression`s and compiling them with
`GraphEvaluator::new`; the compiled op-list is what crosses the C ABI (`sb_expr_compile`).  This module
restates that host logic so the benches and tests of this repo can produce the same op-lists:

    Expression, Query, QueryIndexContext      src/polynomial/expression.rs:38-120
    Expression.homogeneous                     src/polynomial/expression.rs:356-429
    challenge_in_degree                        src/polynomial/expression.rs:503-515
    compress_expression                        src/plonk/util.rs:35-55
    GraphEvaluator (compile only)              src/polynomial/graph_evaluator.rs:57-89, 164-351
    MainGate gate polynomial                   src/main_gate.rs:535-583

Field constants are plain Python ints (canonical, not Montgomery); conversion happens when a program is
uploaded.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47

# ---------------------------------------------------------------------------------------------- Expression


@dataclass(frozen=True)
class QueryIndexContext:
    num_selectors: int = 0
    num_fixed: int = 0
    num_advice: int = 0
    num_challenges: int = 0
    num_lookups: int = 0

    def num_fold_vars(self) -> int:
        return self.num_advice + self.num_lookups * 5


class Expression:
    """tagged tuple: ('const', v) | ('poly', index, rotation) | ('chal', index) | ('neg', a) |
    ('sum', a, b) | ('prod', a, b) | ('scaled', a, k)"""

    __slots__ = ("kind", "a", "b")

    def __init__(self, kind, a=None, b=None):
        self.kind, self.a, self.b = kind, a, b

    # constructors
    @staticmethod
    def Constant(v: int) -> "Expression":
        return Expression("const", v)

    @staticmethod
    def Polynomial(index: int, rotation: int = 0) -> "Expression":
        return Expression("poly", index, rotation)

    @staticmethod
    def Challenge(index: int) -> "Expression":
        return Expression("chal", index)

    # operator overloads exactly as expression.rs:455-499
    def __neg__(self):
        return Expression("neg", self)

    def __add__(self, o: "Expression"):
        return Expression("sum", self, o)

    def __sub__(self, o: "Expression"):
        return Expression("sum", self, Expression("neg", o))

    def __mul__(self, o):
        if isinstance(o, Expression):
            return Expression("prod", self, o)
        return Expression("scaled", self, int(o))

    def collect_challenges(self, out: set) -> None:
        k = self.kind
        if k == "chal":
            out.add(self.a)
        elif k in ("neg", "scaled"):
            self.a.collect_challenges(out)
        elif k in ("sum", "prod"):
            self.a.collect_challenges(out)
            self.b.collect_challenges(out)

    def num_challenges(self) -> int:
        s: set = set()
        self.collect_challenges(s)
        return len(s)

    def _is_folded_var(self, ctx: QueryIndexContext) -> bool:
        assert self.kind == "poly"
        return self.a >= ctx.num_selectors + ctx.num_fixed

    def homogeneous(self, ctx: QueryIndexContext) -> Tuple["Expression", int]:
        """expression.rs:356-429 -> (expr, degree); the new challenge has index ctx.num_challenges."""
        new_ch = ctx.num_challenges
        k = self.kind
        if k == "const":
            return self, 0
        if k == "poly":
            return self, (1 if self._is_folded_var(ctx) else 0)
        if k == "chal":
            return self, 1
        if k == "neg":
            e, d = self.a.homogeneous(ctx)
            return Expression("neg", e), d
        if k == "sum":
            (l, ld), (r, rd) = self.a.homogeneous(ctx), self.b.homogeneous(ctx)
            if ld > rd:
                return l + (r * challenge_in_degree(new_ch, ld - rd)), ld
            if ld < rd:
                return (l * challenge_in_degree(new_ch, rd - ld)) + r, rd
            return l + r, ld
        if k == "prod":
            (l, ld), (r, rd) = self.a.homogeneous(ctx), self.b.homogeneous(ctx)
            return l * r, ld + rd
        if k == "scaled":
            e, d = self.a.homogeneous(ctx)
            return Expression("scaled", e, self.b), d
        raise ValueError(k)


def challenge_in_degree(new_challenge_index: int, degree: int) -> Expression:
    ch = Expression.Challenge(new_challenge_index)
    res = ch
    for _ in range(2, degree + 1):
        res = res * ch
    return res


def compress_expression(exprs: List[Expression], challenge_index: int) -> Expression:
    """plonk/util.rs:35-55: P_n + (...(P_1 + 0*y)*y...)*y for n > 1."""
    y = Expression.Challenge(challenge_index)
    if len(exprs) > 1:
        acc = Expression.Constant(0)
        for e in exprs:
            acc = Expression("sum", e, Expression("prod", acc, y))
        return acc
    return exprs[0] if exprs else Expression.Constant(0)


@dataclass
class CompressedGates:
    """plonk/mod.rs:68-121 (without the lazily grouped polynomial, which only the CPU reference needs)."""

    compressed: Expression
    homogeneous: Expression
    degree: int
    ctx: QueryIndexContext

    @staticmethod
    def new(gates: List[Expression], ctx: QueryIndexContext) -> "CompressedGates":
        compressed = compress_expression(gates, ctx.num_challenges)
        ctx = QueryIndexContext(ctx.num_selectors, ctx.num_fixed, ctx.num_advice, compressed.num_challenges(), ctx.num_lookups)
        hom, deg = compressed.homogeneous(ctx)
        ctx = QueryIndexContext(ctx.num_selectors, ctx.num_fixed, ctx.num_advice, hom.num_challenges(), ctx.num_lookups)
        return CompressedGates(compressed, hom, deg, ctx)


# ---------------------------------------------------------------------------------------------- GraphEvaluator

# ValueSource kinds (graph_evaluator.rs:57-68); the declaration order is the PartialOrd order used to
# canonicalise Add/Mul operands (:304-314, :333-337)
VS_CONSTANT, VS_INTERMEDIATE, VS_FIXED, VS_POLY, VS_CHALLENGE = 0, 1, 2, 3, 4
# Calculation opcodes (graph_evaluator.rs:72-89); Horner is never emitted by add_expression
OP_ADD, OP_SUB, OP_MUL, OP_SQUARE, OP_DOUBLE, OP_NEGATE, OP_HORNER, OP_STORE = 0, 1, 2, 3, 4, 5, 6, 7

ValueSource = Tuple[int, int, int]  # (kind, index, rotation-index)


@dataclass
class GraphEvaluator:
    modulus: int
    constants: List[int] = field(default_factory=lambda: [0, 1, 2])
    rotations: List[int] = field(default_factory=list)
    num_intermediates: int = 0
    calculations: List[Tuple[int, ValueSource, Optional[ValueSource], int]] = field(default_factory=list)  # (op, a, b, target)

    @staticmethod
    def new(expr: Expression, modulus: int = FR) -> "GraphEvaluator":
        g = GraphEvaluator(modulus)
        vs = g._add_expression(expr)
        g._add_calculation(OP_STORE, vs, None)
        return g

    def _add_rotation(self, rot: int) -> int:
        if rot in self.rotations:
            return self.rotations.index(rot)
        self.rotations.append(rot)
        return len(self.rotations) - 1

    def _add_constant(self, c: int) -> ValueSource:
        c %= self.modulus
        if c in self.constants:
            return (VS_CONSTANT, self.constants.index(c), 0)
        self.constants.append(c)
        return (VS_CONSTANT, len(self.constants) - 1, 0)

    def _add_calculation(self, op: int, a: ValueSource, b: Optional[ValueSource]) -> ValueSource:
        for (o, aa, bb, target) in self.calculations:
            if o == op and aa == a and bb == b:
                return (VS_INTERMEDIATE, target, 0)
        target = self.num_intermediates
        self.calculations.append((op, a, b, target))
        self.num_intermediates += 1
        return (VS_INTERMEDIATE, target, 0)

    def _add_expression(self, e: Expression) -> ValueSource:
        ZERO, ONE, TWO = (VS_CONSTANT, 0, 0), (VS_CONSTANT, 1, 0), (VS_CONSTANT, 2, 0)
        k = e.kind
        if k == "const":
            return self._add_constant(e.a)
        if k == "poly":
            rot_idx = self._add_rotation(e.b)
            return self._add_calculation(OP_STORE, (VS_POLY, e.a, rot_idx), None)
        if k == "chal":
            return self._add_calculation(OP_STORE, (VS_CHALLENGE, e.a, 0), None)
        if k == "neg":
            if e.a.kind == "const":
                return self._add_constant(-e.a.a)
            ra = self._add_expression(e.a)
            if ra == ZERO:
                return ra
            return self._add_calculation(OP_NEGATE, ra, None)
        if k == "sum":
            if e.b.kind == "neg":
                ra = self._add_expression(e.a)
                rb = self._add_expression(e.b.a)
                if ra == ZERO:
                    return self._add_calculation(OP_NEGATE, rb, None)
                if rb == ZERO:
                    return ra
                return self._add_calculation(OP_SUB, ra, rb)
            ra = self._add_expression(e.a)
            rb = self._add_expression(e.b)
            if ra <= rb:
                return self._add_calculation(OP_ADD, ra, rb)
            return self._add_calculation(OP_ADD, rb, ra)
        if k == "prod":
            ra = self._add_expression(e.a)
            rb = self._add_expression(e.b)
            if ra == ZERO or rb == ZERO:
                return ZERO
            if ra == ONE:
                return rb
            if rb == ONE:
                return ra
            if ra == TWO:
                return self._add_calculation(OP_DOUBLE, rb, None)
            if rb == TWO:
                return self._add_calculation(OP_DOUBLE, ra, None)
            if ra == rb:
                return self._add_calculation(OP_SQUARE, ra, None)
            if ra <= rb:
                return self._add_calculation(OP_MUL, ra, rb)
            return self._add_calculation(OP_MUL, rb, ra)
        if k == "scaled":
            f = e.b % self.modulus
            if f == 0:
                return ZERO
            if f == 1:
                return self._add_expression(e.a)
            cst = self._add_constant(f)
            ra = self._add_expression(e.a)
            return self._add_calculation(OP_MUL, ra, cst)
        raise ValueError(k)


# ---------------------------------------------------------------------------------------------- gates


def main_gate_expression(T: int, fixed_base: int, advice_base: int, num_selectors: int, num_fixed_total: int) -> Expression:
    """The MainGate<T> custom gate as `Expression::from_halo2_expr` sees it (main_gate.rs:535-583,
    expression.rs:305-340).  Fixed columns of this gate start at `fixed_base`, advice at `advice_base`
    (column order: q_1[T], q_5[T], q_m[2], q_i, q_o, rc / state[T], input, out)."""
    fx = lambda j: Expression.Polynomial(num_selectors + fixed_base + j, 0)  # noqa: E731
    ad = lambda j: Expression.Polynomial(num_selectors + num_fixed_total + advice_base + j, 0)  # noqa: E731
    state = [ad(i) for i in range(T)]
    inp, out = ad(T), ad(T + 1)
    q_1 = [fx(i) for i in range(T)]
    q_5 = [fx(T + i) for i in range(T)]
    q_m = [fx(2 * T), fx(2 * T + 1)]
    q_i, q_o, rc = fx(2 * T + 2), fx(2 * T + 3), fx(2 * T + 4)

    def pow_5(v):
        v2 = v * v
        return v2 * v2 * v

    init = q_m[0] * state[0] * state[1] + q_i * inp + rc + q_o * out
    if T >= 4:
        init = q_m[1] * state[2] * state[3] + init
    acc = init
    for s, q1, q5 in zip(state, q_1, q_5):
        acc = acc + (q1 * s + q5 * pow_5(s))
    return acc


def tiny_gate_expression() -> Expression:
    """The Cyclefold support circuit's gate `s * (s0*s1*mul + s0*sum0 + s1*sum1 + rc - output)` as
    `Expression::from_halo2_expr` sees it (src/ivc/cyclefold/support_circuit/tiny_gate.rs:38-84): 1 selector, fixed
    columns [mul, sum0, sum1, rc], advice columns [state0, state1, output]; folding degree 2."""
    s = Expression.Polynomial(0, 0)
    mul, sum0, sum1, rc = (Expression.Polynomial(1 + j, 0) for j in range(4))
    state0, state1, output = (Expression.Polynomial(5 + j, 0) for j in range(3))
    return s * ((state0 * state1 * mul) + (state0 * sum0) + (state1 * sum1) + rc - output)
