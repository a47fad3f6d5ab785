"""The reference's own GraphEvaluator unit tests (src/polynomial/graph_evaluator.rs:447-634: constant, sum_const,
product_const, neg_const, poly, challenge, eval) restated for the oracle -- both its literal Python `evaluate` and the C
interpreter the CPU baseline runs -- and for the product's host-side compiler (sirius_b200/polynomial.py), whose compiled
calculation lists must equal the oracle's for every expression of these tests.  Random field elements as in the Rust
tests (seeded here).  The Mock of the reference has 2 rows: rotations wrap modulo 2 (get_rotation_idx, :51-53)."""
import random

import numpy as np
import pytest

from oracle import expr_ref as E
from oracle import pyref as R
from sirius_b200 import polynomial as P

M = R.FR


def to_product(e):
    """expr_ref tuple expression -> sirius_b200.polynomial.Expression"""
    t = e[0]
    if t == "C":
        return P.Expression.Constant(e[1])
    if t == "P":
        return P.Expression.Polynomial(e[1], e[2])
    if t == "H":
        return P.Expression.Challenge(e[1])
    if t == "N":
        return -to_product(e[1])
    if t == "S":
        return to_product(e[1]) + to_product(e[2])
    if t == "M":
        return to_product(e[1]) * to_product(e[2])
    raise ValueError(t)


class Mock:
    """graph_evaluator.rs:413-445: selectors -> fixed -> advice column index map, 2 rows"""

    def __init__(self, advice=(), fixed=(), selectors=(), challenges=()):
        self.advice, self.fixed, self.selectors, self.challenges = list(advice), list(fixed), list(selectors), list(challenges)

    def eval_column_var(self, row, index):
        ns, nf = len(self.selectors), len(self.fixed)
        if index < ns:
            return 1 if self.selectors[index][row] else 0
        if index < ns + nf:
            return self.fixed[index - ns][row]
        return self.advice[index - ns - nf][row]


def evaluate_everywhere(oracle, expr, data: Mock, row: int) -> int:
    """value at `row` by the Python evaluator, cross-checked against the C interpreter and the product's compiler"""
    ev = E.GraphEvaluator(expr, M)
    got = ev.evaluate(data.eval_column_var, data.challenges, row, 2)
    # the product's host-side compiler emits the same program
    pe = P.GraphEvaluator.new(to_product(expr), M)
    assert pe.constants == ev.constants and pe.rotations == ev.rotations
    assert [(o, a, b, t) for o, a, b, t in pe.calculations] == ev.calcs
    # the C interpreter (all rows at once), when there is at least one column to size the table from
    sel = [np.array(s, dtype=np.uint8) for s in data.selectors]
    fx = [R.to_mont_limbs(c, M) for c in data.fixed]
    ad = [R.to_mont_limbs(c, M) for c in data.advice]
    ch = R.to_mont_limbs(data.challenges if data.challenges else [0], M)
    out = E.c_graph_evaluate(R.FIELD_FR, ev, sel, fx, ad, ch, 1)
    assert R.from_mont_limbs(out, M)[row] == got
    return got


@pytest.fixture
def rnd():
    r = random.Random(0x51)
    return lambda: r.randrange(M)


def test_constant(oracle, rnd):
    v = rnd()
    assert evaluate_everywhere(oracle, E.Const(v), Mock(), 0) == v


def test_sum_const(oracle, rnd):
    a, b = rnd(), rnd()
    assert evaluate_everywhere(oracle, E.Sum(E.Const(a), E.Const(b)), Mock(), 0) == (a + b) % M


def test_product_const(oracle, rnd):
    a, b = rnd(), rnd()
    assert evaluate_everywhere(oracle, E.Mul(E.Const(a), E.Const(b)), Mock(), 0) == a * b % M


def test_neg_const(oracle, rnd):
    v = rnd()
    assert evaluate_everywhere(oracle, E.Neg(E.Const(v)), Mock(), 0) == (-v) % M


def test_challenge(oracle, rnd):
    v = rnd()
    assert evaluate_everywhere(oracle, E.Chal(0), Mock(challenges=[v]), 0) == v


def test_poly(oracle, rnd):
    a00, a01, a10, a11, f00, f01, f10, f11 = [rnd() for _ in range(8)]
    s1, s2 = True, False
    data = Mock(advice=[[a00, a10], [a01, a11]], fixed=[[f00, f10], [f01, f11]], selectors=[[s1, s2], [s1, s2]])
    ns, nf = 2, 2
    ev_sel = lambda c, rot, row: evaluate_everywhere(oracle, E.Poly(c, rot), data, row)
    ev_fix = lambda c, rot, row: evaluate_everywhere(oracle, E.Poly(ns + c, rot), data, row)
    ev_adv = lambda c, rot, row: evaluate_everywhere(oracle, E.Poly(ns + nf + c, rot), data, row)
    assert ev_adv(0, 0, 0) == a00
    assert ev_adv(0, 1, 0) == a10
    assert ev_adv(0, 0, 1) == a10
    assert ev_adv(0, -1, 1) == a00
    assert ev_adv(1, 0, 1) == a11
    assert ev_fix(0, 0, 0) == f00
    assert ev_fix(0, 0, 1) == f10
    assert ev_fix(0, -1, 1) == f00
    assert ev_sel(0, 0, 0) == (1 if s1 else 0)
    assert ev_sel(0, 0, 1) == (1 if s2 else 0)
    assert ev_adv(0, 2, 0) == a00      # rotations wrap modulo the 2 rows
    assert ev_adv(0, 1, 1) == a00


def test_eval(oracle, rnd):
    a00, a01, a10, a11, f00, f01, f10, f11 = [rnd() for _ in range(8)]
    data = Mock(advice=[[a00, a10], [a01, a11]], fixed=[[f00, f10], [f01, f11]], selectors=[[False, False], [False, False]])
    ns, nf = 2, 2

    def total(exprs):  # the reference's right-nested `sum` ending in Constant(0)
        return E.Sum(exprs[0], total(exprs[1:])) if exprs else E.Const(0)

    fixed = lambda c, rot: E.Poly(ns + c, rot)
    advice = lambda c, rot: E.Poly(ns + nf + c, rot)
    expr = E.Mul(total([advice(0, 0), advice(1, 0), advice(1, 0)]), total([fixed(0, 0), advice(0, 0)]))
    assert evaluate_everywhere(oracle, expr, data, 0) == (a00 + a01 + a01) * (f00 + a00) % M
