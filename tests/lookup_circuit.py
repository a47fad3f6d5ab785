"""A tiny circuit with one lookup argument, shared by the CPU (oracle) and GPU (product) lookup tests.

Index space (src/polynomial/expression.rs:58-70): no selectors; fixed q=0, tab1=1, tab2=2; advice a=3, b=4, c=5;
the lookup variables (l,t,m,h,g) follow at 6..10.
  gate   : q * (a*b - c)
  lookup : a in tab1                    (plain lookup  -> 2 witness rounds, plonk/mod.rs:501-575)
           (a, b) in (tab1, tab2)       (vector lookup -> 3 witness rounds, plonk/mod.rs:577-660)
"""
import numpy as np

NUM_SELECTORS, NUM_FIXED, NUM_ADVICE = 0, 3, 3
TABLE_ROWS = 8


class OracleAlgebra:
    def __init__(self):
        from oracle import expr_ref as E

        self.E = E

    def poly(self, i):
        return self.E.Poly(i, 0)

    def mul(self, a, b):
        return self.E.Mul(a, b)

    def sub(self, a, b):
        return self.E.Sub(a, b)


class ProductAlgebra:
    def __init__(self):
        from sirius_b200 import polynomial as P

        self.P = P

    def poly(self, i):
        return self.P.Expression.Polynomial(i, 0)

    def mul(self, a, b):
        return a * b

    def sub(self, a, b):
        return a - b


def expressions(A, vector: bool):
    q, tab1, tab2, a, b, c = [A.poly(i) for i in range(6)]
    gate = A.mul(q, A.sub(A.mul(a, b), c))
    inputs = [[a, b]] if vector else [[a]]
    tables = [[tab1, tab2]] if vector else [[tab1]]
    return [gate], inputs, tables


def columns(k: int, modulus: int, seed: int):
    """-> (fixed columns [q, tab1, tab2], advice columns [a, b, c]) as lists of Python ints; the table is padded with
    copies of its first row (as halo2 pads lookup tables), so repeated table values occur."""
    n = 1 << k
    rng = np.random.default_rng(seed)
    tab1 = [i if i < TABLE_ROWS else 0 for i in range(n)]
    tab2 = [(i * i + 1) if i < TABLE_ROWS else 1 for i in range(n)]
    q = [int(v) for v in rng.integers(0, 2, size=n)]
    a = [int(v) for v in rng.integers(0, TABLE_ROWS, size=n)]
    a[0] = 0            # row 0 of the table is the repeated one: make sure it is looked up
    b = [v * v + 1 for v in a]
    c = [(x * y) % modulus for x, y in zip(a, b)]
    return [q, tab1, tab2], [a, b, c]
