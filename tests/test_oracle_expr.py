"""Pins the oracle's expression machinery against the reference's string known-answer tests and checks the
literal cross-term restatement against the defining polynomial identity."""
import numpy as np

from oracle import expr_ref as E
from oracle import pyref as R


def test_main_gate_expr_string_kat():
    # reference src/main_gate.rs:892-907 test_main_gate_expr (T = 2: 9 fixed, 4 advice, no selectors)
    g = E.main_gate_expression(2, 0, 0, 0, 9)
    assert E.visualize(g) == (
        "Z_4 * Z_9 * Z_10 + Z_6 * Z_11 + Z_8 + Z_7 * Z_12 + Z_0 * Z_9 + Z_2 * Z_9 * Z_9 * Z_9 * Z_9 * Z_9 + "
        "Z_1 * Z_10 + Z_3 * Z_10 * Z_10 * Z_10 * Z_10 * Z_10"
    )


def test_main_gate_cross_term_string_kat():
    # reference src/main_gate.rs:909-926 test_main_gate_cross_term
    g = E.main_gate_expression(2, 0, 0, 0, 9)
    cg = E.CompressedGates([g], E.Ctx(num_fixed=9, num_advice=4))
    grouped = cg.grouped()
    assert E.visualize(grouped[0]) == (
        "r_0 * r_0 * r_0 * (Z_10 * Z_9 * Z_4 + r_0 * Z_11 * Z_6 + r_0 * r_0 * Z_8 + r_0 * Z_12 * Z_7) + "
        "r_0 * r_0 * r_0 * r_0 * Z_9 * Z_0 + Z_9 * Z_9 * Z_9 * Z_9 * Z_9 * Z_2 + r_0 * r_0 * r_0 * r_0 * Z_10 * Z_1 + "
        "Z_10 * Z_10 * Z_10 * Z_10 * Z_10 * Z_3"
    )
    assert E.visualize(grouped[5]) == (
        "r_1 * r_1 * r_1 * (Z_14 * Z_13 * Z_4 + r_1 * Z_15 * Z_6 + r_1 * r_1 * Z_8 + r_1 * Z_16 * Z_7) + "
        "r_1 * r_1 * r_1 * r_1 * Z_13 * Z_0 + Z_13 * Z_13 * Z_13 * Z_13 * Z_13 * Z_2 + r_1 * r_1 * r_1 * r_1 * Z_14 * Z_1 + "
        "Z_14 * Z_14 * Z_14 * Z_14 * Z_14 * Z_3"
    )
    assert len(grouped) == 6 and cg.degree == 5


def test_grouped_poly_mul_kat():
    # reference src/polynomial/grouped_poly.rs:379-407 `mul`
    a = [None, None, E.Poly(0), E.Poly(1), E.Poly(2)]
    b = [None, None, E.Poly(3), E.Poly(4), E.Poly(5)]
    got = [f"{d};{E.visualize(x)}" for d, x in enumerate(E.gp_mul(a, b)) if x is not None]
    assert got == [
        "4;Z_3 * Z_0",
        "5;Z_4 * Z_0 + Z_3 * Z_1",
        "6;Z_5 * Z_0 + Z_4 * Z_1 + Z_3 * Z_2",
        "7;Z_5 * Z_1 + Z_4 * Z_2",
        "8;Z_5 * Z_2",
    ]


def test_grouped_poly_add_sub_kat():
    # reference grouped_poly.rs:294-360 simple_add / simple_sub
    big = (1 << 128) - 1
    a = [E.Const(big), E.Poly(0), None, None, None, E.Chal(0)]
    b = [E.Chal(0), None, E.Poly(5, -2), None, None, E.Const(1)]
    assert [f"{d};{E.visualize(x)}" for d, x in enumerate(E.gp_add(a, b)) if x is not None] == [
        "0;0xffffffffffffffffffffffffffffffff + r_0", "1;Z_0", "2;Z_5[-2]", "5;r_0 + 0x1"]
    a2 = [E.Const(big), E.Poly(0), None, None, None, E.Const(1)]
    b2 = [E.Chal(0), None, E.Poly(5, -2), None, None, E.Chal(0)]
    assert [f"{d};{E.visualize(x)}" for d, x in enumerate(E.gp_sub(a2, b2)) if x is not None] == [
        "0;0xffffffffffffffffffffffffffffffff - r_0", "1;Z_0", "2;-Z_5[-2]", "5;0x1 - r_0"]


def test_grouped_poly_creation_kat():
    # reference grouped_poly.rs:409-460 `creation`
    def sum_(xs):
        return E.Sum(xs[0], sum_(xs[1:])) if xs else E.Const(0)

    a, b, c, d, e = [E.Poly(i) for i in range(5)]
    g = E.gp_new(E.Mul(sum_([a, b, c]), sum_([d, e])), E.Ctx(num_advice=5))
    assert [f"{i};{E.visualize(x)}" for i, x in enumerate(g) if x is not None] == [
        "0;(Z_3 + Z_4 + 0x) * (Z_0 + Z_1 + Z_2 + 0x)",
        "1;(Z_8 + Z_9) * (Z_0 + Z_1 + Z_2 + 0x) + (Z_3 + Z_4 + 0x) * (Z_5 + Z_6 + Z_7)",
        "2;(Z_8 + Z_9) * (Z_5 + Z_6 + Z_7)",
    ]


def _toy_structure(k, T_list, seed, modulus=R.FR):
    """MainGate<T> gates side by side (like the Sangria step-folding circuit's column layout)."""
    rng = R.Xoshiro256ss(seed)
    n = 1 << k
    nfix = sum(2 * T + 5 for T in T_list)
    nadv = sum(T + 2 for T in T_list)
    gates, fb, ab = [], 0, 0
    for T in T_list:
        gates.append(E.main_gate_expression(T, fb, ab, 0, nfix))
        fb += 2 * T + 5
        ab += T + 2
    cg = E.CompressedGates(gates, E.Ctx(num_fixed=nfix, num_advice=nadv))
    fixed = [[rng.field(modulus) for _ in range(n)] for _ in range(nfix)]
    S = E.Structure(k, [], fixed, nadv, 0, cg, modulus)
    return S, rng


def test_cross_terms_satisfy_folding_identity():
    """sum_j X^j T_j(row) == P_hom(w1 + X w2, c1 + X c2)(row): the defining property of the cross terms
    (doc comment of VanillaFS, src/nifs/sangria/mod.rs:54-58)."""
    m = R.FR
    S, rng = _toy_structure(3, [2, 2], 99)
    n = 1 << S.k
    W1 = [[rng.field(m) for _ in range(S.num_advice * n)]]
    W2 = [[rng.field(m) for _ in range(S.num_advice * n)]]
    nch = S.gates.ctx.num_challenges - 1  # without u
    c1 = [rng.field(m) for _ in range(nch)]
    c2 = [rng.field(m) for _ in range(nch)]
    u1 = rng.field(m)
    T = E.commit_cross_terms_eval(S, c1, u1, W1, c2, W2)
    assert len(T) == S.gates.degree
    hom = E.GraphEvaluator(S.gates.homogeneous, m)
    nfix = len(S.fixed)
    for X in (0, 1, 5, rng.field(m)):
        ch = [(a + X * b) % m for a, b in zip(c1 + [u1], c2 + [1])]

        def col(row, index):
            if index < nfix:
                return S.fixed[index][row]
            a = index - nfix
            return (W1[0][a * n + row] + X * W2[0][a * n + row]) % m

        for row in range(n):
            lhs = hom.evaluate(col, ch, row, n)
            t0 = E.GraphEvaluator(S.gates.grouped()[0], m).evaluate(
                lambda r, i: S.fixed[i][r] if i < nfix else W1[0][(i - nfix) * n + r], c1 + [u1] + c2 + [1], row, n)
            rhs = (t0 + sum(pow(X, j + 1, m) * T[j][row] for j in range(len(T)))) % m
            assert lhs == rhs


def test_c_interpreter_matches_python(oracle):
    m = R.FR
    S, rng = _toy_structure(4, [2], 7)
    n = 1 << S.k
    W1 = [rng.field(m) for _ in range(S.num_advice * n)]
    W2 = [rng.field(m) for _ in range(S.num_advice * n)]
    ch = [rng.field(m) for _ in range(2 * S.gates.ctx.num_challenges)]
    nfix = len(S.fixed)
    ex = S.gates.grouped()[2]
    ev = E.GraphEvaluator(ex, m)

    def col(row, index):
        if index < nfix:
            return S.fixed[index][row]
        a = index - nfix
        return W1[a * n + row] if a < S.num_advice else W2[(a - S.num_advice) * n + row]

    exp = [ev.evaluate(col, ch, row, n) for row in range(n)]
    fixed = [R.to_mont_limbs(c, m) for c in S.fixed]
    w1 = R.to_mont_limbs(W1, m).reshape(S.num_advice, n, 4)
    w2 = R.to_mont_limbs(W2, m).reshape(S.num_advice, n, 4)
    adv = [w1[i] for i in range(S.num_advice)] + [w2[i] for i in range(S.num_advice)]
    got = E.c_graph_evaluate(R.FIELD_FR, ev, [], fixed, adv, R.to_mont_limbs(ch, m), S.k, threads=2)
    assert R.from_mont_limbs(got, m) == exp


def test_fold_witness_c_vs_python(oracle):
    import ctypes

    m = R.FQ
    rng = R.Xoshiro256ss(3)
    n, d = 64, 3
    w1 = [rng.field(m) for _ in range(n)]
    w2 = [rng.field(m) for _ in range(n)]
    e1 = [rng.field(m) for _ in range(n)]
    T = [[rng.field(m) for _ in range(n)] for _ in range(d)]
    r = rng.field(m)
    W, Ef = E.fold_witness(m, [w1], e1, [w2], T, r)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    lib = oracle.lib()
    a, b = R.to_mont_limbs(w1, m), R.to_mont_limbs(w2, m)
    rl = R.to_mont_limbs([r], m).reshape(4)
    out = np.zeros_like(a)
    lib.so_axpy(1, a.ctypes.data_as(u64p), b.ctypes.data_as(u64p), rl.ctypes.data_as(u64p), out.ctypes.data_as(u64p), ctypes.c_size_t(n))
    assert R.from_mont_limbs(out, m) == W[0]
    Ts = [R.to_mont_limbs(t, m) for t in T]
    ptrs = (u64p * d)(*[t.ctypes.data_as(u64p) for t in Ts])
    e = R.to_mont_limbs(e1, m)
    lib.so_error_fold(1, e.ctypes.data_as(u64p), ptrs, ctypes.c_size_t(d), rl.ctypes.data_as(u64p), out.ctypes.data_as(u64p), ctypes.c_size_t(n))
    assert R.from_mont_limbs(out, m) == Ef
