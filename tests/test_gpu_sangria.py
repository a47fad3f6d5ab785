"""GPU parity: fused cross-term kernel, generic expression evaluation and the witness folds through the C ABI,
against the oracle's literal restatement of commit_cross_terms / GraphEvaluator / RelaxedPlonkWitness::fold."""
import ctypes

import numpy as np
import pytest

from oracle import expr_ref as E
from oracle import pyref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import sirius_b200

    sirius_b200.load()
    return sirius_b200


def _build(sb, oracle, field, k, T_list, seed, selectors=0):
    """Product-side PlonkStructure + oracle-side Structure over identical random columns."""
    from sirius_b200 import polynomial as P
    from sirius_b200 import sangria as SG

    m = R.MODULUS[field]
    n = 1 << k
    nfix = sum(2 * T + 5 for T in T_list)
    nadv = sum(T + 2 for T in T_list)
    gp, go, fb, ab = [], [], 0, 0
    for T in T_list:
        gp.append(P.main_gate_expression(T, fb, ab, selectors, nfix))
        go.append(E.main_gate_expression(T, fb, ab, selectors, nfix))
        fb += 2 * T + 5
        ab += T + 2
    if selectors:  # gate the first polynomial with selector 0 like a halo2 selector would
        gp[0] = P.Expression.Polynomial(0, 0) * gp[0]
        go[0] = E.Mul(E.Poly(0, 0), go[0])
    cp = P.CompressedGates.new(gp, P.QueryIndexContext(num_selectors=selectors, num_fixed=nfix, num_advice=nadv))
    co = E.CompressedGates(go, E.Ctx(num_selectors=selectors, num_fixed=nfix, num_advice=nadv))
    fixed = [oracle.random_field(field, seed * 1000 + i, n) for i in range(nfix)]
    rng = np.random.default_rng(seed)
    sels = [rng.integers(0, 2, size=n).astype(np.uint8) for _ in range(selectors)]
    S = SG.PlonkStructure(field, m, k, sels, fixed, nadv, 0, cp)
    return S, co, fixed, sels, nadv


def _oracle_cross_terms(field, co, fixed, sels, nadv, k, W1, W2, ch_all):
    n = 1 << k
    adv = [W1[i * n:(i + 1) * n] for i in range(nadv)] + [W2[i * n:(i + 1) * n] for i in range(nadv)]
    out = []
    for ex in co.grouped()[1:]:
        if ex is None:
            out.append(np.zeros((n, 4), dtype=np.uint64))
        else:
            out.append(E.c_graph_evaluate(field, E.GraphEvaluator(ex, R.MODULUS[field]), sels, fixed, adv, ch_all, k))
    return out


@pytest.mark.parametrize("field,curve", [(R.FIELD_FR, R.CURVE_BN256), (R.FIELD_FQ, R.CURVE_GRUMPKIN)])
@pytest.mark.parametrize("k,T_list,selectors", [(3, [2], 0), (6, [2, 2], 0), (9, [5, 3], 0), (7, [5], 1), (12, [5, 3], 0)])
def test_commit_cross_terms(sb, oracle, field, curve, k, T_list, selectors):
    from sirius_b200 import sangria as SG

    m = R.MODULUS[field]
    n = 1 << k
    S, co, fixed, sels, nadv = _build(sb, oracle, field, k, T_list, 17 + k, selectors)
    W1 = oracle.random_field(field, 501 + k, nadv * n)
    W2 = oracle.random_field(field, 502 + k, nadv * n)
    nch = co.ctx.num_challenges - 1
    c1 = oracle.random_field(field, 601, nch)
    c2 = oracle.random_field(field, 602, nch)
    u1 = oracle.random_field(field, 603, 1)
    one = R.to_mont_limbs([1], m)
    bases = oracle.running_bases(curve, n)
    ck = sb.CommitmentKey(curve, bases)
    T, commits = SG.VanillaFS.commit_cross_terms(ck, S, c1, u1, [W1], c2, [W2])
    assert len(T) == co.degree == S.degree
    ch_all = np.concatenate([c1, u1, c2, one])
    exp = _oracle_cross_terms(field, co, fixed, sels, nadv, k, W1, W2, ch_all)
    for j, (g, e) in enumerate(zip(T, exp)):
        assert np.array_equal(g, e), f"T_{j+1}"
    for j, e in enumerate(exp):
        assert np.array_equal(commits[j], oracle.msm(curve, e, bases)), f"commit T_{j+1}"
    S.close()
    ck.close()


def test_cross_terms_python_literal_tiny(sb, oracle):
    """the pure-Python literal restatement (no C interpreter) agrees too, on a tiny table"""
    from sirius_b200 import sangria as SG

    field, m, k = R.FIELD_FR, R.FR, 3
    n = 1 << k
    S, co, fixed, sels, nadv = _build(sb, oracle, field, k, [2, 2], 5)
    W1 = oracle.random_field(field, 1, nadv * n)
    W2 = oracle.random_field(field, 2, nadv * n)
    nch = co.ctx.num_challenges - 1
    c1, c2, u1 = oracle.random_field(field, 3, nch), oracle.random_field(field, 4, nch), oracle.random_field(field, 5, 1)
    ck = sb.CommitmentKey(R.CURVE_BN256, oracle.running_bases(R.CURVE_BN256, n))
    T, _ = SG.VanillaFS.commit_cross_terms(ck, S, c1, u1, [W1], c2, [W2])
    St = E.Structure(k, [], [R.from_mont_limbs(f, m) for f in fixed], nadv, 0, co, m)
    exp = E.commit_cross_terms_eval(St, R.from_mont_limbs(c1, m), R.from_mont_limbs(u1, m)[0], [R.from_mont_limbs(W1, m)],
                                    R.from_mont_limbs(c2, m), [R.from_mont_limbs(W2, m)])
    for g, e in zip(T, exp):
        assert R.from_mont_limbs(g, m) == e
    S.close()
    ck.close()


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_expr_eval_generic(sb, oracle, field):
    """sb_expr_eval on a two-instance (grouped) expression == the literal interpreter: exercises the
    PlonkEvalDomain index space (second instance at +num_fold_vars) and rotations."""
    from sirius_b200 import _lib
    from sirius_b200 import polynomial as P
    from sirius_b200 import sangria as SG

    m = R.MODULUS[field]
    k, nadv = 7, 3
    n = 1 << k
    nfix = 2
    # (a0[+1] * b1[-1] + f0 * a2 - c0) * (b0 + a1[+3]) with a = instance 1, b = instance 2 columns
    col = lambda i, r=0: P.Expression.Polynomial(nfix + i, r)  # noqa: E731
    ocol = lambda i, r=0: E.Poly(nfix + i, r)  # noqa: E731
    ep = (col(0, 1) * col(nadv + 1, -1) + P.Expression.Polynomial(0) * col(2) - P.Expression.Challenge(0)) * (col(nadv + 0) + col(1, 3))
    eo = E.Mul(E.Sub(E.Sum(E.Mul(ocol(0, 1), ocol(nadv + 1, -1)), E.Mul(E.Poly(0), ocol(2))), E.Chal(0)), E.Sum(ocol(nadv + 0), ocol(1, 3)))
    fixed = [oracle.random_field(field, 70 + i, n) for i in range(nfix)]
    W1 = oracle.random_field(field, 80, nadv * n)
    W2 = oracle.random_field(field, 81, nadv * n)
    ch = oracle.random_field(field, 82, 2)
    cg = P.CompressedGates.new([ep], P.QueryIndexContext(num_fixed=nfix, num_advice=nadv))
    S = SG.PlonkStructure(field, m, k, [], fixed, nadv, 0, cg)
    prog = SG.Program(field, P.GraphEvaluator.new(ep, m))
    lib = _lib.load()
    a1 = (_lib.u64p * 1)(W1.ctypes.data_as(_lib.u64p))
    a2 = (_lib.u64p * 1)(W2.ctypes.data_as(_lib.u64p))
    l1 = (ctypes.c_size_t * 1)(W1.shape[0])
    out = np.zeros((n, 4), dtype=np.uint64)
    _lib.check(lib.sb_expr_eval(prog._h, S._cols, nadv, 0, a1, l1, 1, a2, l1, 1, ch.ctypes.data_as(_lib.u64p), 2, out.ctypes.data_as(_lib.u64p)))
    adv = [W1[i * n:(i + 1) * n] for i in range(nadv)] + [W2[i * n:(i + 1) * n] for i in range(nadv)]
    exp = E.c_graph_evaluate(field, E.GraphEvaluator(eo, m), [], fixed, adv, ch, k)
    assert np.array_equal(out, exp)
    # a column index past both instances is the reference's ColumnVariableIndexOutOfBoundary
    bad = SG.Program(field, P.GraphEvaluator.new(col(2 * nadv), m))
    with pytest.raises(sb.SiriusB200Error):
        _lib.check(lib.sb_expr_eval(bad._h, S._cols, nadv, 0, a1, l1, 1, a2, l1, 1, ch.ctypes.data_as(_lib.u64p), 2, out.ctypes.data_as(_lib.u64p)))
    S.close()


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_witness_fold(sb, oracle, field):
    """RelaxedPlonkWitness::fold (accumulator.rs:363-404) vs the C restatement, incl. a full-size W (12 * 2^17)."""
    from sirius_b200 import sangria as SG

    u64p = ctypes.POINTER(ctypes.c_uint64)
    lib = oracle.lib()
    for nw, ne, d in [(7 * 64, 64, 5), (12 << 17, 1 << 17, 6)]:
        w1, w2 = oracle.random_field(field, 1, nw), oracle.random_field(field, 2, nw)
        e1 = oracle.random_field(field, 3, ne)
        T = [oracle.random_field(field, 10 + j, ne) for j in range(d)]
        r = oracle.random_field(field, 4, 1).reshape(4)
        got = SG.RelaxedPlonkWitness(field, [w1], e1).fold([w2], T, r)
        exp_w = np.zeros_like(w1)
        lib.so_axpy(field, w1.ctypes.data_as(u64p), w2.ctypes.data_as(u64p), r.ctypes.data_as(u64p), exp_w.ctypes.data_as(u64p), ctypes.c_size_t(nw))
        ptrs = (u64p * d)(*[t.ctypes.data_as(u64p) for t in T])
        exp_e = np.zeros_like(e1)
        lib.so_error_fold(field, e1.ctypes.data_as(u64p), ptrs, ctypes.c_size_t(d), r.ctypes.data_as(u64p), exp_e.ctypes.data_as(u64p), ctypes.c_size_t(ne))
        assert np.array_equal(got.W[0], exp_w)
        assert np.array_equal(got.E, exp_e)


@pytest.mark.parametrize("field,curve", [(R.FIELD_FR, R.CURVE_BN256), (R.FIELD_FQ, R.CURVE_GRUMPKIN)])
def test_prove_then_is_sat(sb, oracle, field, curve):
    """The reference's end-to-end nifs structure (src/nifs/sangria/tests.rs:185-235): fold an accumulator with an
    incoming trace, then the deciders accept the result (E opens to the evaluated relation, commitments re-open);
    a single corrupted cell is reported."""
    from sirius_b200 import sangria as SG

    m = R.MODULUS[field]
    k, T_list = 8, [5, 3]
    n = 1 << k
    S, co, fixed, sels, nadv = _build(sb, oracle, field, k, T_list, 91)
    nch = co.ctx.num_challenges - 1
    W1, W2 = oracle.random_field(field, 1, nadv * n), oracle.random_field(field, 2, nadv * n)
    c1, c2, u1 = oracle.random_field(field, 3, nch), oracle.random_field(field, 4, nch), oracle.random_field(field, 5, 1)
    r = oracle.random_field(field, 6, 1).reshape(4)
    # a satisfied relaxed accumulator: E1 := P_hom(W1; c1, u1), taken from the ORACLE evaluator
    hom = E.GraphEvaluator(co.homogeneous, m)
    adv1 = [W1[i * n:(i + 1) * n] for i in range(nadv)]
    E1 = E.c_graph_evaluate(field, hom, sels, fixed, adv1, np.concatenate([c1, u1]), k)
    bases = oracle.running_bases(curve, nadv * n)
    ck = sb.CommitmentKey(curve, bases)
    SG.VanillaFS.is_sat_accumulation(S, c1, u1, [W1], E1)
    T, commits = SG.VanillaFS.commit_cross_terms(ck, S, c1, u1, [W1], c2, [W2])
    folded = SG.RelaxedPlonkWitness(field, [W1], E1).fold([W2], T, r)
    # folded instance scalars: challenges + r * challenges2, u + r (accumulator.rs:228-235)
    cf = oracle.field_binop("add", field, c1, oracle.field_binop("mul", field, np.tile(r, (nch, 1)), c2))
    uf = oracle.field_binop("add", field, u1, r.reshape(1, 4))
    SG.VanillaFS.is_sat_accumulation(S, cf, uf, folded.W, folded.E)
    # commitments of the folded witness re-open: C(W') == C(W1) + r C(W2) is the verifier's view; here the prover's
    cW, cE = ck.commit(folded.W[0]), ck.commit(folded.E)
    SG.VanillaFS.is_sat_witness_commit(ck, [cW], folded.W, folded.E, cE)
    assert np.array_equal(cW, oracle.msm(curve, folded.W[0], bases))
    bad = folded.E.copy()
    bad[5, 0] ^= np.uint64(1)
    with pytest.raises(SG.EvaluationMismatch) as ei:
        SG.VanillaFS.is_sat_accumulation(S, cf, uf, folded.W, bad)
    assert ei.value.mismatch_count == 1 and ei.value.total_row == n
    with pytest.raises(SG.ECommitmentMismatch):
        SG.VanillaFS.is_sat_witness_commit(ck, [cW], folded.W, bad, cE)
    S.close()
    ck.close()
