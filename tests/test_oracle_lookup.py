"""Pins oracle/lookup_ref.py (CPU, no GPU):
  * the reference's own concatenate_with_padding unit tests (src/util/mod.rs:235-290), restated;
  * evaluate_m's first-occurrence rule (src/plonk/lookup.rs:283-297);
  * the structural invariant of the reference's end-to-end tests: a trace produced by the SPS protocol satisfies
    the compressed gate + lookup relation on every row (PlonkStructure::is_sat, src/plonk/mod.rs:304-361) and
    the log-derivative sum check (:363-397); tampering breaks it;
  * sparse::matrix_multiply / is_sat_permutation on a copy cycle;
and the host-only logic of the product mirror (sirius_b200/lookup.py) that needs no device.
"""
import numpy as np
import pytest

import lookup_circuit as LC
from oracle import expr_ref as E
from oracle import lookup_ref as L
from oracle import pyref as R

M = R.FR


# ---- util::concatenate_with_padding (src/util/mod.rs:235-290)


def test_concatenate_empty():
    assert L.concatenate_with_padding([], 4) == []


def test_single_vector_with_padding():
    assert L.concatenate_with_padding([[1, 2]], 4) == [1, 2, 0, 0]


def test_single_vector_no_padding():
    assert L.concatenate_with_padding([[1, 2, 3, 4]], 4) == [1, 2, 3, 4]


def test_multiple_vectors_with_padding():
    assert L.concatenate_with_padding([[1, 2], [3], [4, 5, 6]], 4) == [1, 2, 0, 0, 3, 0, 0, 0, 4, 5, 6, 0]


def test_pad_size_one():
    assert L.concatenate_with_padding([[1], [2, 3]], 1) == [1, 2, 3]


def test_product_concatenate_matches():
    """the product's host-side concatenate_with_padding (pure data movement) against the same cases"""
    from sirius_b200.lookup import concatenate_with_padding

    cases = [([], 4), ([[1, 2]], 4), ([[1, 2, 3, 4]], 4), ([[1, 2], [3], [4, 5, 6]], 4), ([[1], [2, 3]], 1)]
    for vs, pad in cases:
        got = concatenate_with_padding([R.to_mont_limbs(v, M) for v in vs], pad)
        assert R.from_mont_limbs(got, M) == L.concatenate_with_padding(vs, pad)


# ---- evaluate_m / evaluate_h_g


def test_evaluate_m_first_occurrence():
    l = [5, 7, 5, 9, 5, 0, 0]
    t = [0, 5, 7, 5, 8, 0, 7]
    #    0 twice in l -> row 0 ; 5 three times -> row 1 ; 7 once -> row 2 ; repeats and misses -> 0
    assert L.Arguments.evaluate_m(l, t) == [2, 3, 1, 0, 0, 0, 0]
    assert sum(L.Arguments.evaluate_m(l, t)) == sum(1 for v in l if v in set(t))


def test_evaluate_h_g_zero_denominator():
    r = 11
    l = [3, (M - r) % M, 0]
    t = [1, (M - r) % M, 2]
    m = [4, 9, 0]
    h, g = L.Arguments.evaluate_h_g(l, t, r, m, M)
    assert h[1] == 0 and g[1] == 0 and g[2] == 0
    assert h[0] * (l[0] + r) % M == 1 and h[2] * r % M == 1
    assert g[0] * (t[0] + r) % M == 4


def test_batch_invert_assigned():
    cols = [[("zero",), ("trivial", 7), ("rational", 3, 4), ("rational", 5, 0)]]
    out = L.batch_invert_assigned(cols, M)[0]
    assert out[0] == 0 and out[1] == 7 and out[2] * 4 % M == 3 and out[3] == 0


# ---- sparse / permutation


def test_matrix_multiply_copy_cycle():
    # Z has 6 cells; cells 1 -> 3 -> 4 -> 1 form a copy cycle, everything else maps to itself
    perm = {0: 0, 1: 3, 3: 4, 4: 1, 2: 2, 5: 5}
    P = [(r, c, 1) for r, c in perm.items()]
    Z = [10, 7, 11, 7, 7, 12]
    assert L.matrix_multiply(P, Z, M) == Z
    Zbad = [10, 7, 11, 8, 7, 12]
    Y = L.matrix_multiply(P, Zbad, M)
    assert sum(1 for y, z in zip(Y, Zbad) if y != z) == 2
    with pytest.raises(RuntimeError):
        L.matrix_multiply([(0, 9, 1)], Z, M)
    assert L.permutation_mismatch_count(P, Z[:1], Z[1:] + [99], 0, 5, M) == 0


# ---- the SPS protocol with a lookup argument: structural invariant


def _trace(vector: bool, k: int = 4, seed: int = 3):
    A = LC.OracleAlgebra()
    gates, inputs, tables = LC.expressions(A, vector)
    args = L.Arguments(inputs, tables)
    assert args.has_vector_lookup == vector and args.num_lookups() == 1
    all_gates = gates + args.to_expressions(LC.NUM_SELECTORS, LC.NUM_FIXED, LC.NUM_ADVICE)
    assert len(all_gates) == 1 + 2 + 2
    ctx = E.Ctx(LC.NUM_SELECTORS, LC.NUM_FIXED, LC.NUM_ADVICE, 2 if vector else 1, 1)
    cg = E.CompressedGates(all_gates, ctx)
    fixed, advice = LC.columns(k, M, seed)
    rng = R.Xoshiro256ss(0xABCDEF + seed)
    chal = [rng.field(M) for _ in range(3)]
    W, challenges = L.run_sps_protocol(args, k, [], fixed, advice, M, lambda rnd, Wr: chal[rnd])
    return args, cg, fixed, advice, W, challenges


def _is_sat_rows(cg, fixed, W, challenges, k):
    n = 1 << k
    ev = E.GraphEvaluator(cg.compressed, M)

    def eval_column_var(row, index):
        if index < LC.NUM_FIXED:
            return fixed[index][row]
        return E.eval_advice_var(W, LC.NUM_ADVICE, 1, n, row, index - LC.NUM_FIXED)

    return [ev.evaluate(eval_column_var, challenges, row, n) for row in range(n)]


@pytest.mark.parametrize("vector", [False, True])
def test_sps_trace_satisfies_relation(vector):
    k = 4
    n = 1 << k
    args, cg, fixed, advice, W, challenges = _trace(vector, k)
    assert len(W) == (3 if vector else 2) and len(challenges) == len(W)
    if vector:
        assert [len(w) for w in W] == [3 * n, 3 * n, 2 * n]   # constraint_system_metainfo.rs:58-79
    else:
        assert [len(w) for w in W] == [6 * n, 2 * n]
    # gate-combining challenge: compressed uses challenge index num_challenges-1 = the last squeezed one
    assert cg.compressed is not None
    rows = _is_sat_rows(cg, fixed, W, challenges, k)
    assert rows == [0] * n
    assert L.is_sat_log_derivative(W, k, 1, vector, M)
    # multiplicities: every lookup row hits the table exactly once
    m_col = (W[1] if vector else W[0])[(2 if vector else LC.NUM_ADVICE + 2) * n:][:n]
    assert sum(m_col) == n
    assert m_col[0] >= 1 and all(v == 0 for v in m_col[LC.TABLE_ROWS:])   # padding rows repeat row 0 -> 0 (Q2)


@pytest.mark.parametrize("vector", [False, True])
def test_sps_trace_tamper_detected(vector):
    k = 4
    n = 1 << k
    args, cg, fixed, advice, W, challenges = _trace(vector, k)
    last = [list(w) for w in W]
    last[-1][3] = (last[-1][3] + 1) % M        # h_3 += 1
    assert not L.is_sat_log_derivative(last, k, 1, vector, M)
    assert any(v != 0 for v in _is_sat_rows(cg, fixed, last, challenges, k))
    # a looked-up value outside the table: the relation on l still holds, the sum check fails
    fixed2, advice2 = LC.columns(k, M, 3)
    advice2[0][5] = 100
    advice2[2][5] = advice2[0][5] * advice2[1][5] % M
    W2, ch2 = L.run_sps_protocol(args, k, [], fixed2, advice2, M, lambda rnd, Wr: challenges[rnd])
    assert _is_sat_rows(cg, fixed2, W2, ch2, k) == [0] * n
    assert not L.is_sat_log_derivative(W2, k, 1, vector, M)


def test_product_expressions_match_oracle():
    """the product mirror builds the same lookup expressions as the oracle (compared through the compiled
    GraphEvaluator programs, which is what reaches the device)"""
    from sirius_b200 import lookup as PL
    from sirius_b200 import polynomial as P

    for vector in (False, True):
        go, io, to = LC.expressions(LC.OracleAlgebra(), vector)
        gp, ip, tp = LC.expressions(LC.ProductAlgebra(), vector)
        ao = L.Arguments(io, to)
        ap = PL.Arguments.compress_from(ip, tp)
        assert ap.has_vector_lookup == ao.has_vector_lookup == vector
        eo = go + ao.to_expressions(LC.NUM_SELECTORS, LC.NUM_FIXED, LC.NUM_ADVICE) + ao.lookup_polys + ao.table_polys
        ep = gp + ap.to_expressions(LC.NUM_SELECTORS, LC.NUM_FIXED, LC.NUM_ADVICE) + ap.lookup_polys + ap.table_polys
        assert len(eo) == len(ep)
        for a, b in zip(eo, ep):
            evo = E.GraphEvaluator(a, M)
            evp = P.GraphEvaluator.new(b, M)
            assert [tuple(c) for c in evo.calcs] == [tuple(c) for c in evp.calculations]
            assert list(evo.constants) == list(evp.constants) and list(evo.rotations) == list(evp.rotations)
        co = E.CompressedGates(eo[:5], E.Ctx(0, LC.NUM_FIXED, LC.NUM_ADVICE, 2 if vector else 1, 1))
        cp = P.CompressedGates.new(ep[:5], P.QueryIndexContext(0, LC.NUM_FIXED, LC.NUM_ADVICE, 2 if vector else 1, 1))
        assert co.degree == cp.degree
        assert PL.Arguments.compress_from([], []) is None
