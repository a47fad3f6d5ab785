"""Straight-line cross-term kernels compiled at run time (NVRTC) against the interpreter kernel they replace: identical
T_1..T_d on the bench's MainGate shapes, the Cyclefold support gate (selector), and a gate with rotations and a
challenge; both fields."""
import numpy as np
import pytest

from oracle import pyref as R

pytestmark = pytest.mark.gpu


def _cross_terms(S, ck, c1, u1, W1, c2, W2, jit):
    from sirius_b200 import _lib
    from sirius_b200 import sangria as SG

    _lib.load().sb_expr_jit_enable(1 if jit else 0)
    try:
        T, commits = SG.VanillaFS.commit_cross_terms(ck, S, c1, u1, [W1], c2, [W2])
    finally:
        _lib.load().sb_expr_jit_enable(1)
    return T, commits


@pytest.mark.parametrize("side_name,k", [("PRIMARY", 11), ("SECONDARY", 9), ("SUPPORT", 10)])
def test_jit_equals_interpreter_on_hot_path_shapes(oracle, side_name, k):
    import sirius_b200
    from sirius_b200 import curves
    from sirius_b200 import polynomial as P
    from sirius_b200 import sangria as SG
    from sirius_b200 import workload as WL

    side = getattr(WL, side_name)
    gates, nfix, nadv = WL.compressed_gates(side)
    nsel = WL.num_selectors(side)
    f, n = side["field"], 1 << k
    cg = P.CompressedGates.new(gates, P.QueryIndexContext(num_selectors=nsel, num_fixed=nfix, num_advice=nadv))
    fixed = [oracle.random_field(f, 10 + j, n) for j in range(nfix)]
    sel = [(np.arange(n) % 3 != 0).astype(np.uint8) for _ in range(nsel)]
    S = SG.PlonkStructure(f, curves.SCALAR_FIELD[side["curve"]], k, sel, fixed, nadv, 0, cg)
    ck = sirius_b200.CommitmentKey(side["curve"], oracle.running_bases(side["curve"], n))
    nch = cg.ctx.num_challenges - 1
    c1, c2, u1 = oracle.random_field(f, 3, nch), oracle.random_field(f, 4, nch), oracle.random_field(f, 5, 1)
    W1, W2 = oracle.random_field(f, 1, nadv * n), oracle.random_field(f, 2, nadv * n)
    Tj, Cj = _cross_terms(S, ck, c1, u1, W1, c2, W2, jit=True)
    Ti, Ci = _cross_terms(S, ck, c1, u1, W1, c2, W2, jit=False)
    assert len(Tj) == cg.degree
    for a, b in zip(Tj, Ti):
        assert np.array_equal(a, b)
    assert np.array_equal(Cj, Ci)
    S.close()
    ck.close()


def test_jit_rotations_selector_challenge(oracle):
    import sirius_b200
    from sirius_b200 import fft
    from sirius_b200 import polynomial as P
    from sirius_b200 import sangria as SG

    E = P.Expression
    k, n = 8, 256
    gate = E.Polynomial(0, 0) * (E.Polynomial(1, -1) * E.Polynomial(2, 1) * E.Polynomial(3, 0) - E.Polynomial(2, -3))
    gate2 = E.Polynomial(1, 2) * E.Polynomial(3, -1) - E.Polynomial(2, 0)
    cg = P.CompressedGates.new([gate, gate2], P.QueryIndexContext(num_selectors=1, num_fixed=1, num_advice=2))
    fixed = [oracle.random_field(0, 77, n)]
    sel = [(np.arange(n) % 2).astype(np.uint8)]
    S = SG.PlonkStructure(0, fft.FR_MODULUS, k, sel, fixed, 2, 0, cg)
    ck = sirius_b200.CommitmentKey(0, oracle.running_bases(0, n))
    nch = cg.ctx.num_challenges - 1
    c1, c2, u1 = oracle.random_field(0, 3, nch), oracle.random_field(0, 4, nch), oracle.random_field(0, 5, 1)
    W1, W2 = oracle.random_field(0, 1, 2 * n), oracle.random_field(0, 2, 2 * n)
    Tj, _ = _cross_terms(S, ck, c1, u1, W1, c2, W2, jit=True)
    Ti, _ = _cross_terms(S, ck, c1, u1, W1, c2, W2, jit=False)
    for a, b in zip(Tj, Ti):
        assert np.array_equal(a, b)
    S.close()
    ck.close()
