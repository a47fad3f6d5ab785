"""GPU parity: Protogalaxy F / G / K / e / fold_witness through the C ABI against the literal oracle, in both
leaf-row modes (reference-compatible `& total_row` and corrected `% total_row`)."""
import numpy as np
import pytest

from oracle import expr_ref as E
from oracle import pg_ref as PG
from oracle import pyref as R

pytestmark = pytest.mark.gpu
M = R.FR


@pytest.fixture(scope="module")
def sb():
    import sirius_b200

    sirius_b200.load()
    return sirius_b200


def _setup(sb, oracle, k, T_list, seed, traces_len):
    from sirius_b200 import polynomial as P
    from sirius_b200 import sangria as SG

    n = 1 << k
    nfix = sum(2 * T + 5 for T in T_list)
    nadv = sum(T + 2 for T in T_list)
    gp, go, fb, ab = [], [], 0, 0
    for T in T_list:
        gp.append(P.main_gate_expression(T, fb, ab, 0, nfix))
        go.append(E.main_gate_expression(T, fb, ab, 0, nfix))
        fb, ab = fb + 2 * T + 5, ab + T + 2
    # a gate that reads a challenge and a rotated cell, so challenge folding and rotations are exercised
    gp.append(P.Expression.Polynomial(nfix + 0, 1) * P.Expression.Challenge(0) - P.Expression.Polynomial(nfix + 1, -1))
    go.append(E.Sub(E.Mul(E.Poly(nfix + 0, 1), E.Chal(0)), E.Poly(nfix + 1, -1)))
    cg = P.CompressedGates.new(gp, P.QueryIndexContext(num_fixed=nfix, num_advice=nadv))
    fixed = [oracle.random_field(R.FIELD_FR, seed * 100 + i, n) for i in range(nfix)]
    S = SG.PlonkStructure(R.FIELD_FR, M, k, [], fixed, nadv, 0, cg, gates=gp)
    So = PG.PGStructure(k, [], [R.from_mont_limbs(f, M) for f in fixed], nadv, 0, go, num_challenges=1)
    Ws = [oracle.random_field(R.FIELD_FR, seed * 7 + j, nadv * n) for j in range(traces_len + 1)]
    rng = R.Xoshiro256ss(seed)
    chs = [[rng.field(M)] for _ in range(traces_len + 1)]
    return S, So, Ws, chs, rng


@pytest.mark.parametrize("row_mode", ["compat", "correct"])
@pytest.mark.parametrize("k,T_list,L", [(3, [2], 1), (5, [2, 2], 1), (4, [2], 3)])
def test_protogalaxy_polynomials(sb, oracle, k, T_list, L, row_mode):
    from sirius_b200 import protogalaxy as PGX

    S, So, Ws, chs, rng = _setup(sb, oracle, k, T_list, 11 + k, L)
    mode = PGX.ROW_COMPAT if row_mode == "compat" else PGX.ROW_CORRECT
    ctx, ctxo = PGX.PolyContext(S, L), PG.PolyContext(So, L)
    assert ctx.betas_count() == ctxo.betas_count() and ctx.fft_points_count_F() == ctxo.fft_points_count_F()
    assert ctx.fft_points_count_G == ctxo.fft_points_count_G and ctx.fft_log_domain_size_K() == ctxo.fft_log_domain_size_K()
    t = ctx.betas_count()
    betas = [rng.field(M) for _ in range(t)]
    delta, alpha, gamma = rng.field(M), rng.field(M), rng.field(M)
    Wi = [[R.from_mont_limbs(w, M)] for w in Ws]

    # F
    F_gpu = PGX.compute_F(ctx, betas, delta, Ws[0], chs[0], mode)
    F_ref = PG.compute_F(ctxo, betas, delta, Wi[0], chs[0], row_mode)
    assert F_gpu == F_ref and len(F_gpu) == ctx.fft_points_count_F()
    # e
    assert PGX.evaluate_e_from_trace(S, Ws[0], chs[0], betas, mode) == PG.evaluate_e(So, Wi[0], chs[0], betas, row_mode)
    # G and K
    bs = PGX.beta_stroke(betas, alpha, delta)
    assert bs == PG.beta_stroke(betas, alpha, delta)
    G_gpu = PGX.compute_G(ctx, bs, Ws[0], chs[0], Ws[1:], chs[1:], mode)
    G_ref = PG.compute_G(ctxo, bs, Wi[0], chs[0], Wi[1:], chs[1:], row_mode)
    assert G_gpu == G_ref
    if ctx.fft_log_domain_size_K() <= 12:
        F_alpha = PGX.poly_eval(F_gpu, alpha)
        assert PGX.compute_K_from_G(ctx, G_gpu, F_alpha) == PG.compute_K_from_G(ctxo, G_ref, F_alpha)
    # fold_witness at gamma
    Lg = PGX.eval_lagrange_polys(gamma, ctx.lagrange_domain())
    folded = PGX.fold_witness(Ws[0], Ws[1:], Lg)
    assert R.from_mont_limbs(folded, M) == PG.fold_witness(Wi[0], Wi[1:], Lg)[0]
    S.close()


def test_beta_tree_large_vs_direct_sum(sb, oracle):
    """2^17 leaves x 4 points: the tree equals sum_i leaf_i * prod_{bits of i} c_h (the identity the reference's
    cmp_with_direct_eval tests use, poly/mod.rs:639-728), checked against the C oracle's field ops on a sample
    and against a numpy-free python evaluation of a strided subset via linearity."""
    import ctypes

    import torch

    from sirius_b200 import _lib

    lib = _lib.load()
    t, P = 17, 3
    n = 1 << t
    leaves = oracle.random_field(R.FIELD_FR, 4242, n)
    rng = R.Xoshiro256ss(9)
    c = [[rng.field(M) for _ in range(t)] for _ in range(P)]
    d_leaves = torch.from_numpy(leaves.view(np.int64)).cuda()
    d_out = torch.zeros((P, 4), dtype=torch.int64, device="cuda")
    mul = R.to_mont_limbs([x for row in c for x in row], M)
    _lib.check(lib.sb_beta_tree_device(R.FIELD_FR, ctypes.c_void_p(d_leaves.data_ptr()), t, P, 0, mul.ctypes.data_as(_lib.u64p), ctypes.c_void_p(d_out.data_ptr()), None))
    torch.cuda.synchronize()
    got = R.from_mont_limbs(d_out.cpu().numpy().view(np.uint64), M)
    lv = R.from_mont_limbs(leaves, M)
    for p in range(P):
        # direct sum with incremental weights: w(i) = prod over set bits
        w = [1] * n
        for h in range(t):
            step = 1 << h
            ch = c[p][h]
            for base in range(step, n, 2 * step):
                for i in range(base, base + step):
                    w[i] = w[i] * ch % M
        assert got[p] == sum(a * b for a, b in zip(lv, w)) % M
