"""CPU pinning of the Protogalaxy oracle (oracle/pg_ref.py) by the identities the reference's own tests use
(src/nifs/protogalaxy/poly/mod.rs:639-853: cmp_with_direct_eval_of_F / _of_G, zero_f / zero_g) and by the
Lagrange known-answer vector.  Both leaf-row modes."""
import pytest

from oracle import expr_ref as E
from oracle import pg_ref as PG
from oracle import pyref as R

M = R.FR


def _structure(k, T, seed, satisfied=False):
    rng = R.Xoshiro256ss(seed)
    n = 1 << k
    nfix, nadv = 2 * T + 5, T + 2
    gates = [E.main_gate_expression(T, 0, 0, 0, nfix)]
    if satisfied:  # all-zero fixed columns: the gate vanishes on any witness
        fixed = [[0] * n for _ in range(nfix)]
    else:
        fixed = [[rng.field(M) for _ in range(n)] for _ in range(nfix)]
    S = PG.PGStructure(k, [], fixed, nadv, 0, gates)
    return S, rng, nadv, n


def _pow_weights(count, c):
    t = count.bit_length() - 1
    out = []
    for i in range(count):
        w = 1
        for h in range(t):
            if (i >> h) & 1:
                w = w * c[h] % M
        out.append(w)
    return out


@pytest.mark.parametrize("mode", ["compat", "correct"])
def test_F_equals_direct_sum(mode):
    S, rng, nadv, n = _structure(3, 2, 1)
    W = [[rng.field(M) for _ in range(nadv * n)]]
    ctx = PG.PolyContext(S, 1)
    t = ctx.betas_count()
    betas, delta = [rng.field(M) for _ in range(t)], rng.field(M)
    F = PG.compute_F(ctx, betas, delta, W, [], mode)
    assert len(F) == ctx.fft_points_count_F()
    f = PG.evaluate_witness_fn(S, W, [], mode)
    leaves = [f(i) for i in range(ctx.count)]
    deltas = [delta]
    for _ in range(t - 1):
        deltas.append(deltas[-1] ** 2 % M)
    for X in (0, 1, rng.field(M)):
        c = [(b + X * d) % M for b, d in zip(betas, deltas)]
        assert PG.poly_eval(F, X) == sum(w * v for w, v in zip(_pow_weights(ctx.count, c), leaves)) % M
    # e is the same tree with the betas themselves
    assert PG.evaluate_e(S, W, [], betas, mode) == sum(w * v for w, v in zip(_pow_weights(ctx.count, betas), leaves)) % M


@pytest.mark.parametrize("mode", ["compat", "correct"])
def test_G_equals_direct_sum_and_K_divides(mode):
    S, rng, nadv, n = _structure(2, 2, 2)
    Ws = [[[rng.field(M) for _ in range(nadv * n)]] for _ in range(2)]
    ctx = PG.PolyContext(S, 1)
    t = ctx.betas_count()
    bs = [rng.field(M) for _ in range(t)]
    G = PG.compute_G(ctx, bs, Ws[0], [], [Ws[1]], [[]], mode)
    assert len(G) == ctx.fft_points_count_G
    w = _pow_weights(ctx.count, bs)
    for X in (3, rng.field(M)):
        L = R.eval_lagrange_polys(ctx.lagrange_domain(), X)
        folded = [[sum(L[j] * Ws[j][0][i] for j in range(2)) % M for i in range(nadv * n)]]
        f = PG.evaluate_witness_fn(S, folded, [], mode)
        assert PG.poly_eval(G, X) == sum(wi * f(i) for i, wi in enumerate(w)) % M
    # K interpolates (G - F(alpha) L0) / Z on the coset zeta*H (compute_K_from_G, poly/mod.rs:475-509)
    F_alpha = rng.field(M)
    K = PG.compute_K_from_G(ctx, list(G), F_alpha)
    logK = ctx.fft_log_domain_size_K()
    assert len(K) == 1 << logK
    back = R.coset_fft(K)
    for i, w in enumerate(R.iter_cyclic_subgroup(logK)):
        if i % 37:
            continue
        X = R.FR_ZETA * w % M
        L0 = R.eval_lagrange_polys(ctx.lagrange_domain(), X)[0]
        Z = (pow(X, ctx.instances_to_fold, M) - 1) % M
        assert (F_alpha * L0 + Z * back[i]) % M == PG.poly_eval(G, X)


def test_zero_F_and_G_on_satisfied_structure():
    # reference zero_f / zero_g: a satisfied trace gives the zero polynomials
    S, rng, nadv, n = _structure(2, 2, 3, satisfied=True)
    Ws = [[[rng.field(M) for _ in range(nadv * n)]] for _ in range(2)]
    ctx = PG.PolyContext(S, 1)
    t = ctx.betas_count()
    betas = [rng.field(M) for _ in range(t)]
    assert PG.compute_F(ctx, betas, rng.field(M), Ws[0], []) == [0] * ctx.fft_points_count_F()
    assert PG.compute_G(ctx, betas, Ws[0], [], [Ws[1]], [[]]) == [0] * ctx.fft_points_count_G


def test_poly_context_quirk_F5():
    # fft_log_domain_size_K returns a point COUNT used as a log: L=1, degree-5 gate -> 2^8 points
    S, _, _, _ = _structure(2, 2, 4)
    ctx = PG.PolyContext(S, 1)
    assert ctx.fft_points_count_G == 8 and ctx.fft_log_domain_size_K() == 8
    ctx3 = PG.PolyContext(S, 3)
    assert ctx3.fft_points_count_G == 16 and ctx3.fft_log_domain_size_K() == 16


# ---- oracle/pg_fast.py (C interpreter + array tree, used at k >= 12) pinned to the literal pg_ref.py ----------------
@pytest.mark.parametrize("mode", ["compat", "correct"])
@pytest.mark.parametrize("k,T_list,L", [(3, [2], 1), (4, [2, 2], 1), (3, [2], 3)])
def test_pg_fast_equals_literal_restatement(oracle, mode, k, T_list, L):
    from oracle import pg_fast as PF

    n = 1 << k
    nfix = sum(2 * T + 5 for T in T_list)
    nadv = sum(T + 2 for T in T_list)
    gates, fb, ab = [], 0, 0
    for T in T_list:
        gates.append(E.main_gate_expression(T, fb, ab, 0, nfix))
        fb, ab = fb + 2 * T + 5, ab + T + 2
    gates.append(E.Sub(E.Mul(E.Poly(nfix + 0, 1), E.Chal(0)), E.Poly(nfix + 1, -1)))   # a challenge and rotated cells
    fixed = [oracle.random_field(R.FIELD_FR, 100 * k + i, n) for i in range(nfix)]
    So = PG.PGStructure(k, [], [R.from_mont_limbs(f, M) for f in fixed], nadv, 0, gates, num_challenges=1)
    Sf = PF.Structure(k, [], fixed, nadv, gates)
    ctx = PG.PolyContext(So, L)
    assert Sf.betas_count() == ctx.betas_count() and Sf.count() == ctx.count
    Ws = [oracle.random_field(R.FIELD_FR, 7 * k + j, nadv * n) for j in range(L + 1)]
    Wi = [[R.from_mont_limbs(w, M)] for w in Ws]
    rng = R.Xoshiro256ss(5 + k)
    chs = [[rng.field(M)] for _ in range(L + 1)]
    t = ctx.betas_count()
    betas = [rng.field(M) for _ in range(t)]
    delta, alpha, gamma = rng.field(M), rng.field(M), rng.field(M)
    assert PF.compute_F(Sf, betas, delta, Ws[0], chs[0], mode) == PG.compute_F(ctx, betas, delta, Wi[0], chs[0], mode)
    assert PF.evaluate_e(Sf, Ws[0], chs[0], betas, mode) == PG.evaluate_e(So, Wi[0], chs[0], betas, mode)
    bs = PG.beta_stroke(betas, alpha, delta)
    max_degree = max(PG.gate_degree(g, So.ctx) for g in gates)
    assert PF.compute_G(Sf, max_degree, bs, Ws[0], chs[0], Ws[1:], chs[1:], mode) == PG.compute_G(ctx, bs, Wi[0], chs[0], Wi[1:], chs[1:], mode)
    Lg = R.eval_lagrange_polys(ctx.lagrange_domain(), gamma)
    assert R.from_mont_limbs(PF.fold_witness(Ws[0], Ws[1:], Lg), M) == PG.fold_witness(Wi[0], Wi[1:], Lg)[0]


def test_c_tree_parallel_levels_equal_serial_subtrees(oracle):
    """so_beta_tree runs the nodes of a level on all cores above 4096 nodes; the result must equal the composition of
    2^12-leaf subtrees (computed on the serial path) with a top tree -- guards the ping-pong buffering."""
    from oracle import pg_fast as PF

    t, n = 16, 1 << 16
    lv = oracle.random_field(R.FIELD_FR, 99, n)
    rng = R.Xoshiro256ss(3)
    c = [rng.field(M) for _ in range(t)]
    sub = [PF.tree(lv[i * 4096:(i + 1) * 4096], c[:12]) for i in range(16)]
    assert PF.tree(lv, c) == PF.tree(R.to_mont_limbs(sub, M), c[12:])
    # and the weighted-sum identity on a small tree
    small = oracle.random_field(R.FIELD_FR, 7, 8)
    vals = R.from_mont_limbs(small, M)
    assert PF.tree(small, c[:3]) == sum(w * v for w, v in zip(_pow_weights(8, c[:3]), vals)) % M


@pytest.mark.parametrize("mode", [0, 1])
def test_step_ref_protogalaxy_prove_equals_literal_restatement(oracle, mode):
    """oracle/step_ref.protogalaxy_prove -- the checker of bench.py's cyclefold_poseidon and gate_scaling workloads and of the GPU
    parity tests -- on a gate-scaling structure (MainGate<5> + two MainGate<3>: 3 gates, 2^(k+2) leaves) against pg_ref.py's
    element-by-element compute_F / compute_G / compute_K_from_G / fold_witness, both row modes."""
    from oracle import step_ref
    from sirius_b200 import workload as WL   # shapes only

    k = 3
    side = WL.gate_scaling_side(2)
    gates, nfix, nadv = WL.compressed_gates(side, E)
    n = 1 << k
    fixed = [oracle.random_field(R.FIELD_FR, 400 + i, n) for i in range(nfix)]
    W_acc, W_in = oracle.random_field(R.FIELD_FR, 1, nadv * n), oracle.random_field(R.FIELD_FR, 2, nadv * n)
    So = PG.PGStructure(k, [], [R.from_mont_limbs(f, M) for f in fixed], nadv, 0, gates)
    ctx = PG.PolyContext(So, 1)
    t = ctx.betas_count()
    assert t == k + 2
    rng = R.Xoshiro256ss(11)
    betas = [rng.field(M) for _ in range(t)]
    delta, alpha, gamma = rng.field(M), rng.field(M), rng.field(M)
    got = step_ref.protogalaxy_prove(dict(k=k, row_mode=mode, fixed=fixed, nadv=nadv, W_acc=W_acc, W_in=W_in, betas=betas, delta=delta,
                                          alpha=alpha, gamma=gamma), side)
    name = "correct" if mode == 1 else "compat"
    Wa, Wi = [R.from_mont_limbs(W_acc, M)], [R.from_mont_limbs(W_in, M)]
    F = PG.compute_F(ctx, betas, delta, Wa, [], name)
    bs = PG.beta_stroke(betas, alpha, delta)
    G = PG.compute_G(ctx, bs, Wa, [], [Wi], [[]], name)
    assert got["poly_F"] == F and got["poly_G"] == G
    assert got["poly_K"] == PG.compute_K_from_G(ctx, G, PG.poly_eval(F, alpha))
    Lg = R.eval_lagrange_polys(ctx.lagrange_domain(), gamma)
    assert R.from_mont_limbs(got["W"], M) == PG.fold_witness(Wa, [Wi], Lg)[0]
