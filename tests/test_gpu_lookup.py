"""GPU parity for the SPS lookup columns, the permutation decider and the witness assembly
through the C ABI, against oracle/lookup_ref.py (bit-exact: integer work)."""
import ctypes

import numpy as np
import pytest

import lookup_circuit as LC
from oracle import expr_ref as E
from oracle import lookup_ref as L
from oracle import pyref as R

pytestmark = pytest.mark.gpu

FIELDS = [R.FIELD_FR, R.FIELD_FQ]


@pytest.fixture(scope="module")
def sb():
    import sirius_b200

    sirius_b200.load()
    return sirius_b200


def _mont(vals, m):
    return R.to_mont_limbs([v % m for v in vals], m)


def _p(a):
    from sirius_b200 import _lib

    return a.ctypes.data_as(_lib.u64p)


# ------------------------------------------------------------------------------------------ evaluate_m


def _multiplicity(field, l, t):
    from sirius_b200 import _lib

    lib = _lib.load()
    m = R.MODULUS[field]
    la, ta = _mont(l, m), _mont(t, m)
    out = np.full((len(t), 4), 0xDEADBEEF, dtype=np.uint64)
    _lib.check(lib.sb_lookup_multiplicity(field, _p(la) if len(l) else None, len(l), _p(ta) if len(t) else None, len(t), _p(out) if len(t) else None))
    return R.from_mont_limbs(out, m) if len(t) else []


@pytest.mark.parametrize("field", FIELDS)
def test_multiplicity_small_cases(sb, field):
    l = [5, 7, 5, 9, 5, 0, 0]
    t = [0, 5, 7, 5, 8, 0, 7]
    assert _multiplicity(field, l, t) == [2, 3, 1, 0, 0, 0, 0] == L.Arguments.evaluate_m(l, t)
    assert _multiplicity(field, [], [1, 2, 1]) == [0, 0, 0]
    assert _multiplicity(field, [4, 4], []) == []
    assert _multiplicity(field, [3] * 1000, [3] * 77) == [1000] + [0] * 76     # one hot value, all rows repeat it
    assert _multiplicity(field, [1, 2, 3], [4, 5, 6]) == [0, 0, 0]


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("n_l,n_t,distinct", [(1 << 10, 1 << 10, 300), (1 << 14, 1 << 12, 5000), (3001, 4097, 64), (1 << 16, 1 << 16, 1 << 15)])
def test_multiplicity_random(sb, oracle, field, n_l, n_t, distinct):
    m = R.MODULUS[field]
    pool = R.from_mont_limbs(oracle.random_field(field, 77 + distinct, distinct), m) + [0, 1, 2]
    rng = np.random.default_rng(n_l + n_t)
    # skewed draws: a few values are very hot (unused rows all look up the same cell), many are absent from t
    t = [pool[i] for i in rng.integers(0, len(pool) // 2 + 1, size=n_t)]
    l = [pool[i] for i in (rng.zipf(1.3, size=n_l) - 1) % len(pool)]
    got = _multiplicity(field, l, t)
    assert got == L.Arguments.evaluate_m(l, t)


def test_multiplicity_large_property(sb):
    """2^20 rows, device resident: sum_i m_i = #{j : l_j in t}; m is zero off first occurrences."""
    import torch

    from sirius_b200 import _lib
    from sirius_b200.device import random_field_device

    lib = _lib.load()
    n = 1 << 20
    base = random_field_device(1 << 12, 5)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(9)
    t = base[torch.randint(0, 1 << 11, (n,), device="cuda", generator=gen)].contiguous()     # values 0..2047 of the pool
    l = base[torch.randint(0, 1 << 12, (n,), device="cuda", generator=gen)].contiguous()     # half of them are in t
    mm = torch.zeros_like(t)
    _lib.check(lib.sb_lookup_multiplicity_device(R.FIELD_FR, l.data_ptr(), n, t.data_ptr(), n, mm.data_ptr(), None))
    torch.cuda.synchronize()
    mh = mm.cpu().numpy().view(np.uint64)
    counts = R.from_mont_limbs(mh[np.any(mh != 0, axis=1)], R.FR)
    # membership of l in t, computed on the host from the pool indices via byte keys
    tkeys = set(map(bytes, t.cpu().numpy().view(np.uint8).reshape(n, 32)))
    hits = sum(1 for row in l.cpu().numpy().view(np.uint8).reshape(n, 32) if bytes(row) in tkeys)
    assert sum(counts) == hits
    assert len(counts) <= len(tkeys)


# ------------------------------------------------------------------------------------------ evaluate_h_g


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000, 1 << 14])
def test_lookup_inverses(sb, oracle, field, n):
    from sirius_b200 import _lib

    lib = _lib.load()
    m = R.MODULUS[field]
    r = R.from_mont_limbs(oracle.random_field(field, 3, 1), m)[0]
    l = R.from_mont_limbs(oracle.random_field(field, 10 + n, n), m)
    t = R.from_mont_limbs(oracle.random_field(field, 20 + n, n), m)
    mm = [int(v) for v in np.random.default_rng(n).integers(0, 5, size=n)]
    for i in range(0, n, 7):          # zero denominators and zero multiplicities
        l[i] = (m - r) % m
    for i in range(3, n, 11):
        t[i] = (m - r) % m
    eh, eg = L.Arguments.evaluate_h_g(l, t, r, mm, m)
    h = np.zeros((n, 4), dtype=np.uint64)
    g = np.zeros((n, 4), dtype=np.uint64)
    _lib.check(lib.sb_lookup_inverses(field, _p(_mont(l, m)), _p(_mont(t, m)), _p(_mont(mm, m)), _p(_mont([r], m)), n, _p(h), _p(g)))
    assert R.from_mont_limbs(h, m) == eh
    assert R.from_mont_limbs(g, m) == eg


@pytest.mark.parametrize("field", FIELDS)
def test_batch_invert_assigned(sb, oracle, field):
    from sirius_b200.lookup import batch_invert_assigned

    m = R.MODULUS[field]
    n = 777
    num = R.from_mont_limbs(oracle.random_field(field, 1, n), m)
    den = R.from_mont_limbs(oracle.random_field(field, 2, n), m)
    col = []
    for i in range(n):
        if i % 5 == 0:
            col.append(("zero",)); num[i], den[i] = 0, 1
        elif i % 5 == 1:
            col.append(("trivial", num[i])); den[i] = 1
        elif i % 13 == 2:
            col.append(("rational", num[i], 0)); den[i] = 0
        else:
            col.append(("rational", num[i], den[i]))
    got = batch_invert_assigned(field, _mont(num, m), _mont(den, m))
    assert R.from_mont_limbs(got, m) == L.batch_invert_assigned([col], m)[0]


# ------------------------------------------------------------------------------------------ sums, sparse, assembly


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 100000])
def test_sum_diff(sb, oracle, field, n):
    from sirius_b200 import _lib

    lib = _lib.load()
    m = R.MODULUS[field]
    a = oracle.random_field(field, 31 + n, n)
    b = oracle.random_field(field, 32 + n, n)
    out = np.zeros(4, dtype=np.uint64)
    _lib.check(lib.sb_sum_diff(field, _p(a) if n else None, _p(b) if n else None, n, _p(out)))
    ai, bi = R.from_mont_limbs(a, m), R.from_mont_limbs(b, m)
    assert R.from_mont_limbs(out.reshape(1, 4), m)[0] == (sum(ai) - sum(bi)) % m
    _lib.check(lib.sb_sum_diff(field, _p(a) if n else None, None, n, _p(out)))
    assert R.from_mont_limbs(out.reshape(1, 4), m)[0] == sum(ai) % m


@pytest.mark.parametrize("field", FIELDS)
def test_sparse_permutation(sb, oracle, field):
    from sirius_b200 import _lib
    from sirius_b200 import lookup as PL

    m = R.MODULUS[field]
    k, num_advice, num_io = 8, 3, 4
    n = 1 << k
    N = num_io + n * num_advice
    rng = np.random.default_rng(5)
    # random copy cycles over a third of the cells, identity elsewhere (construct_permutation_matrix shape:
    # exactly one unit entry per row, src/plonk/permutation.rs:311-384)
    perm = list(range(N))
    cells = rng.permutation(N)[: N // 3]
    Zi = R.from_mont_limbs(oracle.random_field(field, 8, N), m)
    pos = 0
    while pos + 2 <= len(cells):
        ln = int(rng.integers(2, 6))
        cyc = [int(c) for c in cells[pos : pos + ln]]
        pos += ln
        for a, b in zip(cyc, cyc[1:] + cyc[:1]):
            perm[a] = b
        for c in cyc:
            Zi[c] = Zi[cyc[0]]
    one = _mont([1], m)[0]
    P_entries = [(r, c, 1) for r, c in enumerate(perm)]
    Pm = PL.SparseMatrix(field, [(r, c, one) for r, c in enumerate(perm)], N)
    assert _lib.load().sb_sparse_dim(Pm._h) == N
    Z = _mont(Zi, m)
    assert Pm.mismatch_count(Z) == 0 == L.permutation_mismatch_count(P_entries, Zi[:num_io], Zi[num_io:], k, num_advice, m)
    PL.is_sat_permutation(Pm, Z[:num_io], Z[num_io:], k, num_advice)
    bad = list(Zi)
    for c in cells[:40]:
        bad[int(c)] = (bad[int(c)] + 1) % m
    exp = L.permutation_mismatch_count(P_entries, bad[:num_io], bad[num_io:], k, num_advice, m)
    assert exp > 0 and Pm.mismatch_count(_mont(bad, m)) == exp
    with pytest.raises(PL.PermCheckFail) as ei:
        PL.is_sat_permutation(Pm, _mont(bad[:num_io], m), _mont(bad[num_io:], m), k, num_advice)
    assert ei.value.mismatch_count == exp
    Pm.close()
    # a general sparse matrix (several weighted entries per row, empty rows) through the same kernel
    ent, N2 = [], 50
    vals = R.from_mont_limbs(oracle.random_field(field, 9, 200), m)
    for e in range(200):
        ent.append((int(rng.integers(0, N2 - 5)), int(rng.integers(0, N2)), vals[e]))
    Z2 = R.from_mont_limbs(oracle.random_field(field, 10, N2), m)
    Y2 = L.matrix_multiply(ent, Z2, m)
    G = PL.SparseMatrix(field, [(r, c, _mont([v], m)[0]) for r, c, v in ent], N2)
    assert G.mismatch_count(_mont(Z2, m)) == sum(1 for y, z in zip(Y2, Z2) if y != z)
    assert G.mismatch_count(_mont(Y2[:N2 - 5] + [0] * 5, m)) >= 0   # shape check only
    G.close()
    with pytest.raises(sb.SiriusB200Error):   # "invalid matrix multiply" (sparse.rs:15-17)
        PL.SparseMatrix(field, [(0, 7, one)], 5)


def test_sparse_mismatch_device_head_tail(sb, oracle):
    """Z split as host instances ++ device witness, as is_sat_permutation assembles it"""
    import torch

    from sirius_b200 import _lib
    from sirius_b200 import lookup as PL

    lib = _lib.load()
    field, m = R.FIELD_FR, R.FR
    head_len, tail_len = 3, 500
    N = head_len + tail_len
    Zi = R.from_mont_limbs(oracle.random_field(field, 4, N), m)
    perm = list(range(N))
    perm[1], perm[100], perm[400] = 100, 400, 1        # instance cell 1 tied to two witness cells
    Zi[100] = Zi[400] = Zi[1]
    one = _mont([1], m)[0]
    Pm = PL.SparseMatrix(field, [(r, c, one) for r, c in enumerate(perm)], N)
    Z = _mont(Zi, m)
    tail = torch.from_numpy(Z[head_len:].view(np.int64)).cuda()
    head = np.ascontiguousarray(Z[:head_len])
    got = ctypes.c_uint64(99)
    _lib.check(lib.sb_sparse_mismatch_device(Pm._h, _p(head), head_len, tail.data_ptr(), tail_len, ctypes.cast(ctypes.byref(got), _lib.u64p), None))
    assert got.value == 0
    head[1, 0] ^= 1
    _lib.check(lib.sb_sparse_mismatch_device(Pm._h, _p(head), head_len, tail.data_ptr(), tail_len, ctypes.cast(ctypes.byref(got), _lib.u64p), None))
    assert got.value == 2        # row 1 reads cell 100 (unchanged) against the new head; row 400 reads the new head
    with pytest.raises(sb.SiriusB200Error):
        _lib.check(lib.sb_sparse_mismatch_device(Pm._h, _p(head), head_len, tail.data_ptr(), tail_len - 1, ctypes.cast(ctypes.byref(got), _lib.u64p), None))
    Pm.close()


def test_concat_pad_device(sb):
    import torch

    from sirius_b200.lookup import concatenate_with_padding_device

    cases = [([], 4), ([[1, 2]], 4), ([[1, 2, 3, 4]], 4), ([[1, 2], [3], [4, 5, 6]], 4), ([[1], [2, 3]], 1), ([[], [9]], 2)]
    for vs, pad in cases:
        exp = L.concatenate_with_padding(vs, pad)
        d = torch.full((len(exp) + 2, 4), -1, dtype=torch.int64, device="cuda")
        got_len = concatenate_with_padding_device([_mont(v, R.FR) for v in vs], pad, d.data_ptr(), len(exp) + 2)
        torch.cuda.synchronize()
        assert got_len == len(exp)
        h = d.cpu().numpy().view(np.uint64)
        assert R.from_mont_limbs(h[: len(exp)], R.FR) == exp
        assert (h[len(exp):] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()     # nothing written past the end
    d = torch.zeros((3, 4), dtype=torch.int64, device="cuda")
    with pytest.raises(sb.SiriusB200Error):
        concatenate_with_padding_device([_mont([1, 2], R.FR), _mont([3], R.FR)], 2, d.data_ptr(), 3)


# ------------------------------------------------------------------------------------------ the SPS protocol end to end


def _structures(sb, vector, k, seed):
    from sirius_b200 import lookup as PL
    from sirius_b200 import polynomial as P
    from sirius_b200 import sangria as SG

    m = R.FR
    go, io, to = LC.expressions(LC.OracleAlgebra(), vector)
    gp, ip, tp = LC.expressions(LC.ProductAlgebra(), vector)
    ao = L.Arguments(io, to)
    ap = PL.Arguments.compress_from(ip, tp)
    nch = 2 if vector else 1
    co = E.CompressedGates(go + ao.to_expressions(0, LC.NUM_FIXED, LC.NUM_ADVICE), E.Ctx(0, LC.NUM_FIXED, LC.NUM_ADVICE, nch, 1))
    cp = P.CompressedGates.new(gp + ap.to_expressions(0, LC.NUM_FIXED, LC.NUM_ADVICE), P.QueryIndexContext(0, LC.NUM_FIXED, LC.NUM_ADVICE, nch, 1))
    fixed, advice = LC.columns(k, m, seed)
    S = SG.PlonkStructure(R.FIELD_FR, m, k, [], [_mont(f, m) for f in fixed], LC.NUM_ADVICE, 1, cp, lookup_arguments=ap)
    return S, ao, co, fixed, advice


@pytest.mark.parametrize("vector", [False, True])
@pytest.mark.parametrize("k", [4, 9])
def test_sps_protocol_with_lookup(sb, oracle, vector, k):
    """run_sps_protocol_{2,3}: every witness round bit-exact with the oracle, the commitments open, the trace
    satisfies is_sat (gate + lookup relation on every row, log-derivative sums, commitments)."""
    from sirius_b200 import lookup as PL
    from sirius_b200 import sangria as SG

    m = R.FR
    n = 1 << k
    S, ao, co, fixed, advice = _structures(sb, vector, k, 3)
    rounds = 3 if vector else 2
    chal = R.from_mont_limbs(oracle.random_field(R.FIELD_FR, 99, rounds), m)
    bases = oracle.running_bases(R.CURVE_BN256, 6 * n)
    ck = sb.CommitmentKey(R.CURVE_BN256, bases)
    W, C, ch = PL.run_sps_protocol(S, [_mont(a, m) for a in advice], ck, lambda rnd, Ci: _mont([chal[rnd]], m)[0])
    We, che = L.run_sps_protocol(ao, k, [], fixed, advice, m, lambda rnd, Wr: chal[rnd])
    assert len(W) == len(We) == rounds
    for i, (g, e) in enumerate(zip(W, We)):
        assert R.from_mont_limbs(g, m) == e, f"witness round {i}"
    for g, Ci in zip(W, C):
        assert np.array_equal(Ci, oracle.msm(R.CURVE_BN256, g, bases[: g.shape[0]]))
    S.is_sat(ck, np.stack(ch), C, W)
    assert PL.is_sat_log_derivative(S, W)
    # tampered h column: relation and sum check both fail
    Wb = [w.copy() for w in W]
    Wb[-1][3] = _mont([(R.from_mont_limbs(Wb[-1][3:4], m)[0] + 1) % m], m)[0]
    assert not PL.is_sat_log_derivative(S, Wb)
    with pytest.raises(SG.EvaluationMismatch):
        S.is_sat(ck, np.stack(ch), C, Wb)
    # a value outside the table: rows still vanish, the log-derivative check rejects
    adv2 = [list(a) for a in advice]
    adv2[0][5] = 100
    adv2[2][5] = adv2[0][5] * adv2[1][5] % m
    W2, C2, ch2 = PL.run_sps_protocol(S, [_mont(a, m) for a in adv2], ck, lambda rnd, Ci: _mont([chal[rnd]], m)[0])
    with pytest.raises(SG.LogDerivativeNotSat):
        S.is_sat(ck, np.stack(ch2), C2, W2)
    S.close()
    ck.close()


@pytest.mark.parametrize("vector", [False, True])
def test_sangria_fold_with_lookup(sb, oracle, vector):
    """commit_cross_terms over the 2- and 3-round layouts (PlonkEvalDomain index map, src/plonk/eval.rs:170-204)
    against the literal oracle, then fold and decide: is_sat_accumulation incl. the log-derivative check."""
    from sirius_b200 import lookup as PL
    from sirius_b200 import sangria as SG

    m, k = R.FR, 5
    n = 1 << k
    S, ao, co, fixed, _ = _structures(sb, vector, k, 3)
    rounds = 3 if vector else 2
    bases = oracle.running_bases(R.CURVE_BN256, 6 * n)
    ck = sb.CommitmentKey(R.CURVE_BN256, bases)
    traces = []
    for seed in (11, 12):
        _, advice = LC.columns(k, m, seed)
        chal = R.from_mont_limbs(oracle.random_field(R.FIELD_FR, 200 + seed, rounds), m)
        W, C, ch = PL.run_sps_protocol(S, [_mont(a, m) for a in advice], ck, lambda rnd, Ci, chal=chal: _mont([chal[rnd]], m)[0])
        S.is_sat(ck, np.stack(ch), C, W)
        traces.append((W, np.stack(ch)))
    (W1, c1), (W2, c2) = traces
    # relaxed accumulator = first trace with u = 1, E = 0 ; incoming = second trace
    u1 = _mont([1], m)
    T, commits = SG.VanillaFS.commit_cross_terms(ck, S, c1, u1, W1, c2, W2)
    St = E.Structure(k, [], fixed, LC.NUM_ADVICE, 1, co, m)
    exp = E.commit_cross_terms_eval(St, R.from_mont_limbs(c1, m), 1, [R.from_mont_limbs(w, m) for w in W1], R.from_mont_limbs(c2, m),
                                    [R.from_mont_limbs(w, m) for w in W2])
    assert len(T) == len(exp) == S.degree
    for j, (g, e) in enumerate(zip(T, exp)):
        assert R.from_mont_limbs(g, m) == e, f"T_{j+1}"
    r = oracle.random_field(R.FIELD_FR, 77, 1)
    ri = R.from_mont_limbs(r, m)[0]
    acc = SG.RelaxedPlonkWitness(R.FIELD_FR, W1, np.zeros((n, 4), dtype=np.uint64)).fold(W2, T, r[0])
    ch_acc = _mont([(a + ri * b) % m for a, b in zip(R.from_mont_limbs(c1, m), R.from_mont_limbs(c2, m))], m)
    u_acc = _mont([(1 + ri) % m], m)
    SG.VanillaFS.is_sat_accumulation(S, ch_acc, u_acc, acc.W, acc.E)
    bad = [w.copy() for w in acc.W]
    bad[-1][2, 0] ^= np.uint64(1)
    with pytest.raises((SG.EvaluationMismatch, SG.LogDerivativeNotSat)):
        SG.VanillaFS.is_sat_accumulation(S, ch_acc, u_acc, bad, acc.E)
    S.close()
    ck.close()
