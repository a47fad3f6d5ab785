"""GPU parity: CUDA field arithmetic and MSM through the C ABI against the CPU oracle (bit-exact)."""
import ctypes

import numpy as np
import pytest

from oracle import pyref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import sirius_b200

    sirius_b200.load()
    return sirius_b200


def p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_device_field_arithmetic(sb, oracle, field):
    from sirius_b200 import _lib

    m = R.MODULUS[field]
    n = 1 << 14
    a = oracle.random_field(field, 11, n)
    b = oracle.random_field(field, 12, n)
    edge = R.to_mont_limbs([0, 1, m - 1, m - 1, 2, (m - 1) // 2, m - 2, 3], m)
    a[:8] = edge
    b[:8] = edge[::-1]
    outs = [np.zeros_like(a) for _ in range(5)]
    _lib.check(_lib.load().sb_selftest_field(field, p(a), p(b), n, *[p(o) for o in outs]))
    exp_mul = oracle.field_binop("mul", field, a, b)
    assert np.array_equal(outs[1], exp_mul), "portable device mul"
    assert np.array_equal(outs[0], exp_mul), "PTX device mul"
    assert np.array_equal(outs[2], oracle.field_binop("add", field, a, b))
    assert np.array_equal(outs[3], oracle.field_binop("sub", field, a, b))
    nz = ~np.all(a == 0, axis=1)
    assert np.array_equal(outs[4][nz][:256], oracle.field_inv(field, a[nz][:256]))


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_device_lazy_domain(sb, field):
    """the PTX forms of the lazy-domain operations (field.cuh; used by the bucket kernel's mixed addition): operands
    anywhere in [0, 2p), results below 2p and congruent to the exact operation"""
    import random

    from sirius_b200 import _lib

    m = R.MODULUS[field]
    rng = random.Random(70 + field)
    edge = [0, 1, 2, m - 2, m - 1, m, m + 1, 2 * m - 2, 2 * m - 1, (1 << 254) - 1, 1 << 254, (1 << 32) - 1, 1 << 32, m + (1 << 128)]
    edge = [e for e in edge if e < 2 * m]
    va = edge * len(edge) + [rng.randrange(2 * m) for _ in range(4000)]
    vb = [e for e in edge for _ in edge] + [rng.randrange(2 * m) for _ in range(4000)]

    def limbs(vals):
        a = np.zeros((len(vals), 4), dtype=np.uint64)
        for i, v in enumerate(vals):
            for k in range(4):
                a[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
        return a

    def ints(a):
        return [sum(int(a[i, k]) << (64 * k) for k in range(4)) for i in range(a.shape[0])]

    a, b = limbs(va), limbs(vb)
    outs = [np.zeros_like(a) for _ in range(4)]
    _lib.check(_lib.load().sb_selftest_lazy(field, p(a), p(b), len(va), *[p(o) for o in outs]))
    rinv = pow(pow(2, 256, m), m - 2, m)
    for x, y, gm, gs, gd, gc in zip(va, vb, *[ints(o) for o in outs]):
        assert gm < 2 * m and gm % m == (x * y * rinv) % m, (hex(x), hex(y))
        assert gs < 2 * m and gs % m == (x - y) % m
        assert gd < 2 * m and gd % m == (2 * x) % m
        flag, val = gc >> 255, gc & ((1 << 255) - 1)
        assert val == x % m and bool(flag) == (x % m == 0)


def _scalars(oracle, curve, n, seed, kind):
    sf = 0 if curve == R.CURVE_BN256 else 1
    sm = R.CURVE_SCALAR[curve]
    s = oracle.random_field(sf, seed, n)
    if kind == "uniform":
        return s
    if kind == "edge":
        vals = [0, 1, sm - 1, 2, sm - 2, (sm - 1) // 2, (sm + 1) // 2, (1 << 253), (1 << 128) - 1, 1 << 16, (1 << 16) - 1]
        s[: len(vals)] = R.to_mont_limbs(vals, sm)[: min(len(vals), n)] if n >= len(vals) else s[:n]
        return s
    if kind == "witness":  # 60% zero, 20% < 2^8, 10% < 2^64, 10% uniform
        rng = np.random.default_rng(seed)
        sel = rng.random(n)
        small = rng.integers(0, 256, size=n)
        mid = rng.integers(0, 2**63, size=n)
        vals = R.to_mont_limbs([0], sm)
        out = s.copy()
        zero_rows = sel < 0.6
        out[zero_rows] = 0
        idx_small = np.where((sel >= 0.6) & (sel < 0.8))[0]
        if len(idx_small):
            table = R.to_mont_limbs(list(range(256)), sm)
            out[idx_small] = table[small[idx_small]]
        idx_mid = np.where((sel >= 0.8) & (sel < 0.9))[0]
        if len(idx_mid):
            out[idx_mid] = R.to_mont_limbs([int(v) for v in mid[idx_mid]], sm)
        del vals
        return out
    if kind == "equal":
        s[:] = s[0]
        return s
    raise ValueError(kind)


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
@pytest.mark.parametrize("n,c", [(0, 0), (1, 0), (2, 0), (31, 0), (1000, 8), (1000, 0), (4096, 11), (20000, 13)])
@pytest.mark.parametrize("kind", ["uniform", "edge"])
def test_msm_small(sb, oracle, curve, n, c, kind):
    n_ck = max(n, 1) + 3
    bases = oracle.running_bases(curve, n_ck)
    ck = sb.CommitmentKey(curve, bases, window_bits=c)
    s = _scalars(oracle, curve, n, 1000 + n, kind) if n else np.zeros((0, 4), dtype=np.uint64)
    got = ck.commit(s)
    exp = oracle.msm(curve, s, bases)
    assert np.array_equal(got, exp)
    ck.close()


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
@pytest.mark.parametrize("kind", ["uniform", "witness", "equal"])
def test_msm_medium_skew(sb, oracle, curve, kind):
    n = 1 << 16
    bases = oracle.running_bases(curve, n)
    bases[17] = 0  # identity generator
    bases[19] = bases[18]  # repeated generator
    ck = sb.CommitmentKey(curve, bases)
    s = _scalars(oracle, curve, n, 77, kind)
    assert np.array_equal(ck.commit(s), oracle.msm(curve, s, bases))
    # prefix commit (v.len() < ck.len(), src/commitment.rs:83)
    assert np.array_equal(ck.commit(s[:12345]), oracle.msm(curve, s[:12345], bases))
    ck.close()


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
def test_msm_batch(sb, oracle, curve):
    """the d cross-term commits of one commit_cross_terms call as one batched pipeline"""
    n = 5000
    bases = oracle.running_bases(curve, n + 5)
    ck = sb.CommitmentKey(curve, bases, window_bits=10)
    vs = [_scalars(oracle, curve, n, 40 + j, kind) for j, kind in enumerate(["uniform", "witness", "equal", "edge", "uniform"])]
    vs[4][:] = 0  # an all-zero cross term (T_d of a satisfied instance) -> identity commitment
    got = ck.commit_batch(vs)
    for j, v in enumerate(vs):
        assert np.array_equal(got[j], oracle.msm(curve, v, bases)), j
    assert not got[4].any()
    ck.close()


def test_msm_multiple_window_tables(sb, oracle):
    """a key with several registered window widths: every commit picks one per call, results never change"""
    curve = R.CURVE_GRUMPKIN
    n = 1 << 14
    bases = oracle.running_bases(curve, n)
    ck = sb.CommitmentKey(curve, bases, window_bits=8)
    s = _scalars(oracle, curve, n, 5, "uniform")
    ref_full, ref_small = oracle.msm(curve, s, bases), oracle.msm(curve, s[:300], bases)
    assert np.array_equal(ck.commit(s), ref_full)
    for wb in (13, 11, 16):
        ck.add_window(wb)
        assert np.array_equal(ck.commit(s), ref_full)
        assert np.array_equal(ck.commit(s[:300]), ref_small)
        got = ck.commit_batch([s, s[::-1].copy(), s])
        assert np.array_equal(got[0], ref_full) and np.array_equal(got[2], ref_full)
        assert np.array_equal(got[1], oracle.msm(curve, s[::-1].copy(), bases))
    ck.close()


def test_msm_public_bn256_vector(sb):
    """commit([2], [(1,2)]) on the GPU equals the public alt_bn128 doubling vector (tests/golden/bn256_g1_kat.json)"""
    import json, os

    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bn256_g1_kat.json")))
    exp = tuple(int(v) for v in kat["two_G_dec"])
    ck = sb.CommitmentKey(R.CURVE_BN256, R.points_to_limbs([tuple(kat["G"])], R.CURVE_BN256))
    assert R.limbs_to_points(ck.commit(R.to_mont_limbs([2], R.FR)), R.CURVE_BN256) == [exp]
    ck.close()


def test_key_cache_file_roundtrip_and_validation(sb, oracle, tmp_path):
    """reference commitment::file_tests::consistency (src/commitment.rs:197-213) + the on-curve check of
    load_or_setup_cache (:148-157), here done on the device"""
    for curve in (R.CURVE_BN256, R.CURVE_GRUMPKIN):
        k = 6
        bases = oracle.running_bases(curve, 1 << k)
        bases[5] = 0  # the identity is a valid key entry
        made = []
        ck = sb.CommitmentKey.load_or_setup_cache(curve, str(tmp_path), f"lbl{curve}", k, setup=lambda kk, lbl: (made.append(kk), bases)[1])
        assert made == [k]
        ck2 = sb.CommitmentKey.load_or_setup_cache(curve, str(tmp_path), f"lbl{curve}", k)  # now read back from the file
        assert np.array_equal(ck2._host, bases)
        s = oracle.random_field(0 if curve == R.CURVE_BN256 else 1, 3, 1 << k)
        assert np.array_equal(ck2.commit(s), ck.commit(s))
        # corrupt one coordinate: the loader must refuse the file
        path = tmp_path / f"lbl{curve}" / f"{k}.bin"
        raw = bytearray(path.read_bytes())
        raw[64 * 9 + 3] ^= 0x5A
        path.write_bytes(bytes(raw))
        with pytest.raises(ValueError, match="out of curve"):
            sb.CommitmentKey.load_or_setup_cache(curve, str(tmp_path), f"lbl{curve}", k)
        with pytest.raises(EOFError):
            sb.CommitmentKey.load_from_file(curve, str(path), k + 1)
        ck.close()
        ck2.close()


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_batch_invert(sb, oracle, field):
    """ff::BatchInvert semantics (zeros stay zero) vs the oracle's Fermat inversion"""
    from sirius_b200 import _lib

    for n in (1, 15, 16, 17, 1000, 1 << 15):
        a = oracle.random_field(field, 9 + n, n)
        a[::7] = 0
        out = np.zeros_like(a)
        _lib.check(_lib.load().sb_batch_invert(field, p(a), p(out), n))
        nz = ~np.all(a == 0, axis=1)
        assert not out[~nz].any()
        if nz.any():
            assert np.array_equal(out[nz], oracle.field_inv(field, a[nz]))


def test_msm_too_long_input(sb, oracle):
    bases = oracle.running_bases(R.CURVE_BN256, 8)
    ck = sb.CommitmentKey(R.CURVE_BN256, bases)
    with pytest.raises(sb.TooLongInput) as ei:
        ck.commit(oracle.random_field(R.FIELD_FR, 1, 9))
    assert ei.value.input_len == 9 and ei.value.limit == 8
    ck.close()


def test_msm_linearity_large(sb, oracle):
    """Size-independent property at a BASELINE size (2^20): commit(a) + commit(b) == commit(a+b), and
    a cancelling pair gives the identity."""
    curve = R.CURVE_BN256
    n = 1 << 20
    bases = oracle.running_bases(curve, n)
    ck = sb.CommitmentKey(curve, bases)
    a = oracle.random_field(R.FIELD_FR, 5, n)
    b = oracle.random_field(R.FIELD_FR, 6, n)
    ab = oracle.field_binop("add", R.FIELD_FR, a, b)
    ca, cb, cab = ck.commit(a), ck.commit(b), ck.commit(ab)
    assert oracle.is_on_curve(curve, cab)
    assert np.array_equal(oracle.point_add(curve, ca, cb), cab)
    neg_a = oracle.field_binop("sub", R.FIELD_FR, np.zeros_like(a), a)
    assert np.array_equal(oracle.point_add(curve, ca, ck.commit(neg_a)), np.zeros(8, dtype=np.uint64))
    # and against the multithreaded CPU oracle on the full size
    assert np.array_equal(ca, oracle.msm(curve, a, bases))
    ck.close()
