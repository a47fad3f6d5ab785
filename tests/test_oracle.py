"""Pins the CPU oracle (oracle/sirius_oracle.c) against the reference's own golden vectors and against
an independent big-int restatement (oracle/pyref.py)."""
import numpy as np
import pytest

from oracle import pyref as R

# src/fft.rs:242-251 -- fft([0..8)) over bn256 Fr
FFT_KAT = [
    28,
    68918385373930674424918168212551896122229959265833979749191472831399925654,
    17631683881184975370165255887551781615748388533673675138856,
    68918385373930639161550405842601155791718184162270748252414405484049647934,
    21888242871839275222246405745257275088548364400416034343698204186575808495613,
    21819324486465344583084855339414673932756646216253763595445789781091758847675,
    21888242871839275204614721864072299718383108512864252727949815652902133356753,
    21819324486465344547821487577044723192426134441150200363949012713744408569955,
]


def test_fft_kat_pyref():
    assert R.fft(list(range(8))) == FFT_KAT


def test_fft_kat_c(oracle):
    a = R.to_mont_limbs(list(range(8)), R.FR)
    for threads in (1, 4):
        out = oracle.fft(a, threads=threads)
        assert R.from_mont_limbs(out, R.FR) == FFT_KAT


def test_lagrange_kat_golden():
    import json, os

    path = os.path.join(os.path.dirname(__file__), "golden", "lagrange_kat.json")
    kat = json.load(open(path))
    got = R.eval_lagrange_polys(kat["log_n"], kat["X"])
    assert [int(v, 16) for v in kat["expected_hex"]] == got


@pytest.mark.parametrize("k", [1, 2, 4, 5, 8])
def test_fft_c_vs_naive_dft(oracle, k):
    rng = R.Xoshiro256ss(0x5349524955530000 + k)
    vals = [rng.field(R.FR) for _ in range(1 << k)]
    a = R.to_mont_limbs(vals, R.FR)
    assert R.from_mont_limbs(oracle.fft(a), R.FR) == R.fft(vals)
    assert R.from_mont_limbs(oracle.fft(a, threads=4), R.FR) == R.fft(vals)
    assert R.from_mont_limbs(oracle.ifft(a), R.FR) == R.ifft(vals)
    assert R.from_mont_limbs(oracle.coset_fft(a), R.FR) == R.coset_fft(vals)
    assert R.from_mont_limbs(oracle.coset_ifft(a), R.FR) == R.coset_ifft(vals)


@pytest.mark.parametrize("k", [4, 5, 6, 7, 8, 12])
def test_fft_roundtrip(oracle, k):
    # src/fft.rs:268-296 fft_random_input_test / coset_fft_random_input_test
    a = oracle.random_field(R.FIELD_FR, 77 + k, 1 << k)
    assert np.array_equal(oracle.ifft(oracle.fft(a, threads=2), threads=2), a)
    assert np.array_equal(oracle.coset_ifft(oracle.coset_fft(a)), a)


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_field_ops_vs_bigint(oracle, field):
    m = R.MODULUS[field]
    rng = R.Xoshiro256ss(1234 + field)
    av = [rng.field(m) for _ in range(64)] + [0, 1, m - 1, m - 1, 0]
    bv = [rng.field(m) for _ in range(64)] + [0, m - 1, m - 1, 1, 5]
    a, b = R.to_mont_limbs(av, m), R.to_mont_limbs(bv, m)
    assert R.from_mont_limbs(oracle.field_binop("mul", field, a, b), m) == [x * y % m for x, y in zip(av, bv)]
    assert R.from_mont_limbs(oracle.field_binop("add", field, a, b), m) == [(x + y) % m for x, y in zip(av, bv)]
    assert R.from_mont_limbs(oracle.field_binop("sub", field, a, b), m) == [(x - y) % m for x, y in zip(av, bv)]
    nz = [v for v in av if v]
    assert R.from_mont_limbs(oracle.field_inv(field, R.to_mont_limbs(nz, m)), m) == [pow(v, -1, m) for v in nz]


def test_random_field_matches_pyref(oracle):
    for field in (R.FIELD_FR, R.FIELD_FQ):
        m = R.MODULUS[field]
        rng = R.Xoshiro256ss(42)
        exp = [rng.field(m) for _ in range(16)]
        assert R.from_mont_limbs(oracle.random_field(field, 42, 16), m) == exp


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
def test_generators_and_running_bases(oracle, curve):
    assert R.is_on_curve(R.CURVE_GEN[curve], curve)
    bases = oracle.running_bases(curve, 9)
    assert R.limbs_to_points(bases, curve) == R.running_bases(9, curve)
    for row in bases:
        assert oracle.is_on_curve(curve, row)


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
@pytest.mark.parametrize("n", [0, 1, 3, 5, 33, 100])
def test_msm_vs_bigint(oracle, curve, n):
    """commit == sum v_i * ck_i (src/commitment.rs:81-90), checked against the affine big-int group law."""
    sm = R.CURVE_SCALAR[curve]
    rng = R.Xoshiro256ss(0xC0FFEE + n + curve)
    sv = [rng.field(sm) for _ in range(n)]
    if n >= 5:
        sv[0], sv[1], sv[2] = 0, 1, sm - 1
    pts = R.running_bases(n, curve)
    if n >= 33:
        pts[7] = None  # identity base (CommitmentKey::default_value)
        pts[9] = pts[8]  # repeated base -> doubling inside a bucket
        sv[9] = sv[8]
    exp = R.msm_naive(sv, pts, curve)
    s, b = R.to_mont_limbs(sv, sm), R.points_to_limbs(pts, curve)
    for threads in (1, 3, 8):
        got = oracle.msm(curve, s, b, threads=threads)
        assert R.limbs_to_points(got, curve) == [exp]
    assert R.limbs_to_points(oracle.msm_naive(curve, s, b), curve) == [exp]


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
def test_msm_pippenger_vs_naive_medium(oracle, curve):
    n = 3000
    s = oracle.random_field(0 if curve == 0 else 1, 99, n)
    b = oracle.running_bases(curve, n)
    assert np.array_equal(oracle.msm(curve, s, b, threads=4), oracle.msm_naive(curve, s, b))


def test_msm_cancellation_gives_identity(oracle):
    curve = R.CURVE_BN256
    pts = R.running_bases(2, curve)
    b = R.points_to_limbs([pts[0], pts[0]], curve)
    s = R.to_mont_limbs([5, R.FR - 5], R.FR)
    out = oracle.msm(curve, s, b)
    assert not out.any()  # identity encoded (0,0)


def test_bn256_public_doubling_vector(oracle):
    """An external anchor for the bn256 group law (the reference itself has no known-answer test for commit):
    2*(1,2) from the public alt_bn128 / EIP-196 vectors, and the group order r*(1,2) = identity."""
    import json, os

    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bn256_g1_kat.json")))
    G = tuple(kat["G"])
    exp = tuple(int(v) for v in kat["two_G_dec"])
    assert R.ec_add(G, G, R.CURVE_BN256) == exp
    b = R.points_to_limbs([G], R.CURVE_BN256)
    two = R.to_mont_limbs([2], R.FR)
    assert R.limbs_to_points(oracle.msm(R.CURVE_BN256, two, b), R.CURVE_BN256) == [exp]
    assert R.limbs_to_points(oracle.msm_naive(R.CURVE_BN256, two, b), R.CURVE_BN256) == [exp]
    # (r-1)*G == -G  <=>  r*G == identity, on both curves of the cycle
    for curve in (R.CURVE_BN256, R.CURVE_GRUMPKIN):
        g = R.CURVE_GEN[curve]
        m1 = R.to_mont_limbs([R.CURVE_SCALAR[curve] - 1], R.CURVE_SCALAR[curve])
        got = oracle.msm(curve, m1, R.points_to_limbs([g], curve))
        assert R.limbs_to_points(got, curve) == [R.ec_neg(g, curve)]


def _bn256_kat():
    import json, os

    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bn256_g1_kat.json")))


def test_msm_oracle_public_precompile_vectors(oracle):
    """The C oracle's commit on the public EIP-196 bn256Add / bn256ScalarMul known answers (tests/golden/
    bn256_g1_kat.json): commit([k], [P]) = k*P and commit([1, 1], [P, Q]) = P + Q.  These pin the MSM oracle to
    vectors that come from outside both this repository and the reference."""
    kat = _bn256_kat()
    C = R.CURVE_BN256
    h = lambda s: int(s, 16)
    for v in kat["scalar_mul"]:
        P, exp = (h(v["x"]), h(v["y"])), (h(v["ex"]), h(v["ey"]))
        k = R.to_mont_limbs([h(v["k"]) % R.FR], R.FR)
        b = R.points_to_limbs([P], C)
        assert R.limbs_to_points(oracle.msm(C, k, b), C) == [exp], v["name"]
        assert R.limbs_to_points(oracle.msm_naive(C, k, b), C) == [exp], v["name"]
    ones = R.to_mont_limbs([1, 1], R.FR)
    for v in kat["add"]:
        P, Q, exp = (h(v["x1"]), h(v["y1"])), (h(v["x2"]), h(v["y2"])), (h(v["ex"]), h(v["ey"]))
        assert R.limbs_to_points(oracle.msm(C, ones, R.points_to_limbs([P, Q], C)), C) == [exp], v["name"]
    # a 6-term commitment built from the vectors: sum k_i * P_i with the answers combined by the big-int group law
    pts = [(h(v["x"]), h(v["y"])) for v in kat["scalar_mul"]]
    ks = [h(v["k"]) % R.FR for v in kat["scalar_mul"]]
    exp = None
    for v in kat["scalar_mul"]:
        exp = R.ec_add(exp, (h(v["ex"]), h(v["ey"])), C)
    assert R.limbs_to_points(oracle.msm(C, R.to_mont_limbs(ks, R.FR), R.points_to_limbs(pts, C)), C) == [exp]
