"""GPU parity of the batched-affine reduction rounds in the commitment pipeline (affine.cuh / k_pair_round): every
round count and batch width must give the bit pattern of the CPU oracle, including on keys with identity and
repeated generators and on skewed scalars (one bucket holding most entries)."""
import numpy as np
import pytest

from oracle import pyref as R
from test_gpu_msm import _scalars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import sirius_b200

    sirius_b200.load()
    return sirius_b200


@pytest.fixture
def tune(sb):
    from sirius_b200 import _lib

    lib = _lib.load()

    def set_(rounds, B=16, sort=0):
        _lib.check(lib.sb_msm_tune(0, rounds))
        _lib.check(lib.sb_msm_tune(1, B))
        _lib.check(lib.sb_msm_tune(2, sort))

    yield set_
    _lib.check(lib.sb_msm_tune(0, -1))
    _lib.check(lib.sb_msm_tune(1, 16))
    _lib.check(lib.sb_msm_tune(2, 0))


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
@pytest.mark.parametrize("rounds,B", [(1, 16), (2, 8), (3, 16), (5, 8), (8, 16)])
def test_affine_rounds_small(sb, oracle, tune, curve, rounds, B):
    tune(rounds, B)
    for n, c in [(1, 0), (2, 0), (31, 0), (1000, 8), (4096, 6), (20000, 13)]:
        bases = oracle.running_bases(curve, n + 3)
        if n > 40:
            bases[17] = 0
            bases[19] = bases[18]
        ck = sb.CommitmentKey(curve, bases, window_bits=c)
        for kind in ("uniform", "edge", "equal"):
            s = _scalars(oracle, curve, n, 2000 + n, kind)
            assert np.array_equal(ck.commit(s), oracle.msm(curve, s, bases)), (n, c, kind)
        ck.close()


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
@pytest.mark.parametrize("rounds,B", [(2, 16), (3, 8), (4, 16)])
def test_affine_rounds_medium_skew_and_batch(sb, oracle, tune, curve, rounds, B):
    tune(rounds, B)
    n = 1 << 16
    bases = oracle.running_bases(curve, n)
    bases[17] = 0
    bases[19] = bases[18]
    ck = sb.CommitmentKey(curve, bases)
    for kind in ("uniform", "witness", "equal"):
        s = _scalars(oracle, curve, n, 77, kind)
        assert np.array_equal(ck.commit(s), oracle.msm(curve, s, bases)), kind
        assert np.array_equal(ck.commit(s[:12345]), oracle.msm(curve, s[:12345], bases)), kind
    vs = [_scalars(oracle, curve, 5000, 40 + j, kind) for j, kind in enumerate(["uniform", "witness", "equal", "edge", "uniform"])]
    vs[4][:] = 0
    got = ck.commit_batch(vs)
    for j, v in enumerate(vs):
        assert np.array_equal(got[j], oracle.msm(curve, v, bases)), j
    ck.close()


def test_affine_rounds_large(sb, oracle, tune):
    """2^20 scalars (BASELINE size) against the multithreaded CPU oracle, three round counts"""
    curve = R.CURVE_BN256
    n = 1 << 20
    bases = oracle.running_bases(curve, n)
    ck = sb.CommitmentKey(curve, bases)
    a = oracle.random_field(R.FIELD_FR, 5, n)
    exp = oracle.msm(curve, a, bases)
    for rounds, B in [(0, 16), (2, 16), (3, 8), (3, 16)]:
        tune(rounds, B)
        assert np.array_equal(ck.commit(a), exp), (rounds, B)
    ck.close()


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
@pytest.mark.parametrize("sort", [1, 2])
def test_sort_paths(sb, oracle, tune, curve, sort):
    """the per-entry atomic counting sort (1) and the two-level partition sort (2) forced on the same inputs:
    uniform, witness-like and all-equal scalars (one bucket per window holds everything), identity / repeated
    generators, prefix commits, batched commits with bucket counts that are not a power of two (5 * 512)"""
    tune(-1, 16, sort)
    n = 1 << 15
    bases = oracle.running_bases(curve, n)
    bases[17] = 0
    bases[19] = bases[18]
    for c in (10, 13):      # K = 512 (the smallest the partition sort takes) and 4096
        ck = sb.CommitmentKey(curve, bases, window_bits=c)
        for kind in ("uniform", "witness", "equal", "edge"):
            s = _scalars(oracle, curve, n, 91, kind)
            assert np.array_equal(ck.commit(s), oracle.msm(curve, s, bases)), (c, kind)
        s = _scalars(oracle, curve, n, 92, "uniform")
        assert np.array_equal(ck.commit(s[:777]), oracle.msm(curve, s[:777], bases))
        vs = [_scalars(oracle, curve, 3000, 50 + j, kind) for j, kind in enumerate(["uniform", "witness", "equal", "edge", "uniform"])]
        vs[4][:] = 0
        got = ck.commit_batch(vs)
        for j, v in enumerate(vs):
            assert np.array_equal(got[j], oracle.msm(curve, v, bases)), (c, j)
        ck.close()
    ck = sb.CommitmentKey(curve, bases, window_bits=6)   # K = 32: too few buckets, mode 2 falls back to the atomic path
    s = _scalars(oracle, curve, n, 93, "uniform")
    assert np.array_equal(ck.commit(s), oracle.msm(curve, s, bases))
    ck.close()


def test_large_commit_sort_paths_and_linearity():
    """2^22 bn256 scalars (BASELINE sweep size, window c = 17, 65 536 buckets): the partition sort and the counting sort
    agree bit for bit and commit(a) + commit(b) == commit(a + b) -- size-independent properties, no CPU MSM
    (tools/check_large_msm.py, also run by hand at 2^23: profiles/r1_large_msm_check.txt)"""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_large_msm.py"), "22"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sort paths agree: True" in r.stdout and "commit(a+b): True" in r.stdout
