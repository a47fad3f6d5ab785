"""The portable (host-compilable) code paths of sirius_b200/csrc/{field,curve}.cuh against the oracle.
The PTX paths are compared with these same portable paths on the GPU (tests/test_gpu_field.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pyref as R

HERE = os.path.dirname(os.path.abspath(__file__))
u64p = ctypes.POINTER(ctypes.c_uint64)


@pytest.fixture(scope="module")
def ha():
    src = os.path.join(HERE, "host", "host_arith.cpp")
    so = os.path.join(HERE, "host", "libhost_arith.so")
    deps = [src] + [os.path.join(HERE, "..", "sirius_b200", "csrc", f) for f in ("field.cuh", "curve.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    return ctypes.CDLL(so)


def p(a):
    return a.ctypes.data_as(u64p)


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_field_portable(ha, oracle, field):
    m = R.MODULUS[field]
    n = 512
    a = oracle.random_field(field, 5, n)
    b = oracle.random_field(field, 6, n)
    edge = R.to_mont_limbs([0, 1, m - 1, m - 1, 2, (m - 1) // 2], m)
    a[:6] = edge
    b[:6] = edge[::-1]
    o = np.zeros_like(a)
    ha.ha_mul(field, p(a), p(b), p(o), n)
    assert np.array_equal(o, oracle.field_binop("mul", field, a, b))
    oa, os_ = np.zeros_like(a), np.zeros_like(a)
    ha.ha_addsub(field, p(a), p(b), p(oa), p(os_), n)
    assert np.array_equal(oa, oracle.field_binop("add", field, a, b))
    assert np.array_equal(os_, oracle.field_binop("sub", field, a, b))
    nz = a[8:40].copy()
    oi = np.zeros_like(nz)
    ha.ha_inv(field, p(nz), p(oi), nz.shape[0])
    assert np.array_equal(oi, oracle.field_inv(field, nz))


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_inv_safegcd_portable(ha, oracle, field):
    """the divstep inversion (field.cuh inv_safegcd) against Python pow on edge values and 20k random elements"""
    m = R.MODULUS[field]
    edge = [0, 1, 2, 3, m - 1, m - 2, (m - 1) // 2, (m + 1) // 2, 1 << 30, (1 << 30) - 1, 1 << 253, (1 << 253) - 1, 1 << 128,
            pow(2, 256, m), pow(2, 512, m), m - pow(2, 256, m)] + [1 << s for s in range(0, 254, 7)] + [m - (1 << s) for s in range(0, 254, 11)]
    vals = edge + R.from_mont_limbs(oracle.random_field(field, 99, 20000), m)
    # the routine inverts the stored Montgomery residue: pass v as the residue of V = v * R^-1
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(4):
            a[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    o = np.zeros_like(a)
    ha.ha_inv_safegcd(field, p(a), p(o), len(vals))
    Rm = pow(2, 256, m)
    for i, v in enumerate(vals):
        got = sum(int(o[i, k]) << (64 * k) for k in range(4))
        # stored v = V*R  ->  expected stored result V^-1 * R = v^-1 * R^2
        exp = (pow(v, m - 2, m) * Rm * Rm) % m if v else 0
        assert got == exp, (i, hex(v))


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
def test_xyzz_group_law(ha, oracle, curve):
    pts = R.running_bases(12, curve)
    cases = [
        (pts[:8], [0] * 8),
        (pts[:8], [1, 0, 1, 0, 0, 1, 1, 0]),
        ([pts[0], pts[0]], [0, 0]),  # doubling through madd
        ([pts[0], pts[0]], [0, 1]),  # cancellation -> identity
        ([pts[2], pts[3], pts[2], pts[3]], [0, 0, 0, 0]),  # lo == hi -> doubling through add
        ([pts[2], pts[3], pts[2], pts[3]], [0, 0, 1, 1]),  # lo == -hi -> identity through add
        ([None, pts[1], None, pts[4]], [0, 0, 0, 1]),
        ([None, None], [0, 0]),
    ]
    for plist, negs in cases:
        exp = None
        for q, s in zip(plist, negs):
            exp = R.ec_add(exp, R.ec_neg(q, curve) if s else q, curve)
        arr = R.points_to_limbs(plist, curve)
        ng = np.array(negs, dtype=np.uint8)
        o1, o2 = np.zeros(8, dtype=np.uint64), np.zeros(8, dtype=np.uint64)
        ha.ha_points(curve, p(arr), ng.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), len(plist), p(o1), p(o2))
        assert R.limbs_to_points(o1, curve) == [exp]
        assert R.limbs_to_points(o2, curve) == [R.ec_add(exp, exp, curve)]


def _to_limbs(vals):
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(4):
            a[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return a


def _from_limbs(a):
    return [sum(int(a[i, k]) << (64 * k) for k in range(4)) for i in range(a.shape[0])]


@pytest.mark.parametrize("field", [R.FIELD_FR, R.FIELD_FQ])
def test_lazy_domain_ops(ha, field):
    """field.cuh lazy domain: operands anywhere in [0, 2p) (incl. p, 2p - 1 and values just around p), results must
    stay in [0, 2p) and be congruent to the canonical operation"""
    import random

    m = R.MODULUS[field]
    rng = random.Random(7 + field)
    edge = [0, 1, 2, m - 2, m - 1, m, m + 1, 2 * m - 2, 2 * m - 1, (1 << 254) - 1, 1 << 254, (1 << 254) + 1, (1 << 32) - 1, 1 << 32,
            m + (1 << 128), 2 * m - (1 << 200)]
    edge = [e for e in edge if e < 2 * m]
    vals_a = edge * len(edge) + [rng.randrange(2 * m) for _ in range(20000)]
    vals_b = [e for e in edge for _ in edge] + [rng.randrange(2 * m) for _ in range(20000)]
    a, b = _to_limbs(vals_a), _to_limbs(vals_b)
    om, os_, od = np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)
    oz = np.zeros(len(vals_a), dtype=np.uint8)
    ha.ha_lazy(field, p(a), p(b), p(om), p(os_), p(od), oz.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), len(vals_a))
    rinv = pow(pow(2, 256, m), m - 2, m)
    for x, y, gm, gs, gd, gz in zip(vals_a, vals_b, _from_limbs(om), _from_limbs(os_), _from_limbs(od), oz):
        assert gm < 2 * m and gm % m == (x * y * rinv) % m, (hex(x), hex(y))
        assert gs < 2 * m and gs % m == (x - y) % m
        assert gd < 2 * m and gd % m == (2 * x) % m
        assert bool(gz) == (x % m == 0)


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
def test_xyzz_madd_lazy(ha, oracle, curve):
    """xyzz_madd_lazy (the bucket kernel's addition) over long runs with repeats, cancellations and identities: same
    affine sum as the oracle, every coordinate below 2p at every step"""
    import random

    rng = random.Random(11 + curve)
    pts = R.running_bases(24, curve)
    runs = [
        [(i, 0) for i in range(24)],
        [(3, 0), (3, 0), (3, 0), (3, 1), (3, 1), (3, 1), (5, 0)],          # doubling, then cancel to the identity, restart
        [(None, 0), (7, 1), (None, 0), (7, 0), (7, 0)],
        [(rng.randrange(24), rng.randrange(2)) for _ in range(3000)],
    ]
    for run in runs:
        plist = [pts[i] if i is not None else None for i, _ in run]
        negs = np.array([s for _, s in run], dtype=np.uint8)
        exp = None
        for q, s in zip(plist, negs):
            exp = R.ec_add(exp, R.ec_neg(q, curve) if (s and q is not None) else q, curve)
        arr = R.points_to_limbs(plist, curve)
        out = np.zeros(8, dtype=np.uint64)
        ok = ctypes.c_uint8(0)
        ha.ha_points_lazy(curve, p(arr), negs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), len(plist), p(out), ctypes.byref(ok))
        assert ok.value == 1
        assert R.limbs_to_points(out, curve) == [exp]
