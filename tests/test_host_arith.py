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
