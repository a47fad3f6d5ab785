"""The 4-warp cooperative XYZZ addition / doubling (csrc/coop.cuh) that the commitment tail runs on: every exceptional
case (generic, P + P through the addition, doubling, P + (-P), identity on either side) against the single-lane forms on
the device and against Python big-integer affine arithmetic (oracle/pyref.py), both curves."""
import numpy as np
import pytest

from oracle import pyref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
def test_coop_group_law(oracle, curve):
    from sirius_b200 import _lib

    lib = _lib.load()
    n = 75   # not a multiple of 32: the last block runs with idle lanes
    bases = oracle.running_bases(curve, n + 5)
    a = bases[:n].copy()
    b = bases[3:n + 3].copy()
    # exceptional inputs: a == b (P + P at the first step), b == -a, identities
    b[5] = a[5]
    neg = R.limbs_to_points(a[6:7], curve)[0]
    b[6] = R.points_to_limbs([R.ec_neg(neg, curve)], curve)[0]
    a[7] = 0
    b[8] = 0
    a[9] = 0
    b[9] = 0
    out_c = np.zeros((n, 3, 8), dtype=np.uint64)
    out_p = np.zeros((n, 3, 8), dtype=np.uint64)
    _lib.check(lib.sb_selftest_coop(curve, np.ascontiguousarray(a).ctypes.data_as(_lib.u64p), np.ascontiguousarray(b).ctypes.data_as(_lib.u64p), n,
                                    out_c.ctypes.data_as(_lib.u64p), out_p.ctypes.data_as(_lib.u64p)))
    assert np.array_equal(out_c, out_p)
    pa, pb = R.limbs_to_points(a, curve), R.limbs_to_points(b, curve)
    for i in range(n):
        t = R.ec_add(pa[i], pb[i], curve)
        t = R.ec_add(t, t, curve)
        t = R.ec_add(t, t, curve)
        t = R.ec_add(t, pb[i], curve)
        u = R.ec_add(t, t, curve)   # (identity + t) doubled
        exp = R.points_to_limbs([t, u, t], curve)
        assert np.array_equal(out_c[i], exp), i
