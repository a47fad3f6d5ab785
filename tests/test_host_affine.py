"""The batched-affine reduction rounds of the MSM bucket sums (sirius_b200/csrc/affine.cuh) emulated on the host:
the kernel's own per-thread phase functions and product-tree steps, run thread by thread, against plain XYZZ bucket
sums and the big-int oracle.  Covers the exceptional cases a round can meet (identity operands, P + P, P + (-P),
odd bucket sizes, empty buckets, outputs spanning several blocks) for every round count the pipeline may choose."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pyref as R

HERE = os.path.dirname(os.path.abspath(__file__))
u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)
NEG = 0x80000000


@pytest.fixture(scope="module")
def haf():
    src = os.path.join(HERE, "host", "host_affine.cpp")
    so = os.path.join(HERE, "host", "libhost_affine.so")
    deps = [src] + [os.path.join(HERE, "..", "sirius_b200", "csrc", f) for f in ("field.cuh", "curve.cuh", "affine.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.haf_bucket_sums.restype = ctypes.c_long
    return lib


def run(haf, curve, table, buckets, rounds, B):
    """buckets: list of lists of entries (table index | NEG).  Returns (out, ref, stats)."""
    KB = len(buckets)
    off = np.zeros(KB + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(b) for b in buckets])
    eidx = np.array([e for b in buckets for e in b] + [0], dtype=np.uint32)
    out = np.zeros((KB, 8), dtype=np.uint64)
    ref = np.zeros((KB, 8), dtype=np.uint64)
    stats = (ctypes.c_long * 8)()
    oob = haf.haf_bucket_sums(curve, table.ctypes.data_as(u64p), eidx.ctypes.data_as(u32p), off.ctypes.data_as(u32p), KB, rounds, B,
                              out.ctypes.data_as(u64p), ref.ctypes.data_as(u64p), stats)
    assert oob == 0, "a round wrote past its buffer bound"
    return out, ref, list(stats)


@pytest.mark.parametrize("curve", [R.CURVE_BN256, R.CURVE_GRUMPKIN])
def test_exceptional_cases_against_oracle(haf, curve):
    pts = R.running_bases(10, curve)
    plist = pts + [None]            # index 10 = identity generator (0,0)
    table = R.points_to_limbs(plist, curve).reshape(-1, 8)
    ID = 10
    buckets = [
        [],                                  # empty
        [3],                                 # single
        [3 | NEG],                           # single, negated
        [1, 2],                              # generic pair
        [1, 1],                              # doubling in round 0
        [1, 1 | NEG],                        # cancellation in round 0
        [1, 2, 1, 2],                        # doubling in round 1
        [1, 2, 1 | NEG, 2 | NEG],            # cancellation in round 1
        [1, 2, 1 | NEG, 2 | NEG, 4],         # identity + point in round 2
        [ID, 5],                             # identity first
        [5, ID],                             # identity second
        [ID, ID],                            # both identity
        [ID, ID, ID, 7 | NEG, ID],           # identities around a point
        [0, 1, 2, 3, 4, 5, 6, 7, 8, 9],
        [0, 1, 2, 3, 4, 5, 6, 7, 8],         # odd count
        [4] * 8,                             # 8 * P by repeated doubling
        [4] * 7 + [4 | NEG] * 7,             # cancels to the identity in the last round
        [2, 2, 2],                           # 2P + P
    ]
    exp = []
    for b in buckets:
        acc = None
        for e in b:
            q = plist[e & 0x7fffffff]
            acc = R.ec_add(acc, R.ec_neg(q, curve) if (e & NEG and q is not None) else q, curve)
        exp.append(acc)
    for rounds in (0, 1, 2, 3, 4, 8):
        for B in (8, 16):
            out, ref, stats = run(haf, curve, table, buckets, rounds, B)
            got = R.limbs_to_points(out.reshape(-1), curve)
            assert got == exp, (rounds, B)
            assert np.array_equal(out, ref)
            if rounds >= 3:
                assert stats[5] > 0 and stats[3] > 0 and stats[2] > 0  # tangent, identity and copy-second slots all exercised


@pytest.mark.parametrize("curve,B", [(R.CURVE_BN256, 16), (R.CURVE_GRUMPKIN, 8)])
def test_random_buckets_span_blocks(haf, oracle, curve, B):
    """~30k entries over 600 buckets (several 256 x B output blocks per round), skewed sizes, duplicates, signs"""
    rng = np.random.default_rng(1234 + curve)
    n = 512
    table = oracle.running_bases(curve, n).reshape(-1, 8).copy()
    table[17] = 0
    table[400] = 0          # identity generators
    table[33] = table[32]   # repeated generator
    sizes = rng.poisson(40, 600)
    sizes[::50] = 0
    sizes[7] = 3000         # one heavy bucket
    sizes[8] = 1
    buckets = []
    for s in sizes:
        e = rng.integers(0, n, int(s)).astype(np.uint32)
        e |= (rng.integers(0, 2, int(s)).astype(np.uint32) << 31)
        buckets.append([int(x) for x in e])
    buckets[9] = [5] * 64 + [5 | NEG] * 63     # nearly cancelling run of one point
    for rounds in (1, 2, 3, 6):
        out, ref, stats = run(haf, curve, table, buckets, rounds, B)
        assert np.array_equal(out, ref), rounds
        assert stats[4] > 10000 >> (6 - min(rounds, 6))  # generic additions dominate
