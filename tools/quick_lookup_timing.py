"""Scratch timing of the lookup / decider kernels (CUDA events on the launching stream).  Not the bench."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sirius_b200
from sirius_b200 import _lib
from sirius_b200.device import random_field_device

lib = _lib.load()
FR = 0
ts = torch.cuda.Stream()
st = ts.cuda_stream


def timed(name, fn, n, bytes_per_cell, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(reps):
        fn()
    e1.record(ts)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} n=2^{n.bit_length()-1}: {ms*1e3:9.1f} us  {n/ms/1e6:8.3f} Gcell/s  {bytes_per_cell*n/ms/1e6:8.1f} GB/s algorithmic", flush=True)


for k in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["17", "20"])]:
    n = 1 << k
    pool = random_field_device(1 << 12, 5)
    g = torch.Generator(device="cuda")
    g.manual_seed(k)
    # table: 2^12 distinct values then padding with row 0 (as halo2 pads); lookups: 3/4 real rows, 1/4 unused (value of row 0)
    t_idx = torch.arange(n, device="cuda") % (1 << 12)
    t_idx[n // 2:] = 0
    l_idx = torch.randint(0, 1 << 12, (n,), device="cuda", generator=g)
    l_idx[3 * n // 4:] = 0
    t = pool[t_idx].contiguous()
    l = pool[l_idx].contiguous()
    m = torch.zeros_like(t)
    h = torch.zeros_like(t)
    gg = torch.zeros_like(t)
    out = torch.zeros(4, dtype=torch.int64, device="cuda")
    r = np.array([0x1234567, 2, 3, 4], dtype=np.uint64)
    timed("evaluate_m (3 kernels)", lambda: _lib.check(lib.sb_lookup_multiplicity_device(FR, l.data_ptr(), n, t.data_ptr(), n, m.data_ptr(), st)), n, 96)
    timed("evaluate_h_g (2 kernels)", lambda: _lib.check(lib.sb_lookup_inverses_device(FR, l.data_ptr(), t.data_ptr(), m.data_ptr(), r.ctypes.data_as(_lib.u64p), n, h.data_ptr(), gg.data_ptr(), st)), n, 160)
    timed("batch invert (reference point)", lambda: _lib.check(lib.sb_batch_invert_device(FR, l.data_ptr(), h.data_ptr(), n, st)), n, 64)
    timed("sum(h - g)", lambda: _lib.check(lib.sb_sum_diff_device(FR, h.data_ptr(), gg.data_ptr(), n, out.data_ptr(), st)), n, 64)
    # permutation decider at the Sangria primary shape: N = 12 columns * 2^k rows, unit entries, 1/3 of the cells in copy cycles
    A = 12 if k <= 17 else 2
    N = A * n
    perm = np.arange(N, dtype=np.uint64)
    rng = np.random.default_rng(1)
    cells = rng.permutation(N)[: N // 3].astype(np.uint64)
    perm[cells] = np.roll(cells, 1)
    rows = np.arange(N, dtype=np.uint64)
    one = np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], dtype=np.uint64)  # R mod r (bn256 Fr one)
    vals = np.tile(one, (N, 1))
    hP = ctypes.c_void_p()
    _lib.check(lib.sb_sparse_register(FR, rows.ctypes.data_as(_lib.u64p), perm.ctypes.data_as(_lib.u64p), vals.ctypes.data_as(_lib.u64p), N, N, ctypes.byref(hP)))
    Z = random_field_device(N, 3)
    cnt = ctypes.c_uint64(0)
    timed(f"is_sat_permutation (N={A}*2^{k}, incl. sync)", lambda: _lib.check(lib.sb_sparse_mismatch_device(hP, None, 0, Z.data_ptr(), N, ctypes.cast(ctypes.byref(cnt), _lib.u64p), st)), N, 32 * 3 + 8)
    lib.sb_sparse_release(hP)
    del Z
