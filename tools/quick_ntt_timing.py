"""Timing of the NTT entry points (CUDA events): forward, inverse (with the 2^-k scale) and coset forms, Gelt/s and the
algorithmic 64 B/element bandwidth.  SB_NTT_MODE=0 selects the stage-by-stage shared-memory pass, 1 (default) the radix-8
register pass.  Not the bench."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
import sirius_b200
from oracle import pyref as R
from sirius_b200 import _lib, fft

lib = sirius_b200.load()
ts = torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for _ in range(reps):
        fn()
    e1.record(ts)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"SB_NTT_MODE={os.environ.get('SB_NTT_MODE', '1')}")
for k in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["17", "20"])]:
    n = 1 << k
    a = oracle.random_field(R.FIELD_FR, 1, n)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    w, wi = fft.get_omega_or_inv(k, False), fft.get_omega_or_inv(k, True)
    div = fft.get_ifft_divisor(k)
    z, z2 = fft.fr_to_limbs(fft.FR_ZETA), fft.fr_to_limbs(fft.FR_ZETA * fft.FR_ZETA % fft.FR_MODULUS)

    def coset_fwd():
        _lib.check(lib.sb_coset_scale_device(0, ctypes.c_void_p(d.data_ptr()), n, z.ctypes.data_as(_lib.u64p), z2.ctypes.data_as(_lib.u64p), ctypes.c_void_p(ts.cuda_stream)))
        fft.ntt_device(d.data_ptr(), k, w, None, ts.cuda_stream)

    for name, fn in (("fft", lambda: fft.ntt_device(d.data_ptr(), k, w, None, ts.cuda_stream)),
                     ("ifft", lambda: fft.ntt_device(d.data_ptr(), k, wi, div, ts.cuda_stream)),
                     ("coset_fft", coset_fwd)):
        ms = timed(fn)
        print(f"{name:10s} k={k}: {ms * 1e3:8.1f} us  {n / ms / 1e6:7.3f} Gelt/s  {64 * n / ms / 1e6:8.1f} GB/s algorithmic", flush=True)
