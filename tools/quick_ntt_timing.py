"""Scratch timing of sb_ntt_device (CUDA events).  Not the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
import sirius_b200
from sirius_b200 import fft
from oracle import pyref as R

for k in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["17", "20"])]:
    a = oracle.random_field(R.FIELD_FR, 1, 1 << k)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    ts = torch.cuda.Stream()
    w = fft.get_omega_or_inv(k, False)
    for _ in range(3):
        fft.ntt_device(d.data_ptr(), k, w, None, ts.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record(ts)
    for _ in range(reps):
        fft.ntt_device(d.data_ptr(), k, w, None, ts.cuda_stream)
    e1.record(ts)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = 1 << k
    print(f"ntt k={k}: {ms*1e3:.1f} us  {n/ms/1e6:.3f} Gelt/s  {64*n/ms/1e6:.1f} GB/s algorithmic", flush=True)
