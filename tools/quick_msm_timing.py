"""Scratch timing of sb_msm_device (CUDA events) for a few sizes / window widths.  Not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
import sirius_b200
from oracle import pyref as R

curve = R.CURVE_BN256
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["17", "20"])]
cs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nmax = 1 << max(sizes)
out = None
t0 = time.time()
bases = oracle.running_bases(curve, nmax)
scal = oracle.random_field(R.FIELD_FR, 1, nmax * batch)
print(f"inputs built in {time.time()-t0:.1f}s", flush=True)
d_s = torch.from_numpy(scal.view(np.int64)).cuda()
d_b = torch.from_numpy(bases.view(np.int64)).cuda()
out = torch.zeros(8 * batch, dtype=torch.int64, device="cuda")
for lg in sizes:
    n = 1 << lg
    for c in cs:
        t0 = time.time()
        ck = sirius_b200.CommitmentKey.from_device(curve, d_b.data_ptr(), n, window_bits=c)
        torch.cuda.synchronize()
        treg = time.time() - t0
        ts = torch.cuda.Stream()
        st = ts.cuda_stream
        for _ in range(2):
            ck.commit_batch_device(d_s.data_ptr(), n, n, batch, out.data_ptr(), 0, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(ts)
        for _ in range(reps):
            ck.commit_batch_device(d_s.data_ptr(), n, n, batch, out.data_ptr(), 0, st)
        e1.record(ts)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"n=2^{lg} batch={batch} c={ck.window_bits} register={treg:.2f}s msm={ms:.3f} ms  {batch*n/ms/1e3:.1f} Mscalar/s", flush=True)
        ck.close()
