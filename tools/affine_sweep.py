"""Timing sweep of the batched-affine reduction rounds (sb_msm_tune) at the bench's commit shapes:
rounds 0..4 x outputs-per-thread 8/16 for the primary W commit (1 572 864 scalars), the secondary W commit
(917 504) and the batched cross-term commits (6 x 131 072), synthetic key of 2^21 points with the bench's window
widths.  Every configuration is checked bit-for-bit against rounds = 0.  Not the bench."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sirius_b200 import _lib, device

lib = _lib.load()
_lib.check(lib.sb_init(0))
stream = torch.cuda.Stream()
st = stream.cuda_stream
curve = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rounds_list = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,3,4".split(","))]
n_key = 12 << 17
ck = device.synthetic_key(curve, n_key, window_bits=16, stream=stream)
for wb in (13, 15, 17):
    ck.add_window(wb, st)
stream.synchronize()
lib.sb_profile_enable(0)
shapes = [("W primary", 12 << 17, 1), ("W secondary", 7 << 17, 1), ("cross terms x6", 1 << 17, 6), ("cross terms x5", 1 << 17, 5), ("msm 2^20", 1 << 20, 1)]
NT = 10
for name, n, batch in shapes:
    s = device.random_field_device(n * batch, 0x5349 + n)
    out = torch.zeros(8 * batch, dtype=torch.int64, device="cuda")
    ref = None
    for rounds in rounds_list:
        for B in ((16,) if rounds == 0 else (8, 16)):
            _lib.check(lib.sb_msm_tune(0, rounds))
            _lib.check(lib.sb_msm_tune(1, B))
            for _ in range(2):
                ck.commit_batch_device(s.data_ptr(), n, n, batch, out.data_ptr(), 0, st)
            stream.synchronize()
            got = out.cpu().clone()
            if ref is None:
                ref = got
            ok = bool(torch.equal(got, ref))
            ms_arr, un_arr, ln_arr = (ctypes.c_double * NT)(), (ctypes.c_uint64 * NT)(), (ctypes.c_uint64 * NT)()
            lib.sb_profile_enable(1)
            lib.sb_profile_collect(ms_arr, un_arr, ln_arr)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record(stream)
            for _ in range(reps):
                ck.commit_batch_device(s.data_ptr(), n, n, batch, out.data_ptr(), 0, st)
            e1.record(stream)
            torch.cuda.synchronize()
            lib.sb_profile_collect(ms_arr, un_arr, ln_arr)
            lib.sb_profile_enable(0)
            ms = e0.elapsed_time(e1) / reps
            parts = " ".join(f"{t}={ms_arr[i] / reps:.3f}" for i, t in enumerate(["dec", "sort", "acc", "fix", "red", "fin"]))
            print(f"{name:16s} n={n:8d} x{batch} rounds={rounds} B={B:2d}: {ms:7.3f} ms  [{parts}]  {'same' if ok else 'MISMATCH'}", flush=True)
_lib.check(lib.sb_msm_tune(0, -1))
