"""Pedersen MSM size sweep (BASELINE.json config 4): bn256 G1, n = 2^16..2^24, uniform (U) and witness-like (W)
scalars, one B200.  Every size is checked bit-for-bit against the CPU oracle, then timed with CUDA events.
Usage: python tools/msm_sweep.py [sizes=16,18,20,22,24] [check=1]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
import sirius_b200
from oracle import pyref as R
from sirius_b200 import _lib, curves, device

sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["16", "18", "20", "22", "24"])]
check = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = sirius_b200.load()
curve = R.CURVE_BN256
st = torch.cuda.Stream()
rows = []
for lg in sizes:
    n = 1 << lg
    d_b = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
    g = curves.generator_limbs(curve)
    _lib.check(lib.sb_index_multiples_device(curve, g.ctypes.data_as(_lib.u64p), 0, n, ctypes.c_void_p(d_b.data_ptr()), ctypes.c_void_p(st.cuda_stream)))
    st.synchronize()
    t0 = time.time()
    ck = sirius_b200.CommitmentKey.from_device(curve, d_b.data_ptr(), n, stream=st.cuda_stream)
    st.synchronize()
    treg = time.time() - t0
    U = device.random_field_device(n, 0x5349524955530000 + lg)
    sel = torch.rand(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(lg))
    Wl = U.clone()
    Wl[sel < 0.6] = 0
    small = (sel >= 0.6) & (sel < 0.8)
    # "small" values: integers < 2^8 in Montgomery form = value * R mod r ; build via a lookup of 256 residues
    table = torch.from_numpy(R.to_mont_limbs(list(range(256)), R.FR).view(np.int64)).cuda()
    idx = torch.randint(0, 256, (int(small.sum().item()),), device="cuda")
    Wl[small] = table[idx]
    out = torch.zeros(8, dtype=torch.int64, device="cuda")
    res = {"log_n": lg, "window_bits": ck.window_bits, "register_s": round(treg, 2)}
    for name, s in (("U", U), ("W", Wl)):
        for _ in range(2):
            ck.commit_device(s.data_ptr(), n, out.data_ptr(), 0, st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if lg <= 22 else 3
        e0.record(st)
        for _ in range(reps):
            ck.commit_device(s.data_ptr(), n, out.data_ptr(), 0, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = {"ms": round(ms, 3), "mscalar_per_s": round(n / ms / 1e3, 1), "gbps_algorithmic": round(96 * n / ms / 1e6, 2)}
        if check:
            got = out.cpu().numpy().view(np.uint64)
            exp = oracle.msm(curve, s.cpu().numpy().view(np.uint64), d_b.cpu().numpy().view(np.uint64), threads=64)
            res[name]["bit_exact_vs_oracle"] = bool(np.array_equal(got, exp))
    print(json.dumps(res), flush=True)
    ck.close()
    del d_b, U, Wl
    torch.cuda.empty_cache()
