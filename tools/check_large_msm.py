"""Size-independent checks of the commitment pipeline at a large size (default 2^23 bn256 scalars, window c = 19,
262 144 buckets = 1024 level-1 partitions): the two sort paths (per-entry atomic counting sort, two-level partition
sort) must agree bit for bit, and commit(a) + commit(b) == commit(a + b).  No CPU MSM needed."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from oracle import pyref as R
from sirius_b200 import _lib, device

lib = _lib.load()
_lib.check(lib.sb_init(0))
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 23
n = 1 << lg
stream = torch.cuda.Stream()
st = stream.cuda_stream
t0 = time.time()
ck = device.synthetic_key(R.CURVE_BN256, n, stream=stream)
print(f"key 2^{lg}: window bits {ck.window_bits}, registered in {time.time() - t0:.1f}s", flush=True)
a = device.random_field_device(n, 1)
b = device.random_field_device(n, 2)
ab = torch.zeros_like(a)
one = R.to_mont_limbs([1], R.FR).reshape(4)
_lib.check(lib.sb_axpy_fold_device(R.FIELD_FR, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), one.ctypes.data_as(_lib.u64p),
                                   ctypes.c_void_p(ab.data_ptr()), n, ctypes.c_void_p(st)))
res = {}
for mode, name in ((2, "partition sort"), (1, "atomic counting sort")):
    _lib.check(lib.sb_msm_tune(2, mode))
    outs = []
    for v in (a, b, ab):
        o = torch.zeros(8, dtype=torch.int64, device="cuda")
        ck.commit_device(v.data_ptr(), n, o.data_ptr(), 0, st)
        stream.synchronize()
        outs.append(o.cpu().numpy().view(np.uint64).copy())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o = torch.zeros(8, dtype=torch.int64, device="cuda")
    e0.record(stream)
    for _ in range(3):
        ck.commit_device(a.data_ptr(), n, o.data_ptr(), 0, st)
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 3:.2f} ms per commit", flush=True)
    res[mode] = outs
_lib.check(lib.sb_msm_tune(2, 0))
ok = all(np.array_equal(x, y) for x, y in zip(res[1], res[2]))
ca, cb, cab = res[2]
lin = np.array_equal(oracle.point_add(R.CURVE_BN256, ca, cb), cab) and oracle.is_on_curve(R.CURVE_BN256, cab) and cab.any()
print("sort paths agree:", ok, " linearity commit(a)+commit(b)==commit(a+b):", bool(lin), flush=True)
sys.exit(0 if ok and lin else 1)
