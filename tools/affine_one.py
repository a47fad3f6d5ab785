"""One primary-W-sized commit (1 572 864 scalars, bn256) with a given number of affine rounds: the target of the ncu
captures of k_pair_round / k_accumulate (profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sirius_b200 import _lib, device

lib = _lib.load()
_lib.check(lib.sb_init(0))
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
stream = torch.cuda.Stream()
n = 12 << 17
ck = device.synthetic_key(0, n, window_bits=17, stream=stream)
s = device.random_field_device(n, 0x5349)
out = torch.zeros(8, dtype=torch.int64, device="cuda")
_lib.check(lib.sb_msm_tune(0, rounds))
_lib.check(lib.sb_msm_tune(1, B))
for _ in range(reps):
    ck.commit_batch_device(s.data_ptr(), n, n, 1, out.data_ptr(), 0, stream.cuda_stream)
torch.cuda.synchronize()
print("done", out.cpu()[:2].tolist())
