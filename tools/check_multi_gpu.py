"""torchrun check of the N>1 commit path: row-sharded MSM on each rank + NCCL all-gather of XYZZ partials +
combine kernel == the CPU oracle's full commitment (bit-exact).  Run: torchrun --nproc-per-node N tools/check_multi_gpu.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
import oracle
import sirius_b200
from oracle import pyref as R
from sirius_b200 import _lib, sharding

lib = sirius_b200.load()
_lib.check(lib.sb_init(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for curve in (R.CURVE_BN256, R.CURVE_GRUMPKIN):
    ncols, n = 3, 1 << 12
    sf = 0 if curve == R.CURVE_BN256 else 1
    W = oracle.random_field(sf, 5 + curve, ncols * n)
    bases = oracle.running_bases(curve, ncols * n)
    W_loc = sharding.shard_column_major(W, ncols, n, rank, world)
    ck_loc = sharding.shard_column_major(bases, ncols, n, rank, world)
    ck = sirius_b200.CommitmentKey(curve, ck_loc)
    d_s = torch.from_numpy(W_loc.view(np.int64)).cuda()
    part = torch.zeros(16, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    ck.commit_device(d_s.data_ptr(), W_loc.shape[0], 0, part.data_ptr(), st.cuda_stream)
    gathered = torch.zeros((world, 16), dtype=torch.int64, device="cuda")
    with torch.cuda.stream(st):
        dist.all_gather_into_tensor(gathered, part)
    out = torch.zeros(8, dtype=torch.int64, device="cuda")
    _lib.check(lib.sb_msm_combine_device(curve, ctypes.c_void_p(gathered.data_ptr()), world, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st.cuda_stream)))
    st.synchronize()
    got = out.cpu().numpy().view(np.uint64)
    exp = oracle.msm(curve, W, bases)
    good = bool(np.array_equal(got, exp))
    ok &= good
    if rank == 0:
        print(f"curve {curve}: sharded commit over {world} ranks {'==' if good else '!='} oracle", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
