"""torchrun check of the N>1 path bench.py times: the row-sharded SangriaStepWorkload (per-rank MSM + exchange of XYZZ
partials + combine kernel, row-local cross terms and folds) against the CPU oracle's full step, bit for bit.
Run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_multi_gpu.py [--k 17]
(launched by tests/test_gpu_workload.py::test_multi_gpu_step_torchrun when the box has >= 2 GPUs)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=12)
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
import oracle
import sirius_b200
from oracle import step_ref
from sirius_b200 import _lib
from sirius_b200 import workload as WL

lib = sirius_b200.load()
_lib.check(lib.sb_init(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
oracle.build()
wl = WL.SangriaStepWorkload(args.k, rank, world, torch.cuda.Stream())
ok = True
for i in range(args.steps):
    snap = wl.snapshot_inputs()
    wl.step(upload=(i % 2 == 0))
    got = wl.snapshot_results()
    if rank == 0:
        exp = step_ref.fold_step(snap, step_ref.bases_for(snap))
        rep = step_ref.compare(got, exp)
        ok &= rep["ok"]
        print(f"step {i}: fold_step over {world} ranks {'==' if rep['ok'] else '!='} oracle {rep['bad']}", flush=True)
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
