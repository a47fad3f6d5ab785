"""Per-rank view of the N-GPU step on ONE GPU: rank 0's row shard of the k=17 fold_step workload as if `--world` ranks took
part, with the exchange replaced by a local combine of the rank's own partial (same launches, no NCCL).  Used to look at
the per-commitment latency tail that bounds strong scaling without spending 8 GPUs.

  python tools/shard_profile.py --world 8 --steps 20            # CUDA-event breakdown
  ncu --metrics gpu__time_duration.sum ... python tools/shard_profile.py --world 8 --steps 2 --warmup 1   # launch list
"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sirius_b200
from sirius_b200 import _lib
from sirius_b200 import workload as WL

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--k", type=int, default=17)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--exchange", default="peer", choices=["peer", "local"])
ap.add_argument("--no-profile", action="store_true", help="no per-kernel-group events in the timed region (what bench.py's value is timed with)")
args = ap.parse_args()

lib = sirius_b200.load()
_lib.check(lib.sb_init(0))
stream = torch.cuda.Stream()


class LocalCombiner(WL.Combiner):
    def commit(self, ck, d_scalars, n, batch, h_out):
        part, gathered, out = self._buffers(batch)
        ck.commit_batch_device(d_scalars, n, n, batch, 0, part.data_ptr(), self.stream.cuda_stream)
        _lib.check(lib.sb_msm_combine_batch_device(ck.curve, ctypes.c_void_p(part.data_ptr()), 1, batch, batch, ctypes.c_void_p(out.data_ptr()),
                                                   ctypes.c_void_p(self.stream.cuda_stream)))
        with torch.cuda.stream(self.stream):
            h_out.copy_(out.view(h_out.shape), non_blocking=True)
        self.stream.synchronize()


if args.world == 1:
    comb = None
elif args.exchange == "peer":
    comb = WL.PeerCombiner(0, args.world, stream, alone=True)   # the fused exchange kernel, exchanging with itself
else:
    comb = LocalCombiner(args.world, stream)                    # the library path's launches without NCCL
wl = WL.SangriaStepWorkload(args.k, 0, args.world, stream, combiner=comb)
for _ in range(args.warmup):
    wl.step(False)
torch.cuda.synchronize()
TAGS = ["decompose", "sort", "accumulate", "fixup", "reduce", "finalize", "cross_terms", "fold", "ntt", "protogalaxy"]
NT = len(TAGS)
ms_arr, un_arr, ln_arr = (ctypes.c_double * NT)(), (ctypes.c_uint64 * NT)(), (ctypes.c_uint64 * NT)()
if not args.no_profile:
    lib.sb_profile_enable(1)
    lib.sb_profile_collect(ms_arr, un_arr, ln_arr)
wl.host_enqueue_s = 0.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = lib.sb_launch_count()
e0.record(stream)
for _ in range(args.steps):
    wl.step(False)
e1.record(stream)
torch.cuda.synchronize()
if not args.no_profile:
    lib.sb_profile_collect(ms_arr, un_arr, ln_arr)
    lib.sb_profile_enable(0)
ms = e0.elapsed_time(e1) / args.steps
bd = {t: round(ms_arr[i] / args.steps, 4) for i, t in enumerate(TAGS) if ln_arr[i]}
print(json.dumps({"world": args.world, "k": args.k, "ms_per_step_rank0": round(ms, 4), "host_enqueue_ms_per_step": round(wl.host_enqueue_s * 1e3 / args.steps, 4), "sum_kernel_groups": round(sum(bd.values()), 4), "breakdown": bd,
                  "launches_per_step": (lib.sb_launch_count() - l0) // args.steps}))
