"""Latency / throughput of the device arithmetic primitives (sb_microbench)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sirius_b200 import _lib
lib = _lib.load()
_lib.check(lib.sb_init(0))
names = {0: "mul chain (2 dependent muls/iter)", 1: "4 independent sqr chains/thread", 2: "outlined mul chain", 3: "xyzz_add_call serial", 4: "xyzz_add inline serial", 5: "xyzz_madd serial", 6: "IMAD.WIDE.U32 x8 accumulators"}
per_iter = {0: 2, 1: 4, 2: 2, 3: 1, 4: 1, 5: 1, 6: 8}
clk = 1.965e9
def run(which, iters, blocks, threads):
    ms = ctypes.c_double()
    _lib.check(lib.sb_microbench(which, iters, blocks, threads, ctypes.byref(ms)))
    ops = per_iter[which] * iters
    total = ops * blocks * threads
    print(f"{names[which]:38s} grid={blocks:5d}x{threads:4d} iters={iters:6d}: {ms.value:9.3f} ms  latency/op={ms.value*1e-3/ops*clk:9.1f} cyc  throughput={total/ms.value/1e6:10.2f} Gop/s", flush=True)
for w in (0, 2, 3, 4, 5):
    run(w, 2000, 1, 32)        # one warp alone: latency
for w in (0, 1, 5):
    for thr, blk in ((128, 148 * 4), (128, 148 * 8), (256, 148 * 4)):
        run(w, 1000, blk, thr)  # full chip: throughput
run(6, 100000, 1, 32)
for thr, blk in ((128, 148 * 4), (256, 148 * 4), (256, 148 * 8)):
    run(6, 20000, blk, thr)
