"""Raw pipe rates on the device (sb_microbench kinds 20..29): cycles per warp-instruction per scheduler for the
instruction kinds the 254-bit product can be built from, alone and mixed.  Decides whether an FP64-pipe product
(52-bit limbs, DFMA) beside the IMAD.WIDE one can pay."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sirius_b200 import _lib
lib = _lib.load()
_lib.check(lib.sb_init(0))
CLK = 1.965e9
KINDS = {
    20: ("IMAD.WIDE.U32 (8 chains)", 8),
    21: ("IMAD.WIDE.U32.X carry chains (mul_ptx pattern)", 8),
    22: ("IMAD lo (8 chains)", 8),
    23: ("IADD3.X carry chains", 16),
    24: ("DFMA.RZ (8 chains)", 8),
    25: ("FP64 limb-product recipe: 8 DFMA + 4 DADD + 16 IADD3", 28),
    26: ("IMAD.WIDE + DFMA interleaved in one warp", 16),
    27: ("even warps IMAD.WIDE / odd warps DFMA", 8),
    28: ("even warps IMAD.WIDE.X chains / odd warps FP64 recipe", 0),
    29: ("64-bit integer add (IADD3 + IADD3.X)", 16),
}
def run(which, iters, blocks, threads):
    ms = ctypes.c_double()
    _lib.check(lib.sb_microbench(which, iters, blocks, threads, ctypes.byref(ms)))
    warps_per_sched = blocks * threads / 32 / (148 * 4)
    cyc_iter = ms.value * 1e-3 * CLK / iters / warps_per_sched   # scheduler cycles per warp-iteration
    name, n = KINDS[which]
    per = f"{cyc_iter / n:6.2f} cyc/instr" if n else ""
    print(f"{name:58s} grid={blocks:5d}x{threads:4d}: {ms.value:8.3f} ms  {cyc_iter:8.2f} cyc/warp-iter/sched  {per}", flush=True)
for which in sorted(KINDS):
    for blocks, threads in ((592, 128), (592, 256), (1184, 256)):
        run(which, 20000, blocks, threads)
