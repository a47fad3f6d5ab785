"""Latency / throughput of the device arithmetic primitives (sb_microbench), one driver for every suite:

  python tools/microbench.py [suite ...]      suites: field, karatsuba, group, inversion, pipes, madd, coop   (default: all)

  field      Montgomery product: one warp alone (latency) and the full chip (throughput); raw IMAD.WIDE rate
  karatsuba  the 128-IMAD.WIDE product against a 112-IMAD.WIDE + ~190-ALU stand-in (prices a Karatsuba product)
  group      XYZZ addition / doubling: single lane, 4-lane groups (quad.cuh)
  inversion  binary extended Euclid against the divstep (safegcd) inversion
  pipes      cycles per warp-instruction per scheduler for IMAD / IMAD.WIDE / IADD3 / DFMA, alone and mixed
  madd       full-chip throughput of the bucket kernel's mixed addition, canonical and lazy-domain
  coop       addition / doubling by the 4 warps of a block (coop.cuh): latency alone and with co-resident blocks
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sirius_b200 import _lib

lib = _lib.load()
_lib.check(lib.sb_init(0))
CLK = 1.965e9


def ms_of(which, iters, blocks, threads):
    ms = ctypes.c_double()
    _lib.check(lib.sb_microbench(which, iters, blocks, threads, ctypes.byref(ms)))
    return ms.value


def lat(which, iters, blocks, threads, name, ops_per_iter=1):
    ms = ms_of(which, iters, blocks, threads)
    print(f"{name:44s} grid={blocks:5d}x{threads:4d} iters={iters:6d}: {ms:9.3f} ms  {ms * 1e-3 / (iters * ops_per_iter) * CLK:9.0f} cycles/op", flush=True)


def thr(which, iters, blocks, threads, name, ops_per_iter=1, unit="Gop/s", lanes_per_thread=1.0):
    ms = ms_of(which, iters, blocks, threads)
    total = ops_per_iter * iters * blocks * threads * lanes_per_thread
    print(f"{name:44s} grid={blocks:5d}x{threads:4d} iters={iters:6d}: {ms:9.3f} ms  {total / ms / 1e6:10.3f} {unit}", flush=True)


def suite_field():
    for w, name, per in ((0, "mul chain (2 dependent muls/iter)", 2), (2, "outlined mul chain", 2)):
        lat(w, 2000, 1, 32, name + " (1 warp)", per)
    for w, name, per in ((0, "mul chain", 2), (1, "4 independent sqr chains/thread", 4)):
        for t, b in ((128, 592), (128, 1184), (256, 592)):
            thr(w, 1000, b, t, name, per, "Gmul/s")
    lat(6, 100000, 1, 32, "IMAD.WIDE.U32 x8 accumulators (1 warp)", 8)
    for t, b in ((128, 592), (256, 592), (256, 1184)):
        thr(6, 20000, b, t, "IMAD.WIDE.U32 x8 accumulators", 8, "G imad.wide/s")


def suite_karatsuba():
    for _ in range(2):
        thr(0, 1000, 592, 128, "mul (128 IMAD.WIDE)", 2, "Gmul/s")
        thr(7, 1000, 592, 128, "fake (112 IMAD.WIDE + ~190 ALU)", 2, "Gmul/s")


def suite_group():
    lat(3, 500, 1, 32, "xyzz_add_call (1 warp)")
    lat(4, 500, 1, 32, "xyzz_add inline (1 warp)")
    lat(5, 500, 1, 32, "xyzz_madd (1 warp)")
    lat(8, 500, 1, 32, "quad_add (1 warp = 8 groups)")
    lat(8, 500, 1, 4, "quad_add (1 group)")
    lat(10, 500, 1, 32, "xyzz_double_call (1 warp)")
    lat(9, 500, 1, 32, "quad_double (1 warp)")
    thr(3, 200, 592, 128, "xyzz_add_call (full chip)", 1, "G add/s")
    thr(8, 200, 592, 128, "quad_add (full chip)", 1, "G add/s", 0.25)


def suite_inversion():
    for which, name in ((11, "inv_binary"), (12, "inv_safegcd")):
        lat(which, 50, 1, 1, f"{name} (1 thread)")
        lat(which, 50, 1, 32, f"{name} (1 warp, divergent)")
        thr(which, 20, 592, 128, f"{name} (full chip, 4 warps/SMSP)", 1, "G inv/s")


def suite_pipes():
    kinds = {
        20: ("IMAD.WIDE.U32 (8 chains)", 8), 21: ("IMAD.WIDE.U32.X carry chains (mul_ptx pattern)", 8), 22: ("IMAD lo (8 chains)", 8),
        23: ("IADD3.X carry chains", 16), 24: ("DFMA.RZ (8 chains)", 8), 25: ("FP64 limb-product recipe: 8 DFMA + 4 DADD + 16 IADD3", 28),
        26: ("IMAD.WIDE + DFMA interleaved in one warp", 16), 27: ("even warps IMAD.WIDE / odd warps DFMA", 8),
        28: ("even warps IMAD.WIDE.X chains / odd warps FP64 recipe", 0), 29: ("64-bit integer add (IADD3 + IADD3.X)", 16),
    }
    for which in sorted(kinds):
        for blocks, threads in ((592, 128), (592, 256), (1184, 256)):
            ms = ms_of(which, 20000, blocks, threads)
            warps_per_sched = blocks * threads / 32 / (148 * 4)
            cyc_iter = ms * 1e-3 * CLK / 20000 / warps_per_sched
            name, n = kinds[which]
            per = f"{cyc_iter / n:6.2f} cyc/instr" if n else ""
            print(f"{name:58s} grid={blocks:5d}x{threads:4d}: {ms:8.3f} ms  {cyc_iter:8.2f} cyc/warp-iter/sched  {per}", flush=True)


def suite_madd():
    for _ in range(2):
        for blocks, threads in ((592, 128), (1184, 128), (592, 256)):
            thr(5, 1000, blocks, threads, "xyzz_madd (canonical)", 1, "G madd/s")
            thr(13, 1000, blocks, threads, "xyzz_madd_lazy", 1, "G madd/s")


def suite_coop():
    lat(3, 500, 1, 32, "xyzz_add_call (1 warp, the form replaced)")
    lat(14, 500, 1, 128, "coop4_add (1 block = 4 warps, 32 additions)")
    lat(14, 500, 148, 128, "coop4_add (1 block per SM)")
    lat(14, 500, 296, 128, "coop4_add (2 blocks per SM)")
    lat(14, 500, 592, 128, "coop4_add (4 blocks per SM)")
    lat(14, 200, 1184, 128, "coop4_add (8 blocks per SM)")
    lat(10, 500, 1, 32, "xyzz_double_call (1 warp, the form replaced)")
    lat(15, 500, 1, 128, "coop4_double (1 block)")
    lat(15, 500, 592, 128, "coop4_double (4 blocks per SM)")
    thr(14, 200, 1184, 128, "coop4_add (8 blocks per SM)", 1, "G add/s", 0.25)
    thr(3, 200, 592, 128, "xyzz_add_call (full chip)", 1, "G add/s")


SUITES = {"field": suite_field, "karatsuba": suite_karatsuba, "group": suite_group, "inversion": suite_inversion, "pipes": suite_pipes,
          "madd": suite_madd, "coop": suite_coop}
if __name__ == "__main__":
    for name in (sys.argv[1:] or list(SUITES)):
        print(f"== {name}", flush=True)
        SUITES[name]()
