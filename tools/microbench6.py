"""Full-chip throughput of the mixed XYZZ addition, canonical (sb_microbench 5) and lazy-domain (13) forms: the
integer-pipe peak bench.py quotes for k_accumulate."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sirius_b200 import _lib
lib = _lib.load()
_lib.check(lib.sb_init(0))
def run(which, iters, blocks, threads, name):
    ms = ctypes.c_double()
    _lib.check(lib.sb_microbench(which, iters, blocks, threads, ctypes.byref(ms)))
    total = iters * blocks * threads
    print(f"{name:34s} grid={blocks}x{threads}: {ms.value:8.3f} ms  {total / ms.value / 1e6:8.3f} G madd/s", flush=True)
for _ in range(2):
    for blocks, threads in ((592, 128), (1184, 128), (592, 256)):
        run(5, 1000, blocks, threads, "xyzz_madd (canonical)")
        run(13, 1000, blocks, threads, "xyzz_madd_lazy")
