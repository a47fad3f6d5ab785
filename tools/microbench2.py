import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sirius_b200 import _lib
lib = _lib.load()
_lib.check(lib.sb_init(0))
def run(which, iters, blocks, threads, name):
    ms = ctypes.c_double()
    _lib.check(lib.sb_microbench(which, iters, blocks, threads, ctypes.byref(ms)))
    total = 2 * iters * blocks * threads
    print(f"{name:30s} grid={blocks}x{threads}: {ms.value:8.3f} ms  {total/ms.value/1e6:8.2f} Gmul/s", flush=True)
for _ in range(2):
    run(0, 1000, 592, 128, "mul (128 IMAD.WIDE)")
    run(7, 1000, 592, 128, "fake (112 IMAD.WIDE + ~190 ALU)")
