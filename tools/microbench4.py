"""inv_binary vs inv_safegcd (field.cuh) latency and throughput.  Not the bench."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sirius_b200 import _lib
lib = _lib.load()
_lib.check(lib.sb_init(0))
clk = 1.965e9
def run(which, iters, blocks, threads, name):
    ms = ctypes.c_double()
    _lib.check(lib.sb_microbench(which, iters, blocks, threads, ctypes.byref(ms)))
    print(f"{name:36s} grid={blocks}x{threads} iters={iters}: {ms.value:8.3f} ms  {ms.value*1e-3/iters*clk:9.0f} cycles/op", flush=True)
for which, name in ((11, "inv_binary"), (12, "inv_safegcd")):
    run(which, 50, 1, 1, f"{name} (1 thread)")
    run(which, 50, 1, 32, f"{name} (1 warp, divergent)")
    run(which, 20, 592, 128, f"{name} (full chip, 4 warps/SMSP)")
