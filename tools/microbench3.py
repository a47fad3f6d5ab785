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
run(3, 500, 1, 32, "xyzz_add_call (1 warp)")
run(8, 500, 1, 32, "quad_add (1 warp = 8 groups)")
run(8, 500, 1, 4, "quad_add (1 group)")
run(10, 500, 1, 32, "xyzz_double_call (1 warp)")
run(9, 500, 1, 32, "quad_double (1 warp)")
run(11, 50, 1, 32, "inv_binary (1 warp)")
run(11, 50, 1, 1, "inv_binary (1 thread)")
run(3, 200, 592, 128, "xyzz_add_call (full chip)")
run(8, 200, 592, 128, "quad_add (full chip, 1/4 the adds)")
