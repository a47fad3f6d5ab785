"""The run-time compiled cross-term kernels (csrc/expr.cu): the code generator and NVRTC need no device (the compiler
cross-compiles for sm_100a), so the CPU suite checks that every program shape the hot path uses -- the bench's MainGate
sides, the Cyclefold support gate with its selector, a gate with rotations and a challenge -- lowers, generates and
compiles, and reports the resource usage.  Bit-exactness of what the kernels compute is the GPU suite's job
(tests/test_gpu_sangria.py runs through them by default; tests/test_gpu_jit.py compares them with the interpreter)."""
import ctypes

import numpy as np
import pytest


def _calcs(ev):
    from sirius_b200 import _lib
    from sirius_b200 import polynomial as P

    calcs = (_lib.sb_calc * max(1, len(ev.calculations)))()
    for i, (op, a, b, target) in enumerate(ev.calculations):
        c = calcs[i]
        c.opcode, c.a_kind, c.a_index, c.a_rot = op, a[0], a[1], a[2]
        if b is not None and op <= P.OP_MUL:
            c.b_kind, c.b_index, c.b_rot = b
        c.target = target
    return calcs


def _compile(field, ev, degree, nsel, nfix, nfv, nch):
    from sirius_b200 import _lib

    lib = _lib.load()
    calcs = _calcs(ev)
    rots = np.array(ev.rotations if ev.rotations else [0], dtype=np.int32)
    log = ctypes.create_string_buffer(1 << 16)
    nb = ctypes.c_size_t()
    rc = lib.sb_expr_jit_selftest(field, ctypes.cast(calcs, ctypes.c_void_p), len(ev.calculations), len(ev.constants),
                                  rots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(ev.rotations), degree, nsel, nfix, nfv, nch, log, 1 << 16, ctypes.byref(nb))
    return rc, nb.value, log.value.decode(errors="replace")


@pytest.mark.parametrize("side_name", ["PRIMARY", "SECONDARY", "SUPPORT"])
def test_bench_programs_compile_to_straight_line_kernels(side_name):
    from sirius_b200 import curves
    from sirius_b200 import polynomial as P
    from sirius_b200 import workload as WL

    side = getattr(WL, side_name)
    gates, nfix, nadv = WL.compressed_gates(side)
    nsel = WL.num_selectors(side)
    cg = P.CompressedGates.new(gates, P.QueryIndexContext(num_selectors=nsel, num_fixed=nfix, num_advice=nadv))
    ev = P.GraphEvaluator.new(cg.homogeneous, curves.SCALAR_FIELD[side["curve"]])
    rc, nbytes, log = _compile(side["field"], ev, cg.degree, nsel, nfix, nadv, cg.ctx.num_challenges)
    assert rc == 0, log
    assert nbytes > 10000
    assert "0 bytes spill stores" in log and "Used" in log, log


def test_rotations_challenges_and_second_instance_operands_compile():
    from sirius_b200 import fft
    from sirius_b200 import polynomial as P

    E = P.Expression
    # selector * (fixed(rot -1) * advice(rot +1) * challenge - advice2(rot 0)) : 1 selector, 1 fixed, 2 advice per instance
    expr = E.Polynomial(0, 0) * (E.Polynomial(1, -1) * E.Polynomial(2, 1) * E.Challenge(0) - E.Polynomial(3 + 2, 0))
    ev = P.GraphEvaluator.new(expr, fft.FR_MODULUS)
    rc, nbytes, log = _compile(0, ev, 3, 1, 1, 2, 1)
    assert rc == 0, log
    assert nbytes > 1000


def test_bad_program_is_rejected_before_code_generation():
    from sirius_b200 import _lib

    lib = _lib.load()
    calcs = (_lib.sb_calc * 1)()
    calcs[0].opcode, calcs[0].a_kind, calcs[0].a_index, calcs[0].target = 6, 0, 0, 0   # Horner: never emitted by the reference
    rots = np.zeros(1, dtype=np.int32)
    rc = lib.sb_expr_jit_selftest(0, ctypes.cast(calcs, ctypes.c_void_p), 1, 3, rots.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 1, 2, 0, 1, 1, 0, None, 0, None)
    assert rc == _lib.SB_ERR_ARG
    assert b"Horner" in lib.sb_last_error()


@pytest.mark.parametrize("num_blend", [0, 2, 4])
def test_plain_and_blended_evaluation_kernels_compile(num_blend):
    """degree 0: the per-row evaluation kernel of sb_expr_eval (deciders); num_blend traces: the Protogalaxy leaf kernel that
    evaluates a gate on the Lagrange blend of the traces (each leaf blended once)."""
    from sirius_b200 import fft
    from sirius_b200 import polynomial as P

    gate = P.main_gate_expression(5, 0, 0, 0, 15)
    ev = P.GraphEvaluator.new(gate, fft.FR_MODULUS)
    rc, nbytes, log = _compile(0, ev, num_blend << 8, 0, 15, 7, 0)
    assert rc == 0, log
    assert nbytes > 1000 and "0 bytes spill stores" in log


def test_long_programs_compile_in_the_call_form():
    """Above SB_EXPR_JIT_INLINE_MAX_OPS calculations (the gate-scaling circuits) the generated kernel calls the product instead of
    inlining it: N = 10 sub-circuits (619 calculations, folding degree 15) must compile in seconds, without a stack frame."""
    import time

    from sirius_b200 import curves
    from sirius_b200 import polynomial as P
    from sirius_b200 import workload as WL

    side = WL.gate_scaling_side(10)
    gates, nfix, nadv = WL.compressed_gates(side)
    cg = P.CompressedGates.new(gates, P.QueryIndexContext(num_selectors=0, num_fixed=nfix, num_advice=nadv))
    ev = P.GraphEvaluator.new(cg.homogeneous, curves.SCALAR_FIELD[0])
    assert len(ev.calculations) > 400 and cg.degree == 15
    t0 = time.time()
    rc, nbytes, log = _compile(0, ev, cg.degree, 0, nfix, nadv, cg.ctx.num_challenges)
    assert rc == 0, log
    assert time.time() - t0 < 60
    assert "0 bytes spill stores" in log and nbytes < 2_000_000, log[-600:]
