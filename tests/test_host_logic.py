"""CPU tests of the host-side mirror: expression compile parity with the oracle's literal restatement, the
C-ABI library exports, and the header/binding agreement.  No compute calls (no GPU here)."""
import ctypes
import os
import re

from oracle import expr_ref as E
from oracle import pyref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _to_oracle(e):
    k = e.kind
    if k == "const":
        return E.Const(e.a)
    if k == "poly":
        return E.Poly(e.a, e.b)
    if k == "chal":
        return E.Chal(e.a)
    if k == "neg":
        return E.Neg(_to_oracle(e.a))
    if k == "sum":
        return E.Sum(_to_oracle(e.a), _to_oracle(e.b))
    if k == "prod":
        return E.Mul(_to_oracle(e.a), _to_oracle(e.b))
    if k == "scaled":
        return E.Scaled(_to_oracle(e.a), e.b)
    raise ValueError(k)


def test_main_gate_and_homogeneous_match_oracle():
    from sirius_b200 import polynomial as P

    for T_list in ([2], [5], [5, 3]):
        nfix = sum(2 * T + 5 for T in T_list)
        nadv = sum(T + 2 for T in T_list)
        gates_p, gates_o, fb, ab = [], [], 0, 0
        for T in T_list:
            gates_p.append(P.main_gate_expression(T, fb, ab, 0, nfix))
            gates_o.append(E.main_gate_expression(T, fb, ab, 0, nfix))
            fb += 2 * T + 5
            ab += T + 2
        for gp, go in zip(gates_p, gates_o):
            assert _to_oracle(gp) == go
        cp = P.CompressedGates.new(gates_p, P.QueryIndexContext(num_fixed=nfix, num_advice=nadv))
        co = E.CompressedGates(gates_o, E.Ctx(num_fixed=nfix, num_advice=nadv))
        assert _to_oracle(cp.homogeneous) == co.homogeneous
        assert cp.degree == co.degree == (5 if len(T_list) == 1 else 6)
        assert cp.ctx.num_challenges == co.ctx.num_challenges
        # compiled calculation lists are identical too (same CSE, same operand ordering)
        gp = P.GraphEvaluator.new(cp.homogeneous, R.FR)
        go = E.GraphEvaluator(co.homogeneous, R.FR)
        assert gp.constants == go.constants and gp.rotations == go.rotations
        assert [(o, a, b, t) for o, a, b, t in gp.calculations] == go.calcs


def test_library_exports_every_declared_symbol():
    """include/sirius_b200.h <-> libsirius_b200.so <-> the ctypes table."""
    from sirius_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "sirius_b200.h")).read()
    declared = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback_without_gpu():
    import numpy as np
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sirius_b200

    with pytest.raises(sirius_b200.SiriusB200Error):
        sirius_b200.CommitmentKey(0, np.zeros((4, 8), dtype=np.uint64))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sirius_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "sirius_oracle" not in src, f


def test_header_is_plain_c():
    """the boundary is a C ABI: the header must compile as C with no torch/CUDA types in the signatures"""
    import subprocess, tempfile

    hdr = os.path.join(ROOT, "include", "sirius_b200.h")
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write(f'#include "{hdr}"\nint main(void) {{ return sb_version() == 0; }}\n')
        path = f.name
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", path])
    os.unlink(path)
    import re

    src = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)  # declarations only; comments may name the types
    src = re.sub(r"//[^\n]*", "", src)
    assert "torch" not in src and "cudaStream_t" not in src and "at::" not in src


def test_setup_smallest_key_size_rule():
    """`setup_smallest_key` (src/commitment.rs:172-186): k = max over (advice + 5*lookups, selectors + fixed) of
    ceil(log2(n * 2^K)), with the reference's float arithmetic"""
    from sirius_b200.commitment import smallest_key_log2, smallest_power

    assert smallest_power(12, 17) == 21 and smallest_power(7, 17) == 20 and smallest_power(16, 17) == 21
    assert smallest_power(1, 17) == 17 and smallest_power(0, 17) == 0 and smallest_power(3, 0) == 2
    assert smallest_key_log2(17, 12, 0, 0, 26) == 22      # sangria_poseidon primary: 26 fixed columns dominate
    assert smallest_key_log2(17, 7, 0, 0, 15) == 21       # secondary
    assert smallest_key_log2(10, 3, 1, 2, 1) == 13        # 3 advice + 5 lookup columns = 8 -> 2^13
    for n in range(1, 70):
        for K in (0, 5, 17, 20):
            w = smallest_power(n, K)
            assert (1 << w) >= n << K and (w == 0 or (1 << (w - 1)) < n << K)


def test_gate_scaling_shapes_and_folding_degree():
    """benches/ivc_gate_scaling.rs widens the primary circuit by N Poseidon sub-circuits: A = 7 + 5N, F = 15 + 11N, N + 1 gates of
    degree 5 compressed with powers of one challenge (src/plonk/util.rs:35-55) -> folding degree 5 + N."""
    from sirius_b200 import polynomial as P
    from sirius_b200 import workload as WL

    for N in (1, 5, 10, 20):
        side = WL.gate_scaling_side(N)
        gates, nfix, nadv = WL.compressed_gates(side)
        assert (len(gates), nfix, nadv) == (N + 1, 15 + 11 * N, 7 + 5 * N)
        cg = P.CompressedGates.new(gates, P.QueryIndexContext(num_selectors=0, num_fixed=nfix, num_advice=nadv))
        assert cg.degree == 5 + N and cg.ctx.num_challenges == 2
    assert WL.gate_scaling_side(1)["T_list"] == WL.PRIMARY["T_list"]


def test_integration_doc_names_only_declared_entry_points():
    """Every `sb_*` function INTEGRATION.md binds or calls is declared in include/sirius_b200.h (the Rust shim a maintainer copies
    from the document must link)."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "sirius_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", header))
    used = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", doc))
    assert used and not (used - declared), sorted(used - declared)
