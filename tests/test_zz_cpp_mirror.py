"""include/sirius_b200.hpp -- the C++ host-side mirror of the Rust interface (CommitmentKey, fft::*, RelaxedPlonkWitness::
fold) -- compiled with g++ and run as a caller of the C ABI would:

  * CPU, real library: host-only behaviour (key file round trip = reference commitment::file_tests::consistency,
    UnexpectedEof, TooLongInput with the reference's message, get_omega_or_inv / get_ifft_divisor against the Python
    mirror) and a loud device error when it tries to compute (no CPU fallback);
  * CPU, the library swapped for an oracle-backed stub of the few entry points (tests/host/fake_sirius_b200.c): the
    mirror's full flow, results equal to the oracle;
  * GPU (`-m gpu`): the same binary against the real library, results equal to the oracle.
(The file sorts last on purpose: it is the newest consumer of the ABI.)"""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyref as R

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
N_PTS, LOG_FFT = 16, 6


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cppmirror") / "cpp_mirror_check")
    libdir = os.path.join(ROOT, "sirius_b200")
    assert os.path.exists(os.path.join(libdir, "libsirius_b200.so")), "build the CUDA library first (__graft_entry__.build())"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-o", out, os.path.join(HERE, "host", "cpp_mirror_check.cpp"),
                           "-L" + libdir, "-lsirius_b200", "-Wl,-rpath," + libdir])
    return out


@pytest.fixture(scope="module")
def inputs(oracle, tmp_path_factory):
    d = tmp_path_factory.mktemp("cppmirror_in")
    pts = oracle.running_bases(R.CURVE_BN256, N_PTS)
    sc = oracle.random_field(R.FIELD_FR, 21, N_PTS)
    poly = oracle.random_field(R.FIELD_FR, 22, 1 << LOG_FFT)
    r = oracle.random_field(R.FIELD_FR, 23, 1)
    path = str(d / "inputs.bin")
    with open(path, "wb") as f:
        f.write(np.array([N_PTS, LOG_FFT], dtype=np.uint64).tobytes())
        for a in (pts, sc, poly, r):
            f.write(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
    return dict(path=path, dir=str(d), pts=pts, sc=sc, poly=poly, r=r)


def run(binary, inputs, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([binary, inputs["dir"], inputs["path"]], capture_output=True, text=True, timeout=300, env=e)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    out = {}
    for line in r.stdout.splitlines():
        parts = line.split()
        if not parts:
            continue
        key = parts[0] if parts[0] != "omega" else ("omega", int(parts[1]))
        rest = parts[1:] if parts[0] != "omega" else parts[2:]
        out[key] = rest
    return out, r.stdout


def words(tokens):
    return np.array([int(t, 16) for t in tokens], dtype=np.uint64)


def check_host_part(out):
    from sirius_b200 import fft as pyfft

    assert "host" in out and out["host"] == ["ok"]
    assert " ".join(out["too_long"]) == f"Can't commit too long input: input len: {N_PTS + 1}, but limit is {N_PTS}"
    for k in range(pyfft.FR_S + 1):
        w = words(out[("omega", k)])
        assert np.array_equal(w[0:4], pyfft.fr_to_limbs(pyfft.get_omega_or_inv(k, False))), k
        assert np.array_equal(w[4:8], pyfft.fr_to_limbs(pyfft.get_omega_or_inv(k, True))), k
        assert np.array_equal(w[8:12], pyfft.fr_to_limbs(pyfft.get_ifft_divisor(k))), k


def check_device_part(out, oracle, inputs):
    import ctypes

    from sirius_b200 import fft as pyfft

    assert out.get("device") == ["ok"], out.get("device_error")
    pts, sc, poly, r = inputs["pts"], inputs["sc"], inputs["poly"], inputs["r"]
    assert np.array_equal(words(out["commit"]), oracle.msm(R.CURVE_BN256, sc, pts))
    assert np.array_equal(words(out["commit_prefix"]), oracle.msm(R.CURVE_BN256, sc[: N_PTS // 2], pts))
    assert np.array_equal(words(out["fft"]).reshape(-1, 4), oracle.fft(poly))
    cos = poly.copy()
    lib, u64p = oracle.lib(), ctypes.POINTER(ctypes.c_uint64)
    z, z2 = pyfft.fr_to_limbs(pyfft.FR_ZETA), pyfft.fr_to_limbs(pyfft.FR_ZETA * pyfft.FR_ZETA % pyfft.FR_MODULUS)
    assert lib.so_coset_scale(R.FIELD_FR, cos.ctypes.data_as(u64p), ctypes.c_size_t(cos.shape[0]), z.ctypes.data_as(u64p), z2.ctypes.data_as(u64p)) == 0
    assert np.array_equal(words(out["coset_fft"]).reshape(-1, 4), oracle.fft(cos))
    # fold: W + r*W2 with W2 = reversed W; E + r*T1 + r^2*T2 with T1 = E, T2 = reversed E
    m = R.FR
    W, E, rr = R.from_mont_limbs(sc, m), R.from_mont_limbs(poly, m), R.from_mont_limbs(r, m)[0]
    expW = [(a + rr * b) % m for a, b in zip(W, W[::-1])]
    expE = [(e + rr * e + rr * rr * e2) % m for e, e2 in zip(E, E[::-1])]
    assert R.from_mont_limbs(words(out["fold_W"]).reshape(-1, 4), m) == expW
    assert R.from_mont_limbs(words(out["fold_E"]).reshape(-1, 4), m) == expE


def test_cpp_mirror_host_logic_and_no_fallback(binary, inputs):
    import torch

    out, text = run(binary, inputs)
    check_host_part(out)
    if not torch.cuda.is_available():
        assert out["device_error"][0] == "-1", text[-500:]   # SB_ERR_CUDA: nothing computes without the device


def test_cpp_mirror_full_flow_against_oracle_stub(binary, inputs, oracle, tmp_path):
    so = tmp_path / "libsirius_b200.so"
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-shared", "-fPIC", "-o", str(so), os.path.join(HERE, "host", "fake_sirius_b200.c"),
                           "-L" + os.path.join(ROOT, "oracle"), "-lsirius_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    out, _ = run(binary, inputs, env={"LD_LIBRARY_PATH": str(tmp_path)})
    check_host_part(out)
    check_device_part(out, oracle, inputs)


@pytest.mark.gpu
def test_cpp_mirror_on_the_device(binary, inputs, oracle):
    out, _ = run(binary, inputs)
    check_host_part(out)
    check_device_part(out, oracle, inputs)


# ---------------------------------------------------------------------------------------------- expression compiler
@pytest.fixture(scope="module")
def expr_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cppexpr") / "cpp_expr_check")
    libdir = os.path.join(ROOT, "sirius_b200")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-o", out, os.path.join(HERE, "host", "cpp_expr_check.cpp"),
                           "-L" + libdir, "-lsirius_b200", "-Wl,-rpath," + libdir])
    return out


def _fmt_program(ev):
    lines = ["rotations" + "".join(f" {r}" for r in ev.rotations), "constants" + "".join(f" {c:064x}" for c in ev.constants)]
    for op, a, b, target in ev.calculations:
        lines.append(f"calc {op} {a[0]} {a[1]} {a[2]} {b[0]} {b[1]} {b[2]} {target}" if b is not None else f"calc {op} {a[0]} {a[1]} {a[2]} - {target}")
    return lines


@pytest.mark.parametrize("T_list", [[2], [5], [5, 3]])
def test_cpp_expression_compiler_matches_python_and_oracle(expr_binary, T_list):
    """include/sirius_b200_expr.hpp (Expression, homogeneous, compress_expression, CompressedGates, GraphEvaluator::new,
    main_gate_expression) emits the calculation lists of the Python mirror -- which tests/test_host_logic.py ties to the
    oracle -- for the MainGate structures of the benches, for both the compressed gate and its homogeneous form"""
    from sirius_b200 import polynomial as P

    r = subprocess.run([expr_binary, ",".join(str(t) for t in T_list)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout.splitlines()
    nfix, nadv = sum(2 * T + 5 for T in T_list), sum(T + 2 for T in T_list)
    gates, fb, ab = [], 0, 0
    for T in T_list:
        gates.append(P.main_gate_expression(T, fb, ab, 0, nfix))
        fb, ab = fb + 2 * T + 5, ab + T + 2
    cg = P.CompressedGates.new(gates, P.QueryIndexContext(num_fixed=nfix, num_advice=nadv))
    assert out[0] == f"degree {cg.degree} num_challenges {cg.ctx.num_challenges}"
    pos = 1
    for which, expr in enumerate((cg.compressed, cg.homogeneous)):
        ev = P.GraphEvaluator.new(expr, P.FR)
        assert out[pos] == f"program {which} calcs {len(ev.calculations)} intermediates {ev.num_intermediates}"
        exp = _fmt_program(ev)
        assert out[pos + 1: pos + 1 + len(exp)] == exp
        pos += 1 + len(exp)
        assert out[pos] == f"const2_mont {(2 << 256) % P.FR:064x}"     # constants cross the ABI in Montgomery form
        pos += 1
    # Negated constant / Scaled / Sub / rotation, over Fq
    e = (P.Expression.Polynomial(3, 1) - P.Expression.Constant(5)) * 7 + (-P.Expression.Constant(9)) * P.Expression.Challenge(0)
    ev = P.GraphEvaluator.new(e, P.FQ)
    assert out[pos] == f"extra calcs {len(ev.calculations)}"
    assert out[pos + 1:] == _fmt_program(ev)[1:]


# ---------------------------------------------------------------------------------------------- commit_cross_terms
CT_T_LIST, CT_K = [2, 2], 4      # two MainGate<2> gates (a compression challenge + the homogenising one), 16 rows


@pytest.fixture(scope="module")
def ct_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cppct") / "cpp_cross_terms_check")
    libdir = os.path.join(ROOT, "sirius_b200")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-o", out, os.path.join(HERE, "host", "cpp_cross_terms_check.cpp"),
                           "-L" + libdir, "-lsirius_b200", "-Wl,-rpath," + libdir])
    return out


@pytest.fixture(scope="module")
def ct_inputs(oracle, tmp_path_factory):
    from oracle import expr_ref as E

    d = tmp_path_factory.mktemp("cppct_in")
    n = 1 << CT_K
    nfix, nadv = sum(2 * T + 5 for T in CT_T_LIST), sum(T + 2 for T in CT_T_LIST)
    gates, fb, ab = [], 0, 0
    for T in CT_T_LIST:
        gates.append(E.main_gate_expression(T, fb, ab, 0, nfix))
        fb, ab = fb + 2 * T + 5, ab + T + 2
    cg = E.CompressedGates(gates, E.Ctx(num_fixed=nfix, num_advice=nadv))
    nch = cg.ctx.num_challenges
    key = oracle.running_bases(R.CURVE_BN256, n)
    fixed = [oracle.random_field(R.FIELD_FR, 300 + j, n) for j in range(nfix)]
    W1, W2 = oracle.random_field(R.FIELD_FR, 31, nadv * n), oracle.random_field(R.FIELD_FR, 32, nadv * n)
    U1c, U1u, U2c = oracle.random_field(R.FIELD_FR, 33, nch - 1), oracle.random_field(R.FIELD_FR, 34, 1), oracle.random_field(R.FIELD_FR, 35, nch - 1)
    path = str(d / "ct_inputs.bin")
    with open(path, "wb") as f:
        f.write(np.array([len(CT_T_LIST)] + CT_T_LIST + [CT_K, n], dtype=np.uint64).tobytes())
        for a in [key] + fixed + [W1, W2, U1c, U1u, U2c]:
            f.write(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
    # expected: the reference's literal path (one GraphEvaluator per degree-grouped expression), Python integers
    m = R.FR
    S = E.Structure(CT_K, [], [R.from_mont_limbs(c, m) for c in fixed], nadv, 0, cg, m)
    exp_T = E.commit_cross_terms_eval(S, R.from_mont_limbs(U1c, m), R.from_mont_limbs(U1u, m)[0], [R.from_mont_limbs(W1, m)],
                                      R.from_mont_limbs(U2c, m), [R.from_mont_limbs(W2, m)])
    return dict(path=path, key=key, exp_T=exp_T, degree=cg.degree, nch=nch)


def _check_cross_terms(out_text, ct_inputs, oracle):
    lines = out_text.splitlines()
    assert lines[0] == f"degree {ct_inputs['degree']} num_challenges {ct_inputs['nch']}"
    assert lines[-1] == "device ok", lines[-1]
    T, C = {}, {}
    for line in lines[1:-1]:
        parts = line.split()
        (T if parts[0] == "cross_term" else C)[int(parts[1])] = words(parts[2:])
    assert len(T) == len(C) == ct_inputs["degree"]
    for j, exp in enumerate(ct_inputs["exp_T"]):
        got = T[j].reshape(-1, 4)
        assert R.from_mont_limbs(got, R.FR) == exp, f"cross term {j + 1}"
        assert np.array_equal(C[j], oracle.msm(R.CURVE_BN256, got, ct_inputs["key"])), f"commitment of cross term {j + 1}"


def test_cpp_commit_cross_terms_against_oracle_stub(ct_binary, ct_inputs, oracle, tmp_path):
    """VanillaFS::commit_cross_terms of the C++ mirror (structure registration, program upload, round vectors and
    challenge vectors in the reference's order, batched commit) executed against the oracle-backed ABI stub and compared
    with the reference's literal degree-grouped evaluation (oracle/expr_ref.py) + the oracle MSM"""
    so = tmp_path / "libsirius_b200.so"
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-shared", "-fPIC", "-o", str(so), os.path.join(HERE, "host", "fake_sirius_b200.c"),
                           "-L" + os.path.join(ROOT, "oracle"), "-lsirius_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    e = dict(os.environ)
    e["LD_LIBRARY_PATH"] = str(tmp_path)
    r = subprocess.run([ct_binary, ct_inputs["path"]], capture_output=True, text=True, timeout=300, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    _check_cross_terms(r.stdout, ct_inputs, oracle)


@pytest.mark.gpu
def test_cpp_commit_cross_terms_on_the_device(ct_binary, ct_inputs, oracle):
    r = subprocess.run([ct_binary, ct_inputs["path"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    _check_cross_terms(r.stdout, ct_inputs, oracle)
