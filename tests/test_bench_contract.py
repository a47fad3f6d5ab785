"""bench.py's output contract, checked on the CPU: the reference arm (`--impl reference`, the oracle timed on the host
cores) prints exactly one JSON line with the keys the driver reads; the GPU arm refuses to run without a device
(no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] == d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("benches/sangria_poseidon k=17")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["data"] == "synthetic"


def test_gpu_arm_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not any(l.strip().startswith("{") for l in r.stdout.splitlines())   # no number is reported


def test_reference_arm_of_the_other_workloads():
    """`--workload cyclefold_poseidon | gate_scaling --impl reference`: the CPU restatement of those hot paths, one contract line
    each (small tables so that the CPU suite stays short)."""
    for extra, needle in ((["--workload", "cyclefold_poseidon", "--k", "8"], "benches/cyclefold_poseidon k=8"),
                          (["--workload", "gate_scaling", "--k", "7", "--gates", "2"], "benches/ivc_gate_scaling")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra,
                           capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
        assert len(lines) == 1, (extra, r.stdout[-500:])
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["unit"] == "ms" and d["value"] > 0 and needle in d["config"]["workload"], d
        assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
