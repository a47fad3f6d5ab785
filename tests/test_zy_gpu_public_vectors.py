"""GPU parity on PUBLIC known answers: the EIP-196 bn256Add / bn256ScalarMul precompile vectors of
tests/golden/bn256_g1_kat.json evaluated as commitments by the CUDA path (the same vectors pin the CPU oracle in
tests/test_oracle.py).  Added after the last GPU run of round 1 (its body was dry-run against an oracle-backed stub of
CommitmentKey), hence the file sorts late: it cannot mask another test under `pytest -x`."""
import numpy as np  # noqa: F401
import pytest

from oracle import pyref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import sirius_b200

    sirius_b200.load()
    return sirius_b200


def test_msm_public_precompile_vectors(sb):
    """the CUDA commit on the public EIP-196 bn256Add / bn256ScalarMul known answers (tests/golden/bn256_g1_kat.json):
    commit([k], [P]) = k*P, commit([1, 1], [P, Q]) = P + Q, and the 5-term commitment of all scalar-mul vectors"""
    import json, os

    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bn256_g1_kat.json")))
    C = R.CURVE_BN256
    h = lambda s: int(s, 16)
    for v in kat["scalar_mul"]:
        P, exp = (h(v["x"]), h(v["y"])), (h(v["ex"]), h(v["ey"]))
        ck = sb.CommitmentKey(C, R.points_to_limbs([P], C))
        assert R.limbs_to_points(ck.commit(R.to_mont_limbs([h(v["k"]) % R.FR], R.FR)), C) == [exp], v["name"]
        ck.close()
    ones = R.to_mont_limbs([1, 1], R.FR)
    for v in kat["add"]:
        P, Q, exp = (h(v["x1"]), h(v["y1"])), (h(v["x2"]), h(v["y2"])), (h(v["ex"]), h(v["ey"]))
        ck = sb.CommitmentKey(C, R.points_to_limbs([P, Q], C))
        assert R.limbs_to_points(ck.commit(ones), C) == [exp], v["name"]
        ck.close()
    pts = [(h(v["x"]), h(v["y"])) for v in kat["scalar_mul"]]
    ks = [h(v["k"]) % R.FR for v in kat["scalar_mul"]]
    exp = None
    for v in kat["scalar_mul"]:
        exp = R.ec_add(exp, (h(v["ex"]), h(v["ey"])), C)
    ck = sb.CommitmentKey(C, R.points_to_limbs(pts, C))
    assert R.limbs_to_points(ck.commit(R.to_mont_limbs(ks, R.FR)), C) == [exp]
    ck.close()
