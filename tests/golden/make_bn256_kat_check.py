"""Self-consistency check of tests/golden/bn256_g1_kat.json with an independent affine big-int implementation
(oracle/pyref.py): every operand is on y^2 = x^3 + 3 over the bn256 base field and every expected point is the
group-law result.  The vectors themselves are the public EIP-196 precompile known answers; this script only guards
the transcription.  Run: python tests/golden/make_bn256_kat_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyref as R

kat = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bn256_g1_kat.json")))
C = R.CURVE_BN256
h = lambda s: int(s, 16)
for v in kat["scalar_mul"]:
    P = (h(v["x"]), h(v["y"]))
    assert R.is_on_curve(P, C), v["name"]
    assert R.ec_mul(h(v["k"]), P, C) == (h(v["ex"]), h(v["ey"])), v["name"]
for v in kat["add"]:
    P, Q = (h(v["x1"]), h(v["y1"])), (h(v["x2"]), h(v["y2"]))
    assert R.is_on_curve(P, C) and R.is_on_curve(Q, C), v["name"]
    assert R.ec_add(P, Q, C) == (h(v["ex"]), h(v["ey"])), v["name"]
print(f"{len(kat['scalar_mul'])} scalar-mul and {len(kat['add'])} add vectors are self-consistent")
