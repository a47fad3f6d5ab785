"""Host-side integer glue of the Protogalaxy path, mirrored from the reference's own unit tests, for
both the oracle (oracle/pyref.py, oracle/pg_ref.py) and the product's host mirror (sirius_b200/protogalaxy.py):
  * lagrange::tests::correctness_for_cyclic_element (src/polynomial/lagrange.rs:95-114): L_i(w^j) = delta_ij on the
    order-2^8 subgroup, which goes through the 0/0 special case (:67-68);
  * lagrange::tests::basic_lagrange_test (:116-128), the 4 known-answer constants;
  * the UnivariatePoly::eval known answers (src/polynomial/univariate.rs:197-247)."""
import json
import os

import pytest

from oracle import pg_ref as G
from oracle import pyref as R
from sirius_b200 import protogalaxy as PG

IMPLS = [
    ("oracle", lambda X, log_n: R.eval_lagrange_polys(log_n, X), lambda log_n: list(R.iter_cyclic_subgroup(log_n)), G.poly_eval),
    ("product host mirror", lambda X, log_n: PG.eval_lagrange_polys(X, log_n), PG.iter_cyclic_subgroup, PG.poly_eval),
]


@pytest.mark.parametrize("name,lagrange,subgroup,poly_eval", IMPLS, ids=[i[0] for i in IMPLS])
def test_correctness_for_cyclic_element(name, lagrange, subgroup, poly_eval):
    LOG_N = 8
    pts = subgroup(LOG_N)
    assert len(pts) == 256 and len(set(pts)) == 256 and pts[0] == 1
    assert pow(pts[1], 256, R.FR) == 1 and pow(pts[1], 128, R.FR) != 1   # a generator of the order-2^8 subgroup
    for j, w_j in enumerate(pts):
        L = lagrange(w_j, LOG_N)
        assert L == [1 if i == j else 0 for i in range(256)], j


@pytest.mark.parametrize("name,lagrange,subgroup,poly_eval", IMPLS, ids=[i[0] for i in IMPLS])
def test_basic_lagrange_known_answers(name, lagrange, subgroup, poly_eval):
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lagrange_kat.json")))
    assert lagrange(kat["X"], kat["log_n"]) == [int(v, 16) for v in kat["expected_hex"]]
    # the same constants as the reference writes them (decimal, src/polynomial/lagrange.rs:119-124)
    assert int(kat["expected_hex"][0], 16) == 5472060717959818805561601436314318772137091100104008585924551046643952123908


@pytest.mark.parametrize("name,lagrange,subgroup,poly_eval", IMPLS, ids=[i[0] for i in IMPLS])
def test_univariate_eval_known_answers(name, lagrange, subgroup, poly_eval):
    assert poly_eval([5], 10) == 5                      # test_constant_polynomial
    assert poly_eval([3, 2], 4) == 11                   # test_linear_polynomial
    assert poly_eval([3, 2, 1], 2) == 11                # test_quadratic_polynomial
    coeff = [5, 1, 2, 3, 4]
    assert poly_eval(coeff, 2) == sum(c * 2**i for i, c in enumerate(coeff))   # test_high_degree_polynomial
    assert poly_eval([], 1) == 0                        # test_zero_polynomial
    assert poly_eval([R.FR - 1, 1], 1) == 0             # wraps modulo the field
