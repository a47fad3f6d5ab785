"""GPU parity: sb_ntt / coset scaling through the C ABI vs the literal best_fft restatement (bit-exact)."""
import numpy as np
import pytest

from oracle import pyref as R

pytestmark = pytest.mark.gpu

# src/fft.rs:242-251
FFT_KAT = [
    28,
    68918385373930674424918168212551896122229959265833979749191472831399925654,
    17631683881184975370165255887551781615748388533673675138856,
    68918385373930639161550405842601155791718184162270748252414405484049647934,
    21888242871839275222246405745257275088548364400416034343698204186575808495613,
    21819324486465344583084855339414673932756646216253763595445789781091758847675,
    21888242871839275204614721864072299718383108512864252727949815652902133356753,
    21819324486465344547821487577044723192426134441150200363949012713744408569955,
]


@pytest.fixture(scope="module")
def fft():
    import sirius_b200
    from sirius_b200 import fft as _fft

    sirius_b200.load()
    return _fft


def test_fft_simple_input(fft):
    """reference fft_simple_input_test (src/fft.rs:240-260) run against the CUDA path."""
    a = R.to_mont_limbs(list(range(8)), R.FR)
    fft.fft(a)
    assert R.from_mont_limbs(a, R.FR) == FFT_KAT


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 17, 20, 21])
def test_fft_vs_oracle(fft, oracle, k):
    a = oracle.random_field(R.FIELD_FR, 300 + k, 1 << k)
    exp = oracle.fft(a, threads=8)
    got = fft.fft(a.copy())
    assert np.array_equal(got, exp)
    if k <= 17:
        assert np.array_equal(fft.ifft(a.copy()), oracle.ifft(a, threads=8))
        assert np.array_equal(fft.coset_fft(a.copy()), oracle.coset_fft(a, threads=8))
        assert np.array_equal(fft.coset_ifft(a.copy()), oracle.coset_ifft(a, threads=8))


@pytest.mark.parametrize("k", [4, 5, 6, 7, 8, 20])
def test_fft_roundtrip(fft, oracle, k):
    """reference fft_random_input_test / coset_fft_random_input_test (src/fft.rs:268-296)."""
    a = oracle.random_field(R.FIELD_FR, 900 + k, 1 << k)
    b = a.copy()
    fft.fft(b)
    fft.ifft(b)
    assert np.array_equal(a, b)
    fft.coset_fft(b)
    fft.coset_ifft(b)
    assert np.array_equal(a, b)
