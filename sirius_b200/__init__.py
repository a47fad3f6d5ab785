"""sirius_b200 -- B200 (sm_100a) implementation of the Sirius folding-prover hot path.

Host-side mirror of the reference's Rust interface for the path (same names, argument meaning and error
behaviour), sitting on the C ABI in include/sirius_b200.h:

    commitment.CommitmentKey.commit      <- src/commitment.rs:81-90
    fft.{fft,ifft,coset_fft,coset_ifft}  <- src/fft.rs:160-198

Field elements are uint64 [.., 4] little-endian Montgomery limbs, affine points uint64 [.., 8] (x, y),
identity (0,0) -- the memory the Rust types hold.
"""
from ._lib import (  # noqa: F401
    CURVE_BN256,
    CURVE_GRUMPKIN,
    FIELD_FQ,
    FIELD_FR,
    SiriusB200Error,
    load,
)
from .commitment import CommitmentKey, TooLongInput, setup_smallest_key, smallest_key_log2  # noqa: F401
