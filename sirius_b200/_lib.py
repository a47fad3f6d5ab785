"""ctypes binding of libsirius_b200.so (the C ABI in include/sirius_b200.h).

There is no CPU fallback: if the CUDA library is missing or no device is present every call raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsirius_b200.so")

SB_OK = 0
SB_ERR_CUDA = -1
SB_ERR_ARG = -2
SB_ERR_OOM = -3
SB_ERR_TOO_LONG = -4
SB_ERR_NCCL = -5

FIELD_FR, FIELD_FQ = 0, 1
CURVE_BN256, CURVE_GRUMPKIN = 0, 1

u64p = ctypes.POINTER(ctypes.c_uint64)
vp = ctypes.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check the exports against include/sirius_b200.h
SIGNATURES = {
    "sb_last_error": (ctypes.c_char_p, []),
    "sb_version": (ctypes.c_int, []),
    "sb_init": (ctypes.c_int, [ctypes.c_int]),
    "sb_init_devices": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "sb_num_devices": (ctypes.c_int, []),
    "sb_shutdown": (None, []),
    "sb_device_count": (ctypes.c_int, []),
    "sb_stream_release": (None, [vp]),
    "sb_ck_register": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(vp)]),
    "sb_ck_register_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, ctypes.c_int, vp, ctypes.POINTER(vp)]),
    "sb_ck_release": (None, [vp]),
    "sb_ck_add_window": (ctypes.c_int, [vp, ctypes.c_int, vp]),
    "sb_ck_len": (ctypes.c_size_t, [vp]),
    "sb_ck_window_bits": (ctypes.c_int, [vp]),
    "sb_msm_tune": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "sb_msm": (ctypes.c_int, [vp, u64p, ctypes.c_size_t, u64p]),
    "sb_msm_device": (ctypes.c_int, [vp, vp, ctypes.c_size_t, vp, vp, vp]),
    "sb_msm_batch": (ctypes.c_int, [vp, ctypes.POINTER(u64p), ctypes.c_size_t, ctypes.c_size_t, u64p]),
    "sb_msm_batch_device": (ctypes.c_int, [vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp, vp, vp]),
    "sb_comm_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(vp), ctypes.c_char_p]),
    "sb_comm_connect": (ctypes.c_int, [vp, ctypes.c_char_p]),
    "sb_comm_destroy": (None, [vp]),
    "sb_comm_status": (ctypes.c_int, [vp, vp]),
    "sb_comm_allsum_points_device": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.c_size_t, vp, vp]),
    "sb_msm_batch_sharded_device": (ctypes.c_int, [vp, vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp, vp]),
    "sb_points_on_curve": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_size_t, u64p]),
    "sb_points_on_curve_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, vp, vp]),
    "sb_index_multiples_device": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_uint64, ctypes.c_size_t, vp, vp]),
    "sb_msm_combine_batch_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, vp, vp]),
    "sb_msm_combine_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_int, vp, vp]),
    "sb_expr_compile": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int32), ctypes.c_size_t, ctypes.POINTER(vp)]),
    "sb_expr_free": (None, [vp]),
    "sb_expr_num_slots": (ctypes.c_uint32, [vp]),
    "sb_columns_register": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint32, ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8)), ctypes.c_size_t, ctypes.POINTER(u64p), ctypes.c_size_t, ctypes.POINTER(vp)]),
    "sb_columns_release": (None, [vp]),
    "sb_expr_eval": (ctypes.c_int, [vp, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(u64p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t, ctypes.POINTER(u64p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t, u64p, ctypes.c_size_t, u64p]),
    "sb_expr_eval_device": (ctypes.c_int, [vp, vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.c_size_t, u64p, ctypes.c_size_t, vp, vp]),
    "sb_cross_terms": (ctypes.c_int, [vp, ctypes.c_uint32, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(u64p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t, ctypes.POINTER(u64p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, ctypes.POINTER(u64p)]),
    "sb_cross_terms_device": (ctypes.c_int, [vp, ctypes.c_uint32, vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, vp, vp]),
    "sb_upload_rows_device": (ctypes.c_int, [vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp, vp]),
    "sb_cross_terms_rows_device": (ctypes.c_int, [vp, ctypes.c_uint32, vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp, vp]),
    "sb_axpy_fold": (ctypes.c_int, [ctypes.c_int, u64p, u64p, u64p, u64p, ctypes.c_size_t]),
    "sb_axpy_fold_device": (ctypes.c_int, [ctypes.c_int, vp, vp, u64p, vp, ctypes.c_size_t, vp]),
    "sb_error_fold": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.POINTER(u64p), ctypes.c_uint32, u64p, u64p, ctypes.c_size_t]),
    "sb_error_fold_device": (ctypes.c_int, [ctypes.c_int, vp, vp, ctypes.c_uint32, u64p, vp, ctypes.c_size_t, vp]),
    "sb_expr_field": (ctypes.c_int, [vp]),
    "sb_columns_log_rows": (ctypes.c_uint32, [vp]),
    "sb_pg_leaves_device": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_size_t, vp, ctypes.POINTER(vp), ctypes.c_size_t, ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, vp, vp]),
    "sb_beta_tree_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_size_t, u64p, vp, vp]),
    "sb_pg_tree": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_size_t, vp, ctypes.c_uint32, ctypes.POINTER(u64p), ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32), u64p]),
    "sb_lincomb": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(u64p), u64p, ctypes.c_size_t, ctypes.c_size_t, u64p]),
    "sb_lincomb_device": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(vp), u64p, ctypes.c_size_t, ctypes.c_size_t, vp, vp]),
    "sb_batch_invert": (ctypes.c_int, [ctypes.c_int, u64p, u64p, ctypes.c_size_t]),
    "sb_batch_invert_device": (ctypes.c_int, [ctypes.c_int, vp, vp, ctypes.c_size_t, vp]),
    "sb_lookup_multiplicity": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_size_t, u64p, ctypes.c_size_t, u64p]),
    "sb_lookup_multiplicity_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, vp, ctypes.c_size_t, vp, vp]),
    "sb_lookup_inverses": (ctypes.c_int, [ctypes.c_int, u64p, u64p, u64p, u64p, ctypes.c_size_t, u64p, u64p]),
    "sb_lookup_inverses_device": (ctypes.c_int, [ctypes.c_int, vp, vp, vp, u64p, ctypes.c_size_t, vp, vp, vp]),
    "sb_scaled_inverse": (ctypes.c_int, [ctypes.c_int, u64p, u64p, u64p, u64p, ctypes.c_size_t]),
    "sb_scaled_inverse_device": (ctypes.c_int, [ctypes.c_int, vp, u64p, vp, vp, ctypes.c_size_t, vp]),
    "sb_sum_diff": (ctypes.c_int, [ctypes.c_int, u64p, u64p, ctypes.c_size_t, u64p]),
    "sb_sum_diff_device": (ctypes.c_int, [ctypes.c_int, vp, vp, ctypes.c_size_t, vp, vp]),
    "sb_count_mismatch_device": (ctypes.c_int, [ctypes.c_int, vp, vp, ctypes.c_size_t, vp, vp]),
    "sb_sparse_register": (ctypes.c_int, [ctypes.c_int, u64p, u64p, u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(vp)]),
    "sb_sparse_release": (None, [vp]),
    "sb_sparse_dim": (ctypes.c_size_t, [vp]),
    "sb_sparse_mismatch": (ctypes.c_int, [vp, u64p, ctypes.c_size_t, u64p]),
    "sb_sparse_mismatch_device": (ctypes.c_int, [vp, u64p, ctypes.c_size_t, vp, ctypes.c_size_t, u64p, vp]),
    "sb_concat_pad_device": (ctypes.c_int, [ctypes.POINTER(u64p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), vp]),
    "sb_ntt": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_uint32, u64p, u64p]),
    "sb_ntt_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_uint32, u64p, u64p, vp]),
    "sb_coset_scale": (ctypes.c_int, [ctypes.c_int, u64p, ctypes.c_size_t, u64p, u64p]),
    "sb_coset_scale_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, u64p, u64p, vp]),
    "sb_launch_count": (ctypes.c_uint64, []),
    "sb_profile_enable": (None, [ctypes.c_int]),
    "sb_profile_collect": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    "sb_microbench": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "sb_expr_jit_enable": (None, [ctypes.c_int]),
    "sb_expr_jit_selftest": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int32), ctypes.c_size_t, ctypes.c_uint32,
                                            ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "sb_selftest_coop": (ctypes.c_int, [ctypes.c_int, u64p, u64p, ctypes.c_size_t, u64p, u64p]),
    "sb_selftest_lazy": (ctypes.c_int, [ctypes.c_int, u64p, u64p, ctypes.c_size_t, u64p, u64p, u64p, u64p]),
    "sb_selftest_field": (ctypes.c_int, [ctypes.c_int, u64p, u64p, ctypes.c_size_t, u64p, u64p, u64p, u64p, u64p]),
}



class sb_calc(ctypes.Structure):
    _fields_ = [
        ("opcode", ctypes.c_uint8), ("a_kind", ctypes.c_uint8), ("b_kind", ctypes.c_uint8), ("_pad", ctypes.c_uint8),
        ("a_index", ctypes.c_uint32), ("a_rot", ctypes.c_uint32), ("b_index", ctypes.c_uint32), ("b_rot", ctypes.c_uint32),
        ("target", ctypes.c_uint32),
    ]


_lib = None


class SiriusB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsirius_b200 error {code}: {msg}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the CUDA library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(sirius_b200 has no CPU fallback)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != SB_OK:
        raise SiriusB200Error(rc, load().sb_last_error().decode(errors="replace"))
