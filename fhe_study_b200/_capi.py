"""ctypes binding of libfhe_b200.so (the C ABI in include/fhe_b200.h).

There is NO fallback: if the CUDA library cannot be loaded, importing this module raises.  Nothing in
this package imports ``oracle`` (the CPU restatement is test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FHE_B200_LIB: a tuning variant of the same CUDA library (see build.py), for A/B measurements only.
LIB_PATH = os.environ.get("FHE_B200_LIB") or os.path.join(_HERE, "libfhe_b200.so")

U64 = C.c_uint64
SZ = C.c_size_t
P = C.c_void_p
I = C.c_int


class FheError(RuntimeError):
    """Raised where the reference panics / returns Err (the C call returned non-zero)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m fhe_study_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    return C.CDLL(LIB_PATH)


lib = _load()

class FileInfo(C.Structure):
    """fhe_file_info of include/fhe_b200_file.h"""

    _fields_ = [("version", C.c_uint32), ("kind", C.c_uint32), ("encoding", C.c_uint32), ("bits", C.c_uint32),
                ("q", U64), ("n", U64), ("k", U64), ("l", U64), ("count", U64), ("words_per_object", U64),
                ("payload_bytes", U64), ("checksum", U64)]


# name -> (restype, argtypes); mirrors include/fhe_b200.h one to one (tests/test_capi_symbols.py checks
# that the header, this table and the .so agree).
SIGNATURES = {
    "fhe_last_error": (C.c_char_p, []),
    "fhe_device_count": (I, [C.POINTER(I)]),
    "fhe_set_device": (I, [I]),
    "fhe_current_device": (I, [C.POINTER(I)]),
    "fhe_set_stream": (I, [P]),
    "fhe_synchronize": (I, []),
    "fhe_launch_count": (U64, []),
    "fhe_int_peak": (I, [I, C.POINTER(C.c_double)]),
    "fhe_ntt_plan_create": (I, [U64, U64, C.POINTER(P)]),
    "fhe_ntt_plan_destroy": (None, [P]),
    "fhe_ntt_plan_info": (I, [P, C.POINTER(U64), C.POINTER(U64), P, P]),
    "fhe_ntt_plan_config": (I, [P, C.POINTER(I)]),
    "fhe_rq_check_canonical": (I, [U64, P, SZ, C.POINTER(U64)]),
    "fhe_ntt_fwd": (I, [P, P, P, SZ]),
    "fhe_ntt_inv": (I, [P, P, P, SZ]),
    "fhe_rq_mul": (I, [P, P, P, P, SZ, I, P]),
    "fhe_ntt_fwd_u32": (I, [P, P, P, SZ]),
    "fhe_ntt_inv_u32": (I, [P, P, P, SZ]),
    "fhe_rq_mul_u32": (I, [P, P, P, P, SZ, I, P]),
    "fhe_ntt_fwd_packed": (I, [P, I, P, P, SZ]),
    "fhe_ntt_inv_packed": (I, [P, I, P, P, SZ]),
    "fhe_rq_mul_packed": (I, [P, I, P, P, P, SZ, I, P]),
    "fhe_pack_bits": (I, [I, P, P, SZ]),
    "fhe_unpack_bits": (I, [I, P, P, SZ]),
    "fhe_bfv_keygen": (I, [P, U64, U64, C.c_double, U64, P, P]),
    "fhe_bfv_rlk_generate": (I, [U64, U64, U64, C.c_double, U64, P, P]),
    "fhe_bfv_mul_const": (I, [U64, U64, U64, U64, P, P, P, P, SZ]),
    "fhe_ckks_keygen": (I, [P, U64, U64, C.c_double, U64, P, P]),
    "fhe_ckks_encrypt": (I, [P, U64, U64, P, P, C.c_double, U64, P, SZ]),
    "fhe_ckks_decrypt": (I, [P, U64, U64, P, P, P, SZ]),
    "fhe_ckks_add": (I, [U64, U64, P, P, P, SZ]),
    "fhe_ckks_sub": (I, [U64, U64, P, P, P, SZ]),
    "fhe_compute_lookup_table": (I, [U64, U64, U64, P]),
    "fhe_file_payload_bytes": (U64, [C.POINTER(FileInfo)]),
    "fhe_file_write": (I, [C.c_char_p, C.POINTER(FileInfo), P]),
    "fhe_file_read_info": (I, [C.c_char_p, C.POINTER(FileInfo)]),
    "fhe_file_read_payload": (I, [C.c_char_p, P, SZ]),
    "fhe_tn_mul": (I, [U64, P, P, P, SZ]),
    "fhe_tn_add": (I, [P, P, P, SZ]),
    "fhe_tn_sub": (I, [P, P, P, SZ]),
    "fhe_tn_neg": (I, [P, P, SZ]),
    "fhe_tn_left_rotate": (I, [U64, P, P, U64, P, SZ]),
    "fhe_tggsw_load": (I, [U64, U64, P, C.POINTER(P)]),
    "fhe_tlwe_encrypt": (I, [U64, P, P, C.c_double, U64, I, P, SZ]),
    "fhe_tlwe_decrypt": (I, [U64, P, P, P, SZ]),
    "fhe_tglwe_encrypt": (I, [U64, U64, P, P, C.c_double, U64, I, P, SZ]),
    "fhe_tglwe_decrypt": (I, [U64, U64, P, P, P, SZ]),
    "fhe_tggsw_generate": (I, [U64, U64, P, P, C.c_double, U64, I, P, C.POINTER(P)]),
    "fhe_tggsw_destroy": (None, [P]),
    "fhe_extprod": (I, [P, P, P, SZ]),
    "fhe_cmux": (I, [P, P, P, P, SZ]),
    "fhe_ksk_load": (I, [U64, U64, U64, P, C.POINTER(P)]),
    "fhe_ksk_generate": (I, [U64, U64, U64, P, P, C.c_double, U64, I, C.POINTER(P)]),
    "fhe_ksk_export": (I, [P, P]),
    "fhe_ksk_destroy": (None, [P]),
    "fhe_key_switch": (I, [P, P, P, SZ]),
    "fhe_tlwe_mod_switch": (I, [P, U64, P, SZ]),
    "fhe_sample_extract": (I, [U64, U64, P, U64, P, SZ]),
    "fhe_blind_rotate": (I, [U64, U64, P, I, P, P, U64, P, SZ]),
    "fhe_bootstrap": (I, [U64, U64, P, P, P, U64, P, SZ]),
    "fhe_rq_glev_load": (I, [P, U64, U64, P, C.POINTER(P)]),
    "fhe_rq_glev_destroy": (None, [P]),
    "fhe_rq_glev_mul": (I, [P, P, P, SZ]),
    "fhe_glwe_rq_key_switch": (I, [P, C.c_uint32, C.c_uint32, P, P, SZ]),
    "fhe_cmux_chain": (I, [U64, U64, P, U64, I, P, P, P, SZ]),
    "fhe_bootstrap_chain": (I, [U64, U64, P, U64, I, P, P, P, U64, P, SZ]),
    "fhe_bfv_tensor": (I, [U64, U64, U64, P, P, P, SZ]),
    "fhe_bfv_relinearize": (I, [U64, U64, U64, P, P, P, SZ]),
    "fhe_bfv_mul_relin": (I, [U64, U64, U64, U64, P, P, P, P, SZ]),
    "fhe_bfv_encrypt": (I, [P, U64, U64, U64, P, P, C.c_double, U64, P, SZ]),
    "fhe_bfv_decrypt": (I, [P, U64, U64, U64, P, P, P, SZ]),
    "fhe_rq_add": (I, [U64, P, P, P, SZ]),
    "fhe_rq_sub": (I, [U64, P, P, P, SZ]),
    "fhe_rq_neg": (I, [U64, P, P, SZ]),
    "fhe_rq_mul_u64": (I, [U64, P, U64, P, SZ]),
    "fhe_rq_remodule": (I, [P, U64, P, SZ]),
    "fhe_rq_mod_switch": (I, [U64, P, U64, P, SZ]),
    "fhe_rq_mul_div_round": (I, [U64, P, U64, U64, P, SZ]),
    "fhe_rq_from_vec": (I, [U64, U64, P, U64, P, SZ]),
    "fhe_rq_decompose": (I, [U64, U64, P, C.c_uint32, C.c_uint32, P, SZ]),
    "fhe_tn_decompose": (I, [U64, P, C.c_uint32, P, SZ]),
    "fhe_tn_mod_switch": (I, [P, U64, P, SZ]),
    "fhe_tn_mul_u64": (I, [P, U64, P, SZ]),
    "fhe_tn_mul_div_round": (I, [P, U64, U64, P, SZ]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


def check(rc: int) -> None:
    if rc != 0:
        raise FheError(lib.fhe_last_error().decode("utf-8", "replace") + f" (code {rc})")


def ptr(x):
    """Address of a numpy array (host) or a torch tensor (host or CUDA); None -> NULL."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):  # torch tensor
        return C.c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):  # numpy array
        return x.ctypes.data_as(C.c_void_p)
    if isinstance(x, int):
        return C.c_void_p(x)
    raise TypeError(f"cannot take the address of {type(x)!r}")
