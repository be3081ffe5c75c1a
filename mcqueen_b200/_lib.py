"""ctypes binding of libmcq_b200.so (include/mcq_b200.h).  The library is the product; this
module only loads it.  It raises - it never falls back to a CPU path - when the shared object
is missing."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmcq_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_char_pp = C.POINTER(C.c_char_p)


class McqDims(C.Structure):
    _fields_ = [("num_sites", C.c_int), ("closest_n", C.c_int), ("total_num_refs", C.c_int),
                ("num_query", C.c_int), ("num_hla_types", C.c_int), ("device", C.c_int)]


class McqParams(C.Structure):
    _fields_ = [("mu", C.c_double), ("kappa", C.c_double), ("phi", C.c_double),
                ("prior_C2", c_double_p), ("R", c_double_p), ("omega", c_double_p),
                ("rate_out_of_codons", c_double_p), ("HLA_select_esc", c_double_p),
                ("HLA_select_rev", c_double_p)]


class McqProposal(C.Structure):
    _fields_ = [("kind", C.c_int), ("start", C.c_int), ("end", C.c_int), ("hla", C.c_int),
                ("split", C.c_int), ("value0", C.c_double), ("value1", C.c_double)]


# every symbol include/mcq_b200.h declares: name -> (restype, argtypes)
_V = C.c_void_p
SYMBOLS = {
    "initialise_F_numerical_recipes": (None, [C.c_int] * 3 + [_V] * 6 + [C.c_int] * 2 + [_V, C.c_int, _V]),
    "update_F_numerical_recipes": (None, [C.c_int] * 3 + [_V] * 8 + [C.c_int] * 2 + [_V, C.c_int, _V]),
    "update_B_numerical_recipes": (None, [C.c_int] * 3 + [_V] * 6 + [C.c_int, _V, C.c_int, _V]),
    "hamming_distance_considering_unknown": (None, [_V, _V, C.c_int, C.c_int, _V]),
    "closest_sequences": (None, [C.c_int, _V, C.c_int, _V]),
    "mcq_last_error": (C.c_char_p, []),
    "mcq_device_count": (C.c_int, []),
    "mcq_abi_version": (C.c_int, []),
    "mcq_ctx_create": (C.c_int, [C.POINTER(_V), C.POINTER(McqDims)]),
    "mcq_ctx_destroy": (None, [_V]),
    "mcq_ctx_upload_static": (C.c_int, [_V] * 9),
    "mcq_ctx_set_params": (C.c_int, [_V, C.POINTER(McqParams)]),
    "mcq_ctx_emission": (C.c_int, [_V, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mcq_ctx_init_state": (C.c_int, [_V, c_double_p]),
    "mcq_ctx_full_llk": (C.c_int, [_V, C.c_int, c_double_p]),
    "mcq_ctx_llk_only": (C.c_int, [_V, C.c_int, c_double_p]),
    "mcq_ctx_tune": (C.c_int, [_V, c_int_p]),
    "mcq_ctx_backward_fill": (C.c_int, [_V]),
    "mcq_ctx_propose": (C.c_int, [_V, C.POINTER(McqProposal), c_double_p]),
    "mcq_ctx_accept": (C.c_int, [_V]),
    "mcq_ctx_reject": (C.c_int, [_V]),
    "mcq_ctx_flush": (C.c_int, [_V]),
    "mcq_ctx_set_option": (C.c_int, [_V, C.c_int, C.c_int]),
    "mcq_ctx_get_matrix": (C.c_int, [_V, C.c_int, C.c_int, _V, _V]),
    "mcq_ctx_get_table": (C.c_int, [_V, C.c_int, C.c_int, _V]),
    "mcq_ctx_get_query_llk": (C.c_int, [_V, C.c_int, _V]),
    "mcq_ctx_get_params": (C.c_int, [_V, c_double_p, _V, _V, _V, _V, _V]),
    "mcq_ctx_device_bytes": (C.c_size_t, [_V]),
    "mcq_ctx_last_timing": (C.c_int, [_V, C.POINTER(C.c_float), c_int_p]),
    "mcq_ctx_enable_timing": (C.c_int, [_V, C.c_int]),
    "mcq_ctx_kernel_time": (C.c_int, [_V, C.c_int, C.POINTER(C.c_float), c_int_p, C.c_int]),
    "mcq_ctx_timer_start": (C.c_int, [_V]),
    "mcq_ctx_timer_stop": (C.c_int, [_V, C.POINTER(C.c_float)]),
    "mcq_ctx_synchronize": (C.c_int, [_V]),
    "mcq_host_register": (C.c_int, [_V, C.c_size_t]),
    "mcq_host_unregister": (C.c_int, [_V]),
    "mcq_ctx_peer_handle": (C.c_int, [_V, _V]),
    "mcq_ctx_attach_peers": (C.c_int, [_V, C.c_int, C.c_int, _V]),
    "mcq_comm_unique_id": (C.c_int, [_V]),
    "mcq_ctx_attach_comm": (C.c_int, [_V, C.c_int, C.c_int, _V]),
    "mcq_ctx_exchange_sum": (C.c_int, [_V, c_double_p]),
    "mcq_codon_class": (C.c_int, [C.c_int, C.c_int]),
    "mcq_closest_batch": (C.c_int, [C.c_int, _V, C.c_size_t, C.c_int, _V, C.c_size_t, C.c_int, C.c_int,
                                    C.c_int, C.c_int, _V]),
    "mcq_closest_groups": (C.c_int, [C.c_int, _V, C.c_size_t, C.c_int, _V, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_size_t)]),
    "mcq_closest_replay": (C.c_int, [_V, C.c_size_t, C.c_int, C.c_int, _V, c_int_p]),
    "mcq_free": (None, [_V]),
    "mcq_consensus_batch": (C.c_int, [C.c_int, _V, C.c_size_t, C.c_int, _V, C.c_int, C.c_int, C.c_int, _V]),
    "mcq_consensus_last_timing": (C.c_int, [C.POINTER(C.c_float)]),
    "mcq_closest_last_timing": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "mcq_hamming_batch": (C.c_int, [C.c_int, _V, C.c_size_t, C.c_int, _V, C.c_size_t, C.c_int, C.c_int, _V]),
}

_lib = None


def load():
    """Returns the loaded library; raises RuntimeError (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m mcqueen_b200.build` "
            "(there is no CPU fallback for the McQueen likelihood path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class McqError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise McqError(load().mcq_last_error().decode() or f"libmcq_b200 call failed ({rc})")
