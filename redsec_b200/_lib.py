"""ctypes binding of include/redsec_b200.h.  Loading fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libredsec_b200.so")

u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
i8p = C.POINTER(C.c_int8)
vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/redsec_b200.h declares
SIGNATURES = {
    "rs_ctx_create": (C.c_int, [C.POINTER(vp), C.c_int]),
    "rs_ctx_destroy": (C.c_int, [vp]),
    "rs_ctx_retain": (C.c_int, [vp]),
    "rs_ctx_release": (C.c_int, [vp]),
    "rs_pool_trim": (C.c_int, [vp]),
    "rs_lanes": (C.c_int, [vp, C.c_int]),
    "rs_lane_count": (C.c_int, [vp]),
    "rs_lane_select": (C.c_int, [vp, C.c_int]),
    "rs_reserve_scratch": (C.c_int, [vp, C.c_size_t]),
    "rs_lane_fork": (C.c_int, [vp]),
    "rs_lane_join": (C.c_int, [vp]),
    "rs_last_error": (C.c_char_p, [vp]),
    "rs_set_stream": (C.c_int, [vp, vp]),
    "rs_sync": (C.c_int, [vp]),
    "rs_load_eval_key": (C.c_int, [vp, u32p, u32p]),
    "rs_lwe_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "rs_lwe_free": (C.c_int, [vp, vp]),
    "rs_lwe_upload": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_lwe_download": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_lwe_copy": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_lwe_axpby": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32]),
    "rs_lwe_add_bias": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_int]),
    "rs_get_stream": (C.c_int, [vp, C.POINTER(vp)]),
    "rs_ctx_device": (C.c_int, [vp]),
    "rs_comm_unique_id": (C.c_int, [vp]),
    "rs_comm_init_rank": (C.c_int, [vp, C.POINTER(vp), vp, C.c_int, C.c_int]),
    "rs_comm_init_all": (C.c_int, [C.POINTER(vp), C.c_int, C.POINTER(vp)]),
    "rs_comm_destroy": (C.c_int, [vp]),
    "rs_comm_rank": (C.c_int, [vp]),
    "rs_comm_world": (C.c_int, [vp]),
    "rs_comm_ctx": (vp, [vp]),
    "rs_comm_last_error": (C.c_char_p, []),
    "rs_allgather": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_net_shard_plan": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "rs_net_layer_forward_sharded": (C.c_int, [vp, C.c_int, vp, vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "rs_net_run": (C.c_int, [vp, vp, vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "rs_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "rs_host_free": (C.c_int, [vp]),
    "rs_pbs_batch": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_uint32]),
    "rs_pbs_lut_batch": (C.c_int, [vp, vp, vp, C.c_size_t, vp, C.c_int]),
    "rs_gate_batch": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_size_t, C.c_uint32]),
    "rs_pbs_batch_host": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_uint32]),
    "rs_gate_batch_host": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_size_t, C.c_uint32]),
    "rs_blind_rotate_batch": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_uint32]),
    "rs_keyswitch_batch": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_ext_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "rs_ext_upload": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_ext_download": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_lwe_lincomb": (C.c_int, [vp, vp, C.c_size_t, vp, vp, vp, vp, vp]),
    "rs_lwe_add_const": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32]),
    "rs_dev_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "rs_dev_free": (C.c_int, [vp, vp]),
    "rs_dev_upload": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_dev_download": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "rs_lwe_conv": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "rs_lwe_interleave": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_int, C.c_int]),
    "rs_modswitch_to_torus32": (C.c_uint32, [C.c_int32, C.c_int32]),
    "rs_modswitch_from_torus32": (C.c_int32, [C.c_uint32, C.c_int32]),
    "rs_selftest_chacha20": (C.c_int, [vp, C.c_uint64, C.c_uint64, vp, C.c_size_t]),
    "rs_keygen_secure": (C.c_int, [vp, vp, vp, vp]),
    "rs_lwe_encrypt_secure": (C.c_int, [vp, vp, C.c_size_t, C.c_double, vp]),
    "rs_keygen": (C.c_int, [C.c_uint64, vp, vp, vp, vp]),
    "rs_lwe_encrypt": (C.c_int, [vp, vp, C.c_size_t, C.c_double, vp, C.c_uint64]),
    "rs_lwe_phase": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "rs_lwe_decrypt": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_int32]),
    "rs_write_secret_key": (C.c_int, [C.c_char_p, vp, vp]),
    "rs_read_secret_key": (C.c_int, [C.c_char_p, vp, vp]),
    "rs_write_eval_key": (C.c_int, [C.c_char_p, vp, vp]),
    "rs_read_eval_key": (C.c_int, [C.c_char_p, vp, vp]),
    "rs_write_ctxt": (C.c_int, [C.c_char_p, vp, C.c_size_t, C.c_double, C.c_int]),
    "rs_read_ctxt": (C.c_int, [C.c_char_p, vp, C.c_size_t]),
    "rs_shard_range": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rs_net_create": (vp, [vp]),
    "rs_net_destroy": (None, [vp]),
    "rs_net_add_layer": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "rs_net_prep": (C.c_int, [vp, C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "rs_net_prep_ex": (C.c_int, [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]),
    "rs_net_build_tables": (C.c_int, [vp, C.c_int, C.c_int]),
    "rs_net_num_layers": (C.c_int, [vp]),
    "rs_net_layer_info": (C.c_int, [vp, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rs_net_layer_forward": (C.c_int, [vp, C.c_int, vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rs_profile_enable": (C.c_int, [vp, C.c_int]),
    "rs_profile_get": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "rs_profile_reset": (C.c_int, [vp]),
    "rs_launch_count": (C.c_uint64, [vp]),
    "rs_fp64_peak": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "rs_fp64_peak_three_operand": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "rs_set_tuning": (C.c_int, [vp, C.c_int]),
    "rs_set_ks_variant": (C.c_int, [vp, C.c_int]),
    "rs_device_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
}

_lib = None


def load():
    """dlopen the in-tree CUDA library; raises if it is missing (build with __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise RuntimeError(
                f"{SO} is missing: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback."
            )
        lib = C.CDLL(SO)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
