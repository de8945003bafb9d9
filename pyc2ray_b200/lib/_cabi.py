"""ctypes binding of libasora_b200.so (include/asora_b200.h).  No fallback: a missing or unloadable
library is an ImportError, and every non-zero return code becomes a RuntimeError."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasora_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `python pyc2ray_b200/_build.py` (there is no CPU fallback)")

_L = ctypes.CDLL(LIB_PATH)

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
_i, _d, _i64 = ctypes.c_int, ctypes.c_double, ctypes.c_int64

# name -> (restype, argtypes): exactly the symbols include/asora_b200.h declares
SIGNATURES = {
    "asora_device_init": (_i, [_i, _i]),
    "asora_device_close": (_i, []),
    "asora_density_to_device": (_i, [c_dp, _i]),
    "asora_photo_table_to_device": (_i, [c_dp, c_dp, _i]),
    "asora_source_data_to_device": (_i, [c_ip, c_dp, _i]),
    "asora_do_all_sources": (_i, [_d, _d, _d, c_dp, c_dp, _i, _i, _d, _d, _i]),
    "asora_do_all_sources_begin": (_i, [_d, _d, _d, c_dp, _i, _i, _d, _d, _i]),
    "asora_do_all_sources_end": (_i, [c_dp]),
    "asora_invalidate_temperature": (_i, []),
    "asora_heat_table_to_device": (_i, [c_dp, c_dp, _i]),
    "asora_set_heating": (_i, [_i]),
    "asora_do_all_sources_heat": (_i, [_d, _d, _d, c_dp, c_dp, c_dp, _i, _i, _d, _d, _i]),
    "asora_global_pass": (_i, [_d, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, _d, _d, _d, _d, _d, _i64,
                               ctypes.POINTER(_i)]),
    "asora_device_buffer": (ctypes.c_void_p, [_i]),
    "asora_buffer_upload": (_i, [_i, c_dp]),
    "asora_buffer_upload_f": (_i, [_i, c_dp]),
    "asora_buffer_download_f": (_i, [_i, c_dp]),
    "asora_buffer_upload_range": (_i, [_i, c_dp, _i64, _i64]),
    "asora_buffer_download": (_i, [_i, c_dp]),
    "asora_buffer_copy": (_i, [_i, _i]),
    "asora_ipc_export": (_i, [_i, ctypes.c_char_p]),
    "asora_ipc_open": (_i, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "asora_ipc_close": (_i, [ctypes.c_void_p]),
    "asora_peer_halo": (_i, [_i, ctypes.c_void_p, _i64, _i64, _i]),
    "asora_raytrace_device": (_i, [_d, _d, _d, _i, _i, _d, _d, _i, _i]),
    "asora_global_pass_device": (_i, [_d, _d, _d, _d, _d, _d, ctypes.POINTER(_i), c_dp, c_dp]),
    "asora_set_active_slab": (_i, [_i, _i]),
    "asora_global_pass_device_range": (_i, [_d, _d, _d, _d, _d, _d, _i64, _i64, ctypes.POINTER(_i), c_dp, c_dp]),
    "asora_sync": (_i, []),
    "asora_set_stream": (_i, [ctypes.c_void_p]),
    "asora_debug_single_source": (_i, [_d, _d, _d, c_dp, _i, _d, _d, _i, c_dp, c_dp]),
    "asora_set_sweep_variant": (_i, [_i]),
    "asora_set_tuning": (_i, [_i, _i]),
    "asora_set_sphere_only": (_i, [_i]),
    "asora_set_deterministic": (_i, [_i]),
    "asora_set_grey_notables": (_i, [_i]),
    "asora_set_octant_shape": (_i, [_i, _i, _i, _i]),
    "asora_set_cluster_shape": (_i, [_i, _i]),
    "asora_plan_builds": (_i, []),
    "asora_plan_export": (_i64, [_i, _d, _d, _i, _i, _i, _i64, c_dp, c_dp, ctypes.POINTER(ctypes.c_uint16),
                                 ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint8),
                                 ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(_i), ctypes.POINTER(_i),
                                 ctypes.POINTER(_i)]),
    "asora_last_sweep_stats": (_i, [ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i64),
                                    ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(ctypes.c_float)]),
    "asora_last_sweep_kernel_ms": (_i, [ctypes.POINTER(ctypes.c_float)]),
    "asora_cells_per_source": (_i64, [_i, _d]),
    "asora_last_error": (ctypes.c_char_p, []),
    "asora_version": (ctypes.c_char_p, []),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(_L, _name)
    _f.restype = _res
    _f.argtypes = _args

BUF_NDENS, BUF_XH_AV, BUF_PHI_ION, BUF_XH, BUF_XH_INTERMED, BUF_TEMP, BUF_COLDENS, BUF_PHI_HEAT = range(8)


def check(rc):
    if rc != 0:
        raise RuntimeError("libasora_b200: " + _L.asora_last_error().decode())


def dptr(a):
    return a.ctypes.data_as(c_dp)


def iptr(a):
    return a.ctypes.data_as(c_ip)


L = _L
