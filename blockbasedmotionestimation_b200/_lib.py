"""ctypes binding of libbbme.so (the C ABI declared in include/bbme.h).

The library is built in-tree by `make` / `__graft_entry__.build()`.  There is no fallback: if the shared
object is missing, importing this module raises, and if no sm_100 device is present `bbme_create` fails.
"""
import ctypes as C
import os

BBME_MAX_LEVELS = 16
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BBME_LIB") or os.path.join(_HERE, "libbbme.so")  # BBME_LIB: tuning builds of the same library


class BbmeShape(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("padded_width", C.c_int), ("padded_height", C.c_int),
        ("padding_x", C.c_int), ("padding_y", C.c_int),
        ("num_levels", C.c_int),
        ("level_width", C.c_int * BBME_MAX_LEVELS),
        ("level_height", C.c_int * BBME_MAX_LEVELS),
        ("block_size", C.c_int * BBME_MAX_LEVELS),
        ("search_size", C.c_int * BBME_MAX_LEVELS),
    ]


class BbmeStats(C.Structure):
    _fields_ = [
        ("ms_total", C.c_float), ("ms_pyramid", C.c_float), ("ms_search", C.c_float),
        ("ms_regularize", C.c_float), ("ms_other", C.c_float),
        ("kernel_launches", C.c_uint32), ("fix_rounds", C.c_uint32), ("fix_blocks", C.c_uint32),
        ("search_kernel_used", C.c_uint32), ("search_launches", C.c_uint32), ("reserved", C.c_uint32),
        ("search_candidates", C.c_uint64), ("search_absdiffs", C.c_uint64),
    ]


class BbmeOptions(C.Structure):
    _fields_ = [
        ("sweeps", C.c_int), ("chunk_pairs", C.c_int), ("slots", C.c_int),
        ("search_kernel", C.c_int), ("collect_stats", C.c_int), ("keep_search_mv", C.c_int), ("search_variant", C.c_int),
    ]


class BbmeHostLink(C.Structure):
    _fields_ = [("h2d_gbs", C.c_double), ("d2h_gbs", C.c_double), ("duplex_gbs_per_direction", C.c_double),
                ("host_stream_write_gbs", C.c_double), ("host_threads", C.c_int)]


class BbmeSearchGeometry(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("planned", "copies", "deep_ring", "key64", "rows_per_lane", "pitch_words", "stages", "stage_bytes",
                                       "smem_bytes", "bands", "segments_per_band", "box_w", "box_h", "two_boxes", "lanes_per_unit")]


# name -> (restype, argtypes); every symbol include/bbme.h declares
_P = C.c_void_p
_I = C.c_int
_SZ = C.c_size_t
SIGNATURES = {
    "bbme_default_options": (None, [C.POINTER(BbmeOptions)]),
    "bbme_version": (_I, []),
    "bbme_status_string": (C.c_char_p, [_I]),
    "bbme_plan_shape": (_I, [_I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(BbmeShape)]),
    "bbme_create": (_I, [C.POINTER(_P), _I]),
    "bbme_destroy": (None, [_P]),
    "bbme_last_error": (C.c_char_p, [_P]),
    "bbme_plan": (_I, [_P, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(BbmeOptions), C.POINTER(BbmeShape)]),
    "bbme_estimate": (_I, [_P, _P, _P, _SZ, _P]),
    "bbme_estimate_batch": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_P), _SZ, C.POINTER(_P)]),
    "bbme_estimate_batch_async": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_P), _SZ, C.POINTER(_P)]),
    "bbme_estimate_sequence": (_I, [_P, _I, C.POINTER(_P), _SZ, C.POINTER(_P)]),
    "bbme_estimate_sequence_async": (_I, [_P, _I, C.POINTER(_P), _SZ, C.POINTER(_P)]),
    "bbme_estimate_sequence_device": (_I, [_P, _I, _P, _SZ, _SZ, _P, _SZ]),
    "bbme_estimate_upsampled": (_I, [_P, _I, _I, C.POINTER(_P), C.POINTER(_P), _SZ, C.POINTER(_P)]),
    "bbme_estimate_upsampled_async": (_I, [_P, _I, _I, C.POINTER(_P), C.POINTER(_P), _SZ, C.POINTER(_P)]),
    "bbme_estimate_upsampled_device": (_I, [_P, _I, _I, _P, _P, _SZ, _SZ, _P, _SZ]),
    "bbme_estimate_device": (_I, [_P, _I, _P, _P, _SZ, _SZ, _P, _SZ]),
    "bbme_estimate_device_compact": (_I, [_P, _I, _P, _P, _SZ, _SZ, _P, _SZ]),
    "bbme_estimate_device_both": (_I, [_P, _I, _P, _P, _SZ, _SZ, _P, _SZ, _P, _SZ]),
    "bbme_expand_compact": (_I, [_P, _I, _I, _P]),
    "bbme_sync": (_I, [_P]),
    "bbme_get_stats": (_I, [_P, C.POINTER(BbmeStats)]),
    "bbme_set_streams": (_I, [_P, _I, C.POINTER(_P)]),
    "bbme_measure_int_peak": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "bbme_get_shape": (_I, [_P, C.POINTER(BbmeShape)]),
    "bbme_measure_host_link": (_I, [_P, _SZ, C.POINTER(BbmeHostLink)]),
    "bbme_mf_open": (_I, [C.POINTER(_P), _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), _I, C.POINTER(BbmeShape)]),
    "bbme_mf_close": (None, [_P]),
    "bbme_mf_cache_clear": (None, []),
    "bbme_pool_create": (_I, [C.POINTER(_P), _I, C.POINTER(_I)]),
    "bbme_pool_destroy": (None, [_P]),
    "bbme_pool_device_count": (_I, [_P]),
    "bbme_pool_last_error": (C.c_char_p, [_P]),
    "bbme_pool_plan": (_I, [_P, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(BbmeOptions), C.POINTER(BbmeShape)]),
    "bbme_pool_estimate_batch": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_P), _SZ, C.POINTER(_P)]),
    "bbme_device_alloc": (_I, [_I, _SZ, C.POINTER(_P)]),
    "bbme_device_free": (None, [_I, _P]),
    "bbme_ipc_export": (_I, [_I, _P, _P]),
    "bbme_ipc_open": (_I, [_I, _P, C.POINTER(_P)]),
    "bbme_ipc_close": (_I, [_I, _P]),
    "bbme_copy_async": (_I, [_I, _P, _P, _SZ, _P]),
    "bbme_host_alloc": (_I, [C.POINTER(_P), _SZ]),
    "bbme_host_free": (None, [_P]),
    "bbme_debug_skip_compute": (_I, [_P, _I]),
    "bbme_debug_set_stamp_epoch": (_I, [_P, C.c_uint32]),
    "bbme_debug_search_geometry": (_I, [_I, _I, _I, _I, _I, C.POINTER(BbmeSearchGeometry)]),
    "bbme_debug_div_magic": (_I, [C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "bbme_debug_level_image": (_I, [_P, _I, _I, _I, _P]),
    "bbme_debug_level_mv": (_I, [_P, _I, _I, _I, _P]),
    "bbme_stage_pyrdown": (_I, [_P, _P, _I, _I, _P]),
    "bbme_stage_resize": (_I, [_P, _P, _I, _I, _I, _P]),
    "bbme_stage_search": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _I, C.POINTER(BbmeStats)]),
    "bbme_stage_search_raster": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "bbme_stage_compensate": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "bbme_stage_regularize": (_I, [_P, _P, _P, _I, _I, _I, C.c_float, _I, _P, C.POINTER(C.c_uint32)]),
    "bbme_stage_divide": (_I, [_P, _P, _I, _I, _P]),
    "bbme_stage_copy_mvs": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "bbme_flo_read_header": (_I, [C.c_char_p, C.POINTER(_I), C.POINTER(_I)]),
    "bbme_flo_read": (_I, [C.c_char_p, _P, _I, _I]),
    "bbme_flo_write": (_I, [C.c_char_p, _P, _I, _I]),
    "bbme_flow_aee": (C.c_double, [_P, _P, _I, _I]),
    "bbme_flow_to_color": (_I, [_P, _I, _I, C.c_float, _P, _P]),
    "bbme_flow_strip_subsample": (_I, [_P, C.POINTER(BbmeShape), _I, _P]),
}

_lib = None


def load():
    """Load libbbme.so and bind every declared symbol.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
