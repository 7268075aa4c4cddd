"""ctypes loader for libhannoy_b200.so (C-ABI declared in include/hannoy_b200.h).

There is no Python/CPU fallback: if the CUDA library is missing the import of a compute entry point
fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("HB_LIB_VARIANT", "")  # dev builds only (hannoy_b200/build.py)
SO_PATH = os.path.join(_HERE, f"libhannoy_b200{'_' + _VARIANT if _VARIANT else ''}.so")

HB_OK, HB_EINVAL, HB_EDIM, HB_EFORMAT, HB_ECUDA, HB_ENOMEM, HB_ENCCL = range(7)
HB_EMISSING_METADATA, HB_EUNMATCHING_DISTANCE, HB_ENEED_BUILD, HB_ESTATE = 7, 8, 9, 10
HB_N_CTR = 8
CTR_DIST_UPPER, CTR_DIST_L0, CTR_EXP_UPPER, CTR_EXP_L0, CTR_DEG_UPPER, CTR_DEG_L0, CTR_FLAGS = range(7)
FLAG_FALLBACK, FLAG_LINEAR, FLAG_SLOW_PATH, FLAG_CANCELLED = 1, 2, 4, 8
LEN_CANCELLED, LEN_NONE = 0x80000000, 0xFFFFFFFF

# development entry points (include/hannoy_b200_dev.h): not part of the drop-in boundary
DEV_EXPORTS = ["hb_tune", "hb_debug_phases", "hb_debug_trace"]

# every symbol include/hannoy_b200.h declares
EXPORTS = [
    "hb_metric_name", "hb_metric_from_name", "hb_index_begin", "hb_index_push_kv", "hb_index_push_lmdb", "hb_index_open_lmdb",
    "hb_lmdb_scan", "hb_index_save", "hb_index_load", "hb_index_build_graph", "hb_index_export_kv", "hb_index_from_arrays",
    "hb_index_finalize", "hb_index_free", "hb_index_dimensions", "hb_index_n_items", "hb_index_n_entry_points",
    "hb_index_max_level", "hb_index_version", "hb_index_item_ids", "hb_index_contains_item", "hb_index_item_vector",
    "hb_search_by_vector", "hb_search_by_item", "hb_search_by_vector_device", "hb_exact_knn", "hb_merge_topk_device",
    "hb_launch_count", "hb_last_error", "hb_index_replicate", "hb_index_finalize_replicated", "hb_index_n_devices", "hb_index_device",
    "hb_index_n_layers", "hb_index_entry_points", "hb_index_layer_csr",
    "hb_cancel_token_create", "hb_cancel_token_cancel", "hb_cancel_token_reset", "hb_cancel_token_is_cancelled",
    "hb_cancel_token_free", "hb_shard_group_create", "hb_shard_group_connect", "hb_search_sharded_device", "hb_shard_group_free",
]


class QueryOpts(C.Structure):
    _fields_ = [("candidates", C.c_void_p), ("n_candidates", C.c_uint64), ("has_candidates", C.c_int),
                ("linear_below", C.c_uint32), ("linear_below_ratio", C.c_float),
                ("cancel", C.c_void_p), ("cancel_after_polls", C.c_uint64)]


class BuildOpts(C.Structure):
    _fields_ = [("M", C.c_uint32), ("M0", C.c_uint32), ("ef_construction", C.c_uint32), ("alpha", C.c_float),
                ("seed", C.c_uint64), ("batch_max", C.c_uint32), ("dimensions", C.c_uint32)]


KV_VISIT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_size_t, C.POINTER(C.c_ubyte), C.c_size_t)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -m hannoy_b200.build` (nvcc, sm_100a). "
            "hannoy_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, u16, u32, u64, i32, f32 = C.c_void_p, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int, C.c_float
    sz = C.c_size_t
    sig = {
        "hb_metric_name": (C.c_char_p, [i32]), "hb_metric_from_name": (i32, [C.c_char_p]),
        "hb_index_begin": (i32, [i32, u16, C.POINTER(vp)]),
        "hb_index_push_kv": (i32, [vp, C.c_char_p, sz, C.c_char_p, sz]),
        "hb_index_build_graph": (i32, [vp, C.POINTER(BuildOpts), i32, vp]),
        "hb_index_export_kv": (i32, [vp, i32, KV_VISIT, vp]),
        "hb_index_save": (i32, [vp, C.c_char_p]), "hb_index_load": (i32, [vp, C.c_char_p]),
        "hb_index_push_lmdb": (i32, [vp, C.c_char_p, C.c_char_p, C.POINTER(u64)]),
        "hb_index_open_lmdb": (i32, [C.c_char_p, C.c_char_p, i32, u16, i32, C.POINTER(vp)]),
        "hb_lmdb_scan": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, sz, KV_VISIT, vp, C.POINTER(u64)]),
        "hb_index_from_arrays": (i32, [vp, u32, vp, u64, vp, vp, u32, vp, vp, vp, u32, u32]),
        "hb_index_finalize": (i32, [vp, i32]), "hb_index_free": (None, [vp]),
        "hb_index_replicate": (i32, [vp, vp, i32]), "hb_index_finalize_replicated": (i32, [vp, vp, i32]),
        "hb_index_n_devices": (i32, [vp]), "hb_index_device": (i32, [vp, i32]),
        "hb_index_n_layers": (u32, [vp]), "hb_index_entry_points": (u32, [vp, vp, u32]),
        "hb_index_layer_csr": (i32, [vp, u32, vp, vp, u64, C.POINTER(u64)]),
        "hb_index_dimensions": (u32, [vp]), "hb_index_n_items": (u64, [vp]),
        "hb_index_n_entry_points": (u32, [vp]), "hb_index_max_level": (u32, [vp]),
        "hb_index_version": (i32, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "hb_index_item_ids": (u64, [vp, vp, u64]), "hb_index_contains_item": (i32, [vp, u32]),
        "hb_index_item_vector": (i32, [vp, u32, vp]),
        "hb_search_by_vector": (i32, [vp, vp, u64, u32, u32, u32, C.POINTER(QueryOpts), vp, vp, vp, vp]),
        "hb_search_by_item": (i32, [vp, vp, u64, u32, u32, C.POINTER(QueryOpts), vp, vp, vp, vp]),
        "hb_search_by_vector_device": (i32, [vp, vp, u64, u32, u32, vp, vp, vp, vp, vp]),
        "hb_exact_knn": (i32, [vp, vp, u64, u32, u32, vp, vp]),
        "hb_merge_topk_device": (i32, [i32, vp, vp, u32, u64, u32, vp, vp, vp, vp]),
        "hb_cancel_token_create": (i32, [i32, C.POINTER(vp)]), "hb_cancel_token_cancel": (i32, [vp]),
        "hb_cancel_token_reset": (i32, [vp]), "hb_cancel_token_is_cancelled": (i32, [vp]), "hb_cancel_token_free": (None, [vp]),
        "hb_shard_group_create": (i32, [i32, i32, i32, u64, u32, C.POINTER(vp), vp]),
        "hb_shard_group_connect": (i32, [vp, vp]),
        "hb_search_sharded_device": (i32, [vp, vp, vp, u64, u32, u32, vp, vp, vp, vp]),
        "hb_shard_group_free": (None, [vp]),
        "hb_tune": (i32, [C.c_char_p, i32]), "hb_debug_phases": (None, [vp]), "hb_debug_trace": (u32, [vp, u32]), "hb_launch_count": (u64, []), "hb_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name, None)
        if fn is None:
            if _VARIANT:   # dev builds of older sources may lack newer entry points
                continue
            raise ImportError(f"{SO_PATH} lacks {name}: rebuild it with `python -m hannoy_b200.build`")
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L
