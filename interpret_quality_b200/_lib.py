"""ctypes binding of libiq_b200.so (the C ABI declared in include/iq_b200.h).

The library is built in-tree by interpret_quality_b200/build.py (nvcc, sm_100a).
There is no CPU fallback: if the shared object is missing, or a compute entry is
called without a CUDA device, the call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libiq_b200.so")

_vp, _i64, _int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int

# name -> (restype, argtypes); mirrors include/iq_b200.h one to one
SIGNATURES = {
    "iq_version": (_int, []),
    "iq_last_error": (ctypes.c_char_p, []),
    "iq_launch_count": (ctypes.c_uint64, []),
    "iq_debug_reload_env": (_int, []),
    "iq_f16_paths": (_int, []),
    "iq_split_f16_host": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "iq_profile_enable": (_int, [_int]),
    "iq_profile_report": (_int, [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_double),
                                 ctypes.POINTER(ctypes.c_longlong), _int]),
    "iq_fps": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "iq_square_distance3": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "iq_region_id": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "iq_center": (_int, [_vp, _i64, _vp, _vp]),
    "iq_mask_shapley": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "iq_mask_interaction": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _int, _vp, _vp]),
    "iq_mask_interaction_pairs": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _int, _vp, _vp]),
    "iq_reward": (_int, [_vp, _i64, _i64, _i64, _int, _vp, _vp]),
    "iq_shapley_accumulate": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "iq_interaction_reduce": (_int, [_vp, _i64, _i64, _i64, _i64, _int, _vp, _vp]),
    "iq_model_create": (_vp, [ctypes.c_char_p, _int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_vp),
                              ctypes.POINTER(_i64), _int, _int]),
    "iq_model_destroy": (None, [_vp]),
    "iq_model_set_chunk": (_int, [_vp, _int]),
    "iq_model_set_lanes": (_int, [_vp, _int]),
    "iq_model_get_lanes": (_int, [_vp]),
    "iq_model_get_chunk": (_int, [_vp]),
    "iq_model_workspace_bytes": (_i64, [_vp, _i64, _i64]),
    "iq_model_forward": (_int, [_vp, _vp, _int, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "iq_model_forward_coalitions": (_int, [_vp, _vp, _int, _i64, _i64, _vp, _vp, _vp, _i64, _vp]),
    "iq_collapse_plan": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "iq_model_last_row_fraction": (ctypes.c_double, [_vp]),
    "iq_model_last_buckets": (_int, [_vp, ctypes.POINTER(_i64), _int]),
    "iq_ball_query": (_int, [_vp, _vp, _i64, _i64, _i64, ctypes.c_double, _int, _vp, _vp]),
    "iq_knn_xyz": (_int, [_vp, _i64, _i64, _int, _vp, _vp]),
    "iq_knn_features": (_int, [_vp, _i64, _i64, _i64, _int, _vp, _vp, _vp]),
    "iq_region_smoothness_epoch": (_int, [_vp] * 12 + [_i64, _i64, _i64, _int, _int, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_double, ctypes.c_double, _int, _int, _vp]),
    "iq_topk_rows": (_int, [_vp, _i64, _i64, _i64, _int, _int, _vp, _vp]),
    "iq_linear": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp]),
    "iq_linear_pool": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _vp, _vp, _vp]),
    "iq_grouped_mlp_max": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "iq_model_set_engine": (_int, [_vp, _int]),
}

_lib = None


class IQError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IQError("libiq_b200.so is not built (run `python interpret_quality_b200/build.py` or "
                      "__graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise IQError(load().iq_last_error().decode())


def last_error():
    return load().iq_last_error().decode()


def launch_count():
    return int(load().iq_launch_count())


def f16_paths():
    """Bit mask of the DGCNN / GCNN products running on kind::f16 MMAs: 1 conv5, 2 EdgeConv products, 4 Gram kNN."""
    return int(load().iq_f16_paths())


def profile_enable(on=True):
    load().iq_profile_enable(1 if on else 0)


def profile_report(cap=64):
    """{kernel name: (total ms, launches)} since profile_enable(True)."""
    names = (ctypes.c_char_p * cap)()
    ms = (ctypes.c_double * cap)()
    counts = (ctypes.c_longlong * cap)()
    n = min(load().iq_profile_report(names, ms, counts, cap), cap)
    return {names[i].decode(): (ms[i], int(counts[i])) for i in range(n)}
