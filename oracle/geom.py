"""ORACLE (test infrastructure): ctypes front-end of geom_oracle.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libiq_oracle.so")
_lib = None


def build(force=False):
    """gcc-compile geom_oracle.c into oracle/_build/ (idempotent)."""
    src = os.path.join(_HERE, "geom_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libiq_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_I = ctypes.c_int64


def fps(xyz, npoint):
    """final_save_fps.py:10-31 -> (B,npoint) int64."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), np.int64)
    lib().oracle_fps(_p(xyz), _I(B), _I(N), _I(npoint), _p(out))
    return out


def square_distance3(src, dst):
    """tools/final_util.py:134-147 for C=3 -> (B,N,M) float32."""
    src, dst = _f32(src), _f32(dst)
    B, N, _ = src.shape
    M = dst.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().oracle_square_distance3(_p(src), _p(dst), _I(B), _I(N), _I(M), _p(out))
    return out


def region_id(xyz, fps_index):
    """final_shapley_value.py:20-35 -> (N,) int64."""
    xyz = _f32(xyz).reshape(-1, 3)
    fps_index = _i64(fps_index)
    out = np.empty((xyz.shape[0],), np.int64)
    lib().oracle_region_id(_p(xyz), _p(fps_index), _I(xyz.shape[0]), _I(fps_index.shape[0]), _p(out))
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    """models/pointnet2.py:70-91 -> (B,S,nsample) int64."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), np.int64)
    lib().oracle_ball_query(_p(xyz), _p(new_xyz), _I(B), _I(N), _I(S), _I(nsample), ctypes.c_double(radius), _p(out))
    return out


def mask_shapley(data, center, orders, region_ids):
    """tools/final_common.py:46-61 -> ((R+1)*bs, N, 3) float32."""
    data = _f32(data).reshape(-1, 3)
    center = _f32(center).reshape(3)
    orders = _i64(orders)
    region_ids = _i64(region_ids)
    bs, R = orders.shape
    N = data.shape[0]
    out = np.empty(((R + 1) * bs, N, 3), np.float32)
    lib().oracle_mask_shapley(_p(data), _p(center), _p(orders), _p(region_ids), _I(bs), _I(R), _I(N), _p(out))
    return out


def mask_interaction(data, center, contexts, region_i, region_j, region_ids, num_regions):
    """final_point_binary_interaction_logits.py:42-56 -> (4*ctx, 3, N) float32."""
    data = _f32(data).reshape(-1, 3)
    center = _f32(center).reshape(3)
    contexts = _i64(contexts)
    ctx, m = contexts.shape
    region_ids = _i64(region_ids)
    N = data.shape[0]
    out = np.empty((4 * ctx, 3, N), np.float32)
    lib().oracle_mask_interaction(_p(data), _p(center), _p(contexts), _I(ctx), _I(m), _I(int(region_i)),
                                  _I(int(region_j)), _p(region_ids), _I(num_regions), _I(N), _p(out))
    return out
