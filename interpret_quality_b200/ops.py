"""Tensor-level wrappers over the C ABI: argument checking + pointer marshalling.

torch is used for device memory and streams only; all arithmetic happens in
libiq_b200.so.  Every function requires CUDA tensors and raises otherwise.
"""
import numpy as np
import torch

from . import _lib


def _call(t, name, *args):
    """Run one library entry on the device (and that device's current stream) of tensor `t`: the library launches on
    the CUDA device that is current, so the guard keeps models / inputs on cuda:1 off cuda:0's stream."""
    with torch.cuda.device(t.device):
        _lib.check(getattr(_lib.load(), name)(*args, torch.cuda.current_stream(t.device).cuda_stream))


def _chk(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise _lib.IQError("%s must live on a CUDA device (iq_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def to_dev_i64(a, device):
    """numpy / tensor integer array -> contiguous int64 CUDA tensor."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.int64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.int64)).to(device)


def fps(xyz, npoint):
    _chk(xyz, torch.float32, "xyz")
    B, N, C = xyz.shape
    if C != 3:
        raise ValueError("xyz must be (B,N,3)")
    out = torch.empty((B, npoint), dtype=torch.int64, device=xyz.device)
    _call(xyz, "iq_fps", xyz.data_ptr(), B, N, npoint, out.data_ptr())
    return out


def square_distance3(src, dst):
    _chk(src, torch.float32, "src")
    _chk(dst, torch.float32, "dst")
    B, N, _ = src.shape
    M = dst.shape[1]
    out = torch.empty((B, N, M), dtype=torch.float32, device=src.device)
    _call(src, "iq_square_distance3", src.data_ptr(), dst.data_ptr(), B, N, M, out.data_ptr())
    return out


def region_id(xyz, fps_index):
    _chk(xyz, torch.float32, "xyz")
    _chk(fps_index, torch.int64, "fps_index")
    N = xyz.shape[-2]
    out = torch.empty((N,), dtype=torch.int64, device=xyz.device)
    _call(xyz, "iq_region_id", xyz.data_ptr(), fps_index.data_ptr(), N, fps_index.numel(), out.data_ptr())
    return out


def center(xyz):
    _chk(xyz, torch.float32, "xyz")
    N = xyz.shape[-2]
    out = torch.empty((3,), dtype=torch.float32, device=xyz.device)
    _call(xyz, "iq_center", xyz.data_ptr(), N, out.data_ptr())
    return out


def mask_shapley(data, center_t, orders, region_ids, out=None, in_place=False):
    """data (N,3) | None, center (3), orders (bs,R) i64, region_ids (N) i64 -> ((R+1)*bs, N, 3)."""
    _chk(center_t, torch.float32, "center")
    _chk(orders, torch.int64, "orders")
    _chk(region_ids, torch.int64, "region_id")
    bs, R = orders.shape
    N = region_ids.numel()
    if in_place:
        _chk(out, torch.float32, "masked_data")
        if out.numel() != (R + 1) * bs * N * 3:
            raise ValueError("masked_data must be ((num_regions+1)*bs, num_points, 3)")
        dptr = 0
    else:
        _chk(data, torch.float32, "data")
        if out is None:
            out = torch.empty(((R + 1) * bs, N, 3), dtype=torch.float32, device=data.device)
        dptr = data.data_ptr()
    _call(out, "iq_mask_shapley", dptr, center_t.data_ptr(), orders.data_ptr(), region_ids.data_ptr(), bs, R,
                                           N, out.data_ptr(), 1 if in_place else 0)
    return out


def mask_interaction(data, center_t, contexts, region_i, region_j, region_ids, num_regions, point_major=False,
                     out=None):
    """data (N,3), contexts (ctx,m) i64 -> (4*ctx,3,N) (reference layout) or (4*ctx,N,3)."""
    _chk(data, torch.float32, "data")
    _chk(center_t, torch.float32, "center")
    _chk(contexts, torch.int64, "contexts")
    _chk(region_ids, torch.int64, "region_id")
    ctx, m = contexts.shape
    N = region_ids.numel()
    if out is None:
        shape = (4 * ctx, N, 3) if point_major else (4 * ctx, 3, N)
        out = torch.empty(shape, dtype=torch.float32, device=data.device)
    cptr = contexts.data_ptr() if contexts.numel() else 0
    _call(data, "iq_mask_interaction", data.data_ptr(), center_t.data_ptr(), cptr, ctx, m, int(region_i),
                                               int(region_j), region_ids.data_ptr(), num_regions, N,
                                               1 if point_major else 0, out.data_ptr())
    return out


def mask_interaction_pairs(data, center_t, pairs, contexts, region_ids, num_regions, point_major=True, out=None):
    """data (N,3), pairs (P,2) i64, contexts (P,ctx,m) i64 -> (P*ctx*4, N, 3) (or (P*ctx*4, 3, N)): the
    4-clouds-per-context blocks of every pair in one launch."""
    _chk(data, torch.float32, "data")
    _chk(center_t, torch.float32, "center")
    _chk(pairs, torch.int64, "pairs")
    _chk(contexts, torch.int64, "contexts")
    _chk(region_ids, torch.int64, "region_id")
    P, ctx, m = contexts.shape
    if pairs.shape != (P, 2):
        raise ValueError("pairs must be (P,2) with P = contexts.shape[0]")
    N = region_ids.numel()
    shape = (4 * P * ctx, N, 3) if point_major else (4 * P * ctx, 3, N)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=data.device)
    elif out.numel() != 4 * P * ctx * N * 3:
        raise ValueError("out must hold %s floats" % (shape,))
    cptr = contexts.data_ptr() if contexts.numel() else 0
    _call(data, "iq_mask_interaction_pairs", data.data_ptr(), center_t.data_ptr(), pairs.data_ptr(), cptr, P, ctx, m,
          region_ids.data_ptr(), num_regions, N, 1 if point_major else 0, out.data_ptr())
    return out


def reward(logits, lbl, softmax_type="modified", out=None):
    _chk(logits, torch.float32, "logits")
    B, C = logits.shape
    if out is None:
        out = torch.empty((B,), dtype=torch.float32, device=logits.device)
    _call(logits, "iq_reward", logits.data_ptr(), B, C, int(lbl), 1 if softmax_type == "normal" else 0,
                                     out.data_ptr())
    return out


def shapley_accumulate(v, orders, phi_sum):
    _chk(v, torch.float32, "v")
    _chk(orders, torch.int64, "orders")
    _chk(phi_sum, torch.float64, "phi_sum")
    bs, R = orders.shape
    if v.numel() != bs * (R + 1) or phi_sum.numel() != R:
        raise ValueError("shape mismatch in shapley_accumulate")
    _call(v, "iq_shapley_accumulate", v.data_ptr(), orders.data_ptr(), bs, R, phi_sum.data_ptr())
    return phi_sum


def interaction_reduce(all_logits, lbl, softmax_type="modified"):
    _chk(all_logits, torch.float32, "all_logits")
    P, rows, C = all_logits.shape
    ctx = rows // 4
    out = torch.empty((P, ctx), dtype=torch.float64, device=all_logits.device)
    _call(all_logits, "iq_interaction_reduce", all_logits.data_ptr(), P, ctx, C, int(lbl),
                                                 1 if softmax_type == "normal" else 0, out.data_ptr())
    return out


def knn_xyz(xyz, k):
    """xyz (B,N,3) point-major -> (B,N,k) int32 neighbour sets (models/dgcnn.py:12-18 for 3-d input)."""
    _chk(xyz, torch.float32, "xyz")
    B, N, _ = xyz.shape
    out = torch.empty((B, N, k), dtype=torch.int32, device=xyz.device)
    _call(xyz, "iq_knn_xyz", xyz.data_ptr(), B, N, k, out.data_ptr())
    return out


def knn_features(x, k, return_counts=False):
    """x (B,N,C) point-major features, C in {64,128} -> (B,N,k) int32: the SET of the k nearest points by (distance, index), in no particular order:
    the fused tcgen05 Gram + candidate selection + exact re-rank of csrc/knn_tc.cu (models/dgcnn.py:12-18)."""
    _chk(x, torch.float32, "x")
    B, N, C = x.shape
    out = torch.empty((B, N, k), dtype=torch.int32, device=x.device)
    cnt = torch.empty((B, N), dtype=torch.int32, device=x.device) if return_counts else None
    _call(x, "iq_knn_features", x.data_ptr(), B, N, C, k, out.data_ptr(), cnt.data_ptr() if return_counts else None)
    return (out, cnt) if return_counts else out


SMOOTHNESS_MODES = {"linearity": 0, "planarity": 1, "scattering": 2}


def region_smoothness_epoch(data, data_orig, offsets, members, orient, var_ub, var_lb, smoothness, alive, max_region, mode,
                            objective, step, enum_step, dist_threshold, stop_ratio, max_iteration, clamp=False):
    """One epoch of the geometry ascent / descent over every region whose `alive` flag is set
    (final_smoothness_center_enum_all.py:184-243, :305-321), csrc/smoothness.cu.  data (N,3) is updated in place,
    smoothness (R) float64 and alive (R) int32 too.  Returns (iters (R) int32, last_var (R,3) float32,
    stop_flags (R) int32), all on the device."""
    _chk(data, torch.float32, "data")
    _chk(data_orig, torch.float32, "data_orig")
    _chk(offsets, torch.int32, "offsets")
    _chk(members, torch.int32, "members")
    _chk(orient, torch.float32, "orient")
    _chk(var_ub, torch.float32, "var_ub")
    _chk(var_lb, torch.float32, "var_lb")
    _chk(smoothness, torch.float64, "smoothness")
    _chk(alive, torch.int32, "alive")
    N, R = data.shape[0], offsets.shape[0] - 1
    if data.shape != (N, 3) or data_orig.shape != (N, 3) or members.shape[0] != N:
        raise ValueError("region_smoothness_epoch: data / data_orig must be (N,3) and members (N,)")
    if orient.shape != (R, 3, 3) or var_ub.shape != (R, 3) or var_lb.shape != (R, 3) or smoothness.shape != (R,) \
            or alive.shape != (R,):
        raise ValueError("region_smoothness_epoch: per-region arrays must have %d rows" % R)
    iters = torch.empty((R,), dtype=torch.int32, device=data.device)
    last_var = torch.zeros((R, 3), dtype=torch.float32, device=data.device)
    flags = torch.zeros((R,), dtype=torch.int32, device=data.device)
    _call(data, "iq_region_smoothness_epoch", 
        data.data_ptr(), data_orig.data_ptr(), offsets.data_ptr(), members.data_ptr(), orient.data_ptr(), var_ub.data_ptr(),
        var_lb.data_ptr(), smoothness.data_ptr(), alive.data_ptr(), iters.data_ptr(), last_var.data_ptr(), flags.data_ptr(),
        N, R, int(max_region), SMOOTHNESS_MODES[mode], 1 if objective == "inc" else 0, float(step), float(enum_step),
        float(dist_threshold), float(stop_ratio), int(max_iteration), 1 if clamp else 0)
    return iters, last_var, flags


def topk_rows(keys, k, largest=True):
    """keys (rows,N) -> (rows,k) int32, unordered exact top-k with lowest-index tie-breaking."""
    _chk(keys, torch.float32, "keys")
    rows, N = keys.shape
    out = torch.empty((rows, k), dtype=torch.int32, device=keys.device)
    _call(keys, "iq_topk_rows", keys.data_ptr(), rows, N, N, k, 1 if largest else 0, out.data_ptr())
    return out


def linear(x, w, b=None, act=0, engine=0):
    """act(x @ w.T + b) through the library's GEMM (engine 0: exact fp32 SIMT, 1: tcgen05 3xTF32, 2: tcgen05 kind::f16 on
    two-term fp16 splits)."""
    _chk(x, torch.float32, "x")
    _chk(w, torch.float32, "w")
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    bp = _chk(b, torch.float32, "b").data_ptr() if b is not None else 0
    _call(x, "iq_linear", x.data_ptr(), w.data_ptr(), bp, M, N, K, act, engine, y.data_ptr())
    return y


def linear_pool(x, w, b, clouds, points, act=0, engine=0, want_mean=True, want_arg=False):
    """Pooled 1x1 conv: x (clouds*points, K), w (N,K) -> (max (clouds,N), mean | None, argmax | None)."""
    _chk(x, torch.float32, "x")
    _chk(w, torch.float32, "w")
    K = x.shape[1]
    N = w.shape[0]
    mx = torch.empty((clouds, N), dtype=torch.float32, device=x.device)
    mean = torch.empty((clouds, N), dtype=torch.float32, device=x.device) if want_mean else None
    arg = torch.empty((clouds, N), dtype=torch.int64, device=x.device) if want_arg else None
    bp = _chk(b, torch.float32, "b").data_ptr() if b is not None else 0
    _call(x, "iq_linear_pool", x.data_ptr(), w.data_ptr(), bp, clouds, points, N, K, act, engine,
                                          mx.data_ptr(), mean.data_ptr() if want_mean else 0,
                                          arg.data_ptr() if want_arg else 0)
    return mx, mean, arg


def ball_query(radius, nsample, xyz, new_xyz):
    """xyz (B,N,3), new_xyz (B,S,3) -> (B,S,nsample) int32 (models/pointnet2.py:70-91)."""
    _chk(xyz, torch.float32, "xyz")
    _chk(new_xyz, torch.float32, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty((B, S, nsample), dtype=torch.int32, device=xyz.device)
    _call(xyz, "iq_ball_query", xyz.data_ptr(), new_xyz.data_ptr(), B, N, S, float(radius), nsample,
                                         out.data_ptr())
    return out


def grouped_mlp_max(U, V, b1, idx, clouds, S, K, W2, b2, W3, b3):
    """max over each group of K rows of relu(relu(relu(U[idx] - V + b1) W2^T + b2) W3^T + b3): csrc/chain_tc.cu
    (one PointNet++ set-abstraction scale, models/pointnet2.py:215-232).  U (clouds*nsrc, C1), V (clouds*S, C1),
    idx (clouds*S*K) int32 -> (clouds*S, C3)."""
    for t, n in ((U, "U"), (V, "V"), (b1, "b1"), (W2, "W2"), (b2, "b2"), (W3, "W3"), (b3, "b3")):
        _chk(t, torch.float32, n)
    _chk(idx, torch.int32, "idx")
    C1, C2, C3 = U.shape[1], W2.shape[0], W3.shape[0]
    nsrc = U.shape[0] // clouds
    out = torch.empty((clouds * S, C3), dtype=torch.float32, device=U.device)
    _call(U, "iq_grouped_mlp_max", U.data_ptr(), V.data_ptr(), b1.data_ptr(), idx.data_ptr(), clouds, S, K, nsrc, C1,
          W2.data_ptr(), b2.data_ptr(), C2, W3.data_ptr(), b3.data_ptr(), C3, out.data_ptr())
    return out
