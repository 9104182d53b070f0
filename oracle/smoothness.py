"""ORACLE (test infrastructure, not product code).

CPU restatement of the geometry-ascent loop of final_smoothness_center_enum_all.py (ada-shen/Interpret_quality):
every region's points are pushed by normalised gradient steps so that the region's linearity / planarity /
scattering (ratios of the variances along its three original principal orientations) rises or falls by
ENUM_STEP per epoch, under a variance bound and a per-point distance bound.  torch fp32 on the CPU with
autograd, like the reference; torch.symeig (:41, removed from torch) is torch.linalg.eigh here.

Pinned to tests/golden/smoothness.npz (make_golden.py smoothness: the unmodified reference, CPU).
"""
import numpy as np
import torch

# final_smoothness_center_enum_all.py:13-19
STEP, ENUM_STEP, EPOCH = 1e-3, 0.05, 50
VAR_THRESHOLD, DIST_THRESHOLD, STOP_RATIO, MAX_ITERATION = 0.003, 0.03, 0.5, 100
# clamp=False is the reference's observable behaviour: apply_distance_bound :103-120 assigns to the .data of a
# temporary row view (`data_region_i[i].data = ...`), which leaves data_region_i untouched, so points are only
# COUNTED against the distance bound, never pulled back.  clamp=True is what its docstring describes.
HP = dict(step=STEP, enum_step=ENUM_STEP, var_threshold=VAR_THRESHOLD, dist_threshold=DIST_THRESHOLD,
          stop_ratio=STOP_RATIO, max_iteration=MAX_ITERATION, clamp=False)


def principal_orientations(pts):
    """cal_principal_orientation :22-45.  pts (S,3) fp32 tensor -> (3,3) tensor, rows o1,o2,o3 = unit eigenvectors of
    the (S-1)-normalised covariance for the largest, middle, smallest eigenvalue."""
    c = pts - pts.mean(dim=0)
    cov = (c.unsqueeze(2) * c.unsqueeze(1)).sum(dim=0) / (pts.shape[0] - 1)
    _, vec = torch.linalg.eigh(cov)                    # ascending eigenvalues, eigenvectors in columns
    return torch.stack([vec[:, 2], vec[:, 1], vec[:, 0]]).detach().clone()


def variances(pts, orient, two_pass=False):
    """cal_variance :48-62: unbiased variance of the projections on o1,o2,o3 -> three scalar tensors.
    two_pass=True is the same quantity with a different fp32 rounding (explicit mean, then squared deviations);
    well_conditioned_regions() uses it to find the regions whose trajectory does not depend on the last bit."""
    if not two_pass:
        return [torch.var(torch.matmul(pts, orient[k])) for k in range(3)]
    out = []
    for k in range(3):
        p = (pts * orient[k]).sum(dim=1)
        d = p - p.mean()
        out.append((d * d).sum() / (pts.shape[0] - 1))
    return out


def smoothness_of(v, mode):
    """:205-221 / :143-159: the mode's ratio of the sorted variances (np.argsort order, :85-100)."""
    order = np.argsort(np.array([x.item() for x in v])).tolist()
    s_min, s_mid, s_max = v[order[0]], v[order[1]], v[order[2]]
    if mode == "linearity":
        return (s_max - s_mid) / s_max, (s_max, s_mid)
    if mode == "planarity":
        return (s_mid - s_min) / s_max, (s_max, s_mid, s_min)
    return s_min / s_max, (s_max, s_min)


def region_info(pts, mode, hp=HP):
    """get_original_region_info :245-268 -> (orient (3,3), ub (3,), lb (3,), smoothness float)."""
    orient = principal_orientations(pts)
    v = variances(pts, orient)
    ub = torch.stack([x + hp["var_threshold"] for x in v])
    lb = torch.stack([x - hp["var_threshold"] for x in v])
    with torch.no_grad():
        s, _ = smoothness_of(v, mode)
    return orient, ub, lb, s.item()


def update_region(pts, pts_orig, orient, ub, lb, smooth0, mode, objective, hp=HP):
    """update_region :184-243 on one region's points (S,3); returns (new pts, smoothness, if_update, iterations,
    last variances).  `smoothness` is the value measured BEFORE the last step, as the reference reports it."""
    rising = objective == "inc"
    target = smooth0 + hp["enum_step"] if rising else smooth0 - hp["enum_step"]
    smooth, if_update, it = smooth0, True, 0
    pts = pts.clone()
    last_v = None
    while (smooth < target) if rising else (smooth > target):
        x = pts.clone().detach().requires_grad_(True)
        v = variances(x, orient, hp.get("two_pass", False))
        v = [vk.detach() if (vk > ub[k] or vk < lb[k]) else vk for k, vk in enumerate(v)]      # :65-73
        f, used = smoothness_of(v, mode)
        smooth = f.item()
        if any(u.requires_grad for u in used):
            f.backward()
        grad_none = x.grad is None                                                               # :123-139
        if not grad_none:
            g = x.grad
            n = torch.norm(g)
            delta = hp["step"] * g / n if n != 0 else 1e-8
            x = (x.detach() + delta) if rising else (x.detach() - delta)
        else:
            x = x.detach()
        diff = x - pts_orig                                                                      # :103-120
        dist = torch.norm(diff, dim=1)
        far = dist > hp["dist_threshold"]
        count = int(far.sum())
        if hp.get("clamp"):
            x = torch.where(far[:, None], pts_orig + hp["dist_threshold"] * diff / dist[:, None], x)
        pts = x
        last_v = [vk.item() for vk in v]
        it += 1
        if count / pts.shape[0] > hp["stop_ratio"] or grad_none or it > hp["max_iteration"]:     # :167-181
            if_update = False
            break
    return pts, smooth, if_update, it, last_v


def run_epochs(data, region_id, num_regions, mode, objective, epochs, hp=HP):
    """The geometry part of test_all_region :281-350 (no classifier): returns (clouds (E,1,N,3) float32,
    smoothness (E,R) float64, iterations (E,R) int)."""
    data = torch.as_tensor(data, dtype=torch.float32).reshape(1, -1, 3)
    region_id = np.asarray(region_id)
    sel = [torch.from_numpy(np.nonzero(region_id == r)[0]) for r in range(num_regions)]
    orig = [data[0, s].clone() for s in sel]
    info = [region_info(o, mode, hp) for o in orig]
    smooth = [i[3] for i in info]
    alive = [True] * num_regions
    cur = data.clone()
    clouds, smooths, iters = [], [], []
    for _ in range(epochs):
        row_it = [0] * num_regions
        for r in range(num_regions):
            if alive[r]:
                pts, smooth[r], alive[r], row_it[r], _ = update_region(cur[0, sel[r]], orig[r], info[r][0], info[r][1],
                                                                       info[r][2], smooth[r], mode, objective, hp)
                cur[0, sel[r]] = pts
        clouds.append(cur.numpy().copy())
        smooths.append(list(smooth))
        iters.append(row_it)
        if not any(alive):
            break
    return np.stack(clouds), np.array(smooths), np.array(iters)


def well_conditioned_regions(data, region_id, num_regions, mode, objective, epochs, hp=HP, tol=1e-4):
    """(R,) bool.  The loop is a threshold process: near a tie of two variances, or where the smoothness is ~0, the
    choice of which variance a step pushes flips on the last bit and trajectories separate by up to STEP per step --
    in the reference itself (CPU vs CUDA, or two torch builds).  A region is well conditioned when re-running the
    oracle with the two-pass variance moves none of its points by more than `tol` in any epoch; implementations
    are compared point by point on those regions and by invariants on the others."""
    a, _, _ = run_epochs(data, region_id, num_regions, mode, objective, epochs, hp)
    b, _, _ = run_epochs(data, region_id, num_regions, mode, objective, epochs, dict(hp, two_pass=True))
    n = min(a.shape[0], b.shape[0])
    per_point = np.abs(a[:n] - b[:n]).max(axis=(0, 1, 3))
    region_id = np.asarray(region_id)
    ok = np.array([per_point[region_id == r].max() <= tol for r in range(num_regions)])
    return ok if a.shape[0] == b.shape[0] else np.zeros(num_regions, dtype=bool)
