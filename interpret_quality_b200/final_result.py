"""Consumers of the files the coalition path writes (final_result.py:66-176 of ada-shen/Interpret_quality):
Table 2 sensitivity, Table 3 Pearson correlation between sensitivity and mean |Shapley value|, Table 4 spatial
non-smoothness.  Host-side float64 numpy like the reference (a few kB per cloud, no device work); this module
exists so that the artefacts of tools.final_common.test() / final_shapley_value.shap_sampling() can be checked
end to end against the reference's own table code.

The reference reads model / dataset / folder names from a global argparse namespace and its dataset loaders;
here the experiment folder and the per-cloud folder names are explicit arguments.  Plotting is out of scope."""
import numpy as np

BALL_QUERY_COEF = 0.25          # tools/final_util.py:68
GEOMETRY_MODES = ("linearity", "planarity", "scattering")
POSE_MODES = ("trans", "rotate", "scale")


def square_distance_np(x):
    """(N,F) -> (N,N) squared distances in the reference's expanded form (tools/final_util.py:122-132)."""
    x = np.asarray(x)
    sq = np.sum(x ** 2, axis=1, keepdims=True)
    return sq + sq.T - 2 * np.matmul(x, x.T)


def ball_query(x, r):
    """(R,d) centres -> (R,R) bool, True where the squared distance is < r^2; a centre is its own neighbour
    (tools/final_util.py:150-160)."""
    return square_distance_np(x) < r ** 2


def load_region_shapley_values(base_folder, mode):
    """(poses, R) float64 Shapley values of one cloud under `mode`.  Geometry modes concatenate the ascent and the
    descent runs (final_result.py:87-90), pose modes read <mode>_all/region_shapley_value.npy (:92)."""
    if any(g in mode for g in GEOMETRY_MODES):
        parts = [np.load(base_folder + "%s_all/allregion_%s/region_shapley_value.npy" % (mode, d)) for d in ("inc", "dec")]
        return np.concatenate(parts, axis=0)
    return np.load(base_folder + "%s_all/region_shapley_value.npy" % mode)


def sensitivity_of(region_shapley_values):
    """(poses,R) -> (R,): range of each region's value over the poses, normalised by the mean L1 norm of a pose's
    values (final_result.py:94-102)."""
    v = np.asarray(region_shapley_values)
    return (v.max(axis=0) - v.min(axis=0)) / np.abs(v).sum(axis=1).mean()


def cal_sensitivity(base_folder, mode):
    """final_result.py:83-102."""
    return sensitivity_of(load_region_shapley_values(base_folder, mode))


def cal_sensitivity_all_pc(exp_folder, folder_name_list, mode):
    """(num_pc, R) sensitivities of every cloud folder (final_result.py:106-121)."""
    return np.array([cal_sensitivity(exp_folder + "%s/" % name, mode) for name in folder_name_list])


def cal_mean_sv_intensity(exp_folder, folder_name_list, mode):
    """(num_pc, R) mean |Shapley value| over the poses (final_result.py:61-79); pose modes only."""
    assert mode in POSE_MODES
    return np.array([np.abs(np.load(exp_folder + "%s/%s_all/region_shapley_value.npy" % (name, mode))).mean(axis=0)
                     for name in folder_name_list])


def pearson_r(x, y):
    """Pearson correlation of two 1-d arrays, float64 (what scipy.stats.pearsonr returns first, :136)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xm, ym = x - x.mean(), y - y.mean()
    nx, ny = np.linalg.norm(xm), np.linalg.norm(ym)
    return float(np.clip(np.dot(xm / nx, ym / ny), -1.0, 1.0))


def cal_correlation_coef(exp_folder, folder_name_list, mode, verbose=True):
    """Mean over the clouds of the Pearson r between a cloud's region sensitivities and its mean |Shapley value|s
    (final_result.py:124-140).  Returns (mean r, (num_pc,) r's); the reference returns the mean and prints the
    mean and the ddof=1 standard deviation."""
    assert mode in POSE_MODES
    sens = cal_sensitivity_all_pc(exp_folder, folder_name_list, mode)
    inten = cal_mean_sv_intensity(exp_folder, folder_name_list, mode)
    r = np.array([pearson_r(s, m) for s, m in zip(sens, inten)])
    if verbose:
        std = r.std(ddof=1) if r.size > 1 else float("nan")
        print("mean Pearson r=%f±%f" % (r.mean(), std))
    return r.mean(), r


def cal_shapley_smoothness_metric_single_pc(data, region_shapley_values, region_id, num_regions=None):
    """Spatial non-smoothness of one cloud (final_result.py:144-176).

    data (N,3), region_shapley_values (poses,R), region_id (N,).  Neighbours of region i are the regions whose
    centroid lies within BALL_QUERY_COEF * cloud diameter of i's centroid (i included).  For every pose and region
    the mean |phi_i - phi_j| over the neighbours j, divided by the mean |sum_i phi_i| over the poses.
    Returns (metric, metric_all_poses (poses,), denominator)."""
    data = np.asarray(data)
    v = np.asarray(region_shapley_values, dtype=np.float64)
    region_id = np.asarray(region_id)
    R = int(num_regions) if num_regions is not None else v.shape[1]
    centers = np.stack([data[region_id == i].mean(axis=0) for i in range(R)]).astype(np.float64)
    diameter = np.sqrt(np.maximum(square_distance_np(data), 0)).max()
    neighbor = ball_query(centers, r=BALL_QUERY_COEF * diameter)                     # (R,R) bool
    denominator = np.abs(v.sum(axis=1)).mean()
    gap = np.abs(v[:, :, None] - v[:, None, :])                                      # (poses,R,R)
    fraction = (gap * neighbor[None]).sum(axis=2) / neighbor.sum(axis=1)[None] / denominator
    return fraction.mean(), fraction.mean(axis=1), denominator


def cal_shapley_smoothness_metric(exp_folder, samples, mode, verbose=True):
    """Table 4 over clouds (final_result.py:179-214).  samples: iterable of (data (N,3) array-like, folder_name);
    the reference takes them from its dataset loader and skips ShapeNet's Knife class (:201).
    Returns the (num_pc,) metrics."""
    assert mode in ("trans", "rotate")
    out = []
    for data, name in samples:
        if name[:5] == "Knife":
            continue
        base_folder = exp_folder + "%s/" % name
        data = np.asarray(data.cpu() if hasattr(data, "cpu") else data).reshape(-1, 3)
        region_id = np.load(base_folder + "region_id.npy")
        v = np.load(base_folder + "%s_all/region_shapley_value.npy" % mode)
        metric, _, _ = cal_shapley_smoothness_metric_single_pc(data, v, region_id)
        if verbose:
            print("%s, metric=%f" % (name, metric))
        out.append(metric)
    return np.array(out)
