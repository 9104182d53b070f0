"""Consumers of the coalition path's files: final_result's Table 2-4 code against the
reference's own functions run on the same value arrays (tests/golden/result_tables.npz, make_golden.py result_tables)."""
import os

import numpy as np
import pytest

from interpret_quality_b200 import final_result as res
from interpret_quality_b200 import synthetic

NAMES = ["cloud%d" % i for i in range(4)]
TOL = 1e-12        # float64 numpy on both sides; only the summation order of the neighbour mean differs


def lay_out(g, root):
    """Write the golden value arrays in the reference's folder layout (tools/final_common.py:150-172)."""
    for name in NAMES:
        base = root + name + "/"
        for mode in ("scale", "rotate"):
            os.makedirs(base + "%s_all/" % mode)
            np.save(base + "%s_all/region_shapley_value.npy" % mode, g["%s_%s_values" % (name, mode)])
        for direction in ("inc", "dec"):
            os.makedirs(base + "linearity_all/allregion_%s/" % direction)
            np.save(base + "linearity_all/allregion_%s/region_shapley_value.npy" % direction,
                    g["%s_linearity_%s_values" % (name, direction)])
        np.save(base + "region_id.npy", g["region_id"])


def close(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= TOL * max(np.abs(b).max(), 1.0)


def test_sensitivity_matches_the_reference(golden, tmp_path):
    g, root = golden("result_tables"), str(tmp_path) + "/"
    lay_out(g, root)
    for name in NAMES:
        for mode in ("scale", "rotate", "linearity"):               # linearity: inc and dec runs concatenated
            got = res.cal_sensitivity(root + name + "/", mode)
            assert got.shape == (32,) and close(got, g["%s_%s_sensitivity" % (name, mode)])
    for mode in ("scale", "rotate"):
        assert close(res.cal_sensitivity_all_pc(root, NAMES, mode), g["sens_all_%s" % mode])
        assert close(res.cal_mean_sv_intensity(root, NAMES, mode), g["intensity_all_%s" % mode])


def test_pearson_table_matches_the_reference(golden, tmp_path, capsys):
    g, root = golden("result_tables"), str(tmp_path) + "/"
    lay_out(g, root)
    for mode in ("scale", "rotate"):
        mean_r, all_r = res.cal_correlation_coef(root, NAMES, mode)
        assert all_r.shape == (4,) and abs(mean_r - float(g["pearson_mean_%s" % mode])) <= 1e-12
    assert "mean Pearson r=" in capsys.readouterr().out
    with pytest.raises(AssertionError):
        res.cal_correlation_coef(root, NAMES, "linearity")           # pose modes only, like the reference


def test_pearson_r_known_answers():
    x = np.arange(10.0)
    assert res.pearson_r(x, 3 * x + 1) == pytest.approx(1.0, abs=1e-15)
    assert res.pearson_r(x, -x) == pytest.approx(-1.0, abs=1e-15)
    assert abs(res.pearson_r([1, 2, 3, 4], [1, -1, -1, 1])) < 1e-15


def test_smoothness_matches_the_reference(golden, tmp_path):
    g, root = golden("result_tables"), str(tmp_path) + "/"
    lay_out(g, root)
    data = synthetic.make_cloud(1024)[0]
    for name in NAMES:
        m, m_poses, den = res.cal_shapley_smoothness_metric_single_pc(data, g["%s_rotate_values" % name], g["region_id"])
        assert close([m, den], g["%s_smooth" % name]) and close(m_poses, g["%s_smooth_poses" % name])
    table = res.cal_shapley_smoothness_metric(root, [(data[None], n) for n in NAMES] + [(data[None], "Knife_x")], "rotate",
                                              verbose=False)
    assert table.shape == (4,) and close(table, [g["%s_smooth" % n][0] for n in NAMES])   # Knife skipped


def test_ball_query_contains_self_and_is_symmetric():
    x = np.random.RandomState(0).rand(32, 3)
    nb = res.ball_query(x, 0.4)
    assert nb.dtype == bool and nb.diagonal().all() and np.array_equal(nb, nb.T)
    d = np.linalg.norm(x[:, None] - x[None], axis=-1)
    off = ~np.eye(32, dtype=bool)
    assert np.array_equal(nb[off], (d < 0.4)[off])
