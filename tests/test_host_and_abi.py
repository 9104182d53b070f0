"""CPU-only checks of the host logic and of the C-ABI shared library: it loads, exports every symbol
include/iq_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import _lib, build, synthetic
from interpret_quality_b200.config import CONFIG
from interpret_quality_b200.tools import final_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "iq_b200.h")).read()
    declared = set(re.findall(r"\b(iq_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), "libiq_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header drifted apart"
    assert lib.iq_version() >= 100


def test_sass_contains_tcgen05_and_tma(lib):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass                            # TMA tensor loads
    assert "LDTM" in sass                               # tcgen05.ld


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    args = types.SimpleNamespace(model="dgcnn", k=20, dataset="shapenet", device="cpu")
    model = final_util.build_model(args, synthetic.make_state_dict("dgcnn"))
    with pytest.raises(_lib.IQError):
        model(torch.zeros(1, 3, 1024))
    from interpret_quality_b200 import ops
    with pytest.raises(_lib.IQError):
        ops.fps(torch.zeros(1, 64, 3), 8)


def test_batch_knobs_match_reference_config():
    want = {"shapley_batch_size": {"pointnet2": 5, "pointnet": 50, "dgcnn": 5, "gcnn": 10, "pointconv": 20},
            "interaction_batch_size": {"pointnet2": 25, "pointnet": 100, "dgcnn": 25, "gcnn": 50, "pointconv": 100}}
    assert CONFIG == want
    for model, bs in want["shapley_batch_size"].items():
        a = types.SimpleNamespace(model=model)
        final_util.set_shapley_batch_size(a)
        final_util.set_interaction_batch_size(a)
        assert a.shapley_batch_size == bs and a.interaction_batch_size == want["interaction_batch_size"][model]
    a = types.SimpleNamespace(model="gcnn_adv")
    final_util.set_shapley_batch_size(a)
    assert a.shapley_batch_size == 10
    with pytest.raises(Exception):
        final_util.set_shapley_batch_size(types.SimpleNamespace(model="nope"))


def test_set_model_args_and_module_prefix():
    a = types.SimpleNamespace(model="dgcnn", dataset="modelnet10")
    final_util.set_model_args(a)
    assert a.k == 20 and a.model_path.endswith("model_best.t7")
    a = types.SimpleNamespace(model="pointnet", dataset="shapenet")
    final_util.set_model_args(a)
    assert a.feature_transform is True
    with pytest.raises(Exception):
        final_util.set_model_args(types.SimpleNamespace(model="dgcnn", dataset="imagenet"))
    sd = {"module." + k: torch.from_numpy(np.asarray(v)) for k, v in synthetic.make_state_dict("gcnn").items()}
    args = types.SimpleNamespace(model="gcnn", k=20, dataset="shapenet", device="cpu")
    model = final_util.build_model(args, sd)
    assert not model.training
    assert list(model.state_dict().keys()) == [k for k, _, _ in synthetic.state_dict_spec("gcnn")]


def test_checkpoint_round_trip_through_load_model(tmp_path):
    path = str(tmp_path / "model_best.t7")
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in synthetic.make_state_dict("pointnet").items()}
    torch.save(sd, path)
    args = types.SimpleNamespace(model="pointnet", dataset="shapenet", feature_transform=True, device="cpu", model_path=path)
    model = final_util.load_model(args)
    got = model.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)


def test_modelnet40_head_has_40_classes():
    a = types.SimpleNamespace(model="dgcnn", k=20, dataset="modelnet40", device="cpu")
    m = final_util.MODEL_CLASSES["dgcnn"](a)
    assert m.state_dict()["linear3.weight"].shape == (40, 256)


def test_seed_replay_of_generate_all_orders():
    from interpret_quality_b200.final_shapley_value import generate_all_orders
    final_util.set_random(1)
    a = types.SimpleNamespace(num_samples_save=20, num_regions=32)
    orders = generate_all_orders(None, a, save=False)
    assert np.array_equal(orders, synthetic.make_orders(20, 32, seed=1))


def test_f16_operand_split_host(lib):
    """The two-term fp16 split of the kind::f16 tensor-core paths (csrc/common.cuh::split_f16, the host half the library
    applies to weights): bit patterns equal a numpy restatement, the pair reproduces the value to 2^-22, the scale is the
    power of two that brings max|w| into [2^9, 2^10), and nothing overflows."""
    rs = np.random.RandomState(3)
    for w in (rs.normal(size=4096).astype(np.float32) * 0.05,
              (rs.normal(size=1000) * rs.uniform(1e-4, 30.0, size=1000)).astype(np.float32),
              np.array([0.0, -0.0, 1.0, -1.0, 3.0e-6, 123.456], np.float32),
              np.zeros(7, np.float32)):
        hi = np.zeros(w.size, np.uint16)
        lo = np.zeros(w.size, np.uint16)
        scale = ctypes.c_float(0.0)
        rc = lib.iq_split_f16_host(w.ctypes.data, w.size, hi.ctypes.data, lo.ctypes.data, ctypes.byref(scale))
        assert rc == 0
        s = np.float32(scale.value)
        mx = np.abs(w).max()
        if mx > 0:
            assert 512.0 <= mx * s < 1024.0 and float(np.log2(s)).is_integer()
        else:
            assert s == 1.0
        xs = w * s                                                   # exact: s is a power of two
        want_hi = xs.astype(np.float16)
        want_lo = (xs - want_hi.astype(np.float32)).astype(np.float16)
        assert np.array_equal(hi, want_hi.view(np.uint16)) and np.array_equal(lo, want_lo.view(np.uint16))
        back = (hi.view(np.float16).astype(np.float64) + lo.view(np.float16).astype(np.float64)) / float(s)
        assert np.isfinite(back).all()
        # 22 significand bits while the low term is a normal fp16 number; the fp16 subnormal spacing below that
        err = np.abs(back - w.astype(np.float64))
        assert (err <= np.maximum(np.abs(w) * 2.0 ** -22, 2.0 ** -25 / float(s))).all()


def test_f16_paths_switches(lib):
    """iq_f16_paths(): every kind::f16 path on by default; IQ_F16_CONV5=0 switches all of them off (the other two read the
    fp16 activation buffers conv5's path writes), the other switches only their own path."""
    keys = ("IQ_F16_CONV5", "IQ_F16_STORE", "IQ_F16_GRAM")
    before = {k: os.environ.pop(k, None) for k in keys}
    try:
        lib.iq_debug_reload_env()
        assert _lib.f16_paths() == 7
        for env, want in (({"IQ_F16_CONV5": "0"}, 0), ({"IQ_F16_STORE": "0"}, 5), ({"IQ_F16_GRAM": "0"}, 3),
                          ({"IQ_F16_STORE": "0", "IQ_F16_GRAM": "0"}, 1)):
            os.environ.update(env)
            lib.iq_debug_reload_env()
            assert _lib.f16_paths() == want, (env, _lib.f16_paths())
            for k in env:
                del os.environ[k]
    finally:
        for k, v in before.items():
            if v is not None:
                os.environ[k] = v
        lib.iq_debug_reload_env()
