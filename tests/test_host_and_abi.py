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
