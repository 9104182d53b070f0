"""Edge cases of the CUDA path the reference's own usage reaches: the 40-class head (dataset modelnet40,
models/dgcnn.py:56-58), wide logits rows in the reward / interaction kernels, single-cloud and empty batches."""
import types

import numpy as np
import pytest
import torch

from interpret_quality_b200 import ops, synthetic
from interpret_quality_b200.tools import final_common, final_util
from oracle import coalition, nets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
R = 32


def masked_clouds(n_perm=1):
    from oracle import geom
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    return geom.mask_shapley(data[0], coalition.center_of(data), synthetic.make_orders(n_perm, R), rid)


@pytest.mark.parametrize("name", ["pointnet", "gcnn"])
def test_forty_class_head(name):
    a = types.SimpleNamespace(model=name, k=20, dataset="modelnet40", feature_transform=True, device=DEV)
    sd = synthetic.make_state_dict(name, num_classes=40)
    model = final_util.build_model(a, sd)
    assert model.output_channels == 40
    x = masked_clouds()[::4]                                            # 9 clouds
    got = model.forward_point_major(torch.from_numpy(x).to(DEV)).cpu().numpy()
    want = nets.forward(name, torch.from_numpy(x).permute(0, 2, 1).contiguous(),
                        {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}).numpy()
    assert got.shape == (9, 40)
    assert np.abs(got - want).max() <= 1e-3 * np.abs(want).max()
    for soft in ("modified", "normal"):
        v = final_common.get_reward(torch.from_numpy(got).to(DEV), torch.tensor([17]), types.SimpleNamespace(softmax_type=soft))
        ref = coalition.reward(torch.from_numpy(got), 17, soft).numpy()
        assert np.abs(v.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


def test_wide_rows_in_the_reduction_kernels():
    """C = 40 and C = 100 logits rows: the shared-memory staging of reward_kernel / interaction_reduce_kernel."""
    from interpret_quality_b200.final_cal_interactions import compute_order_interaction
    rs = np.random.RandomState(4)
    for C, lbl in ((40, 39), (100, 0)):
        logits = (rs.normal(size=(3, 4 * 37, C)) * 3).astype(np.float32)
        got = compute_order_interaction(torch.from_numpy(logits).to(DEV), torch.tensor([lbl]), types.SimpleNamespace(softmax_type="modified"))
        v = coalition.reward(torch.from_numpy(logits.reshape(-1, C)), lbl, "modified").numpy().reshape(3, 37, 4)
        want = (v[..., 0] + v[..., 3]) - v[..., 1] - v[..., 2]
        assert got.shape == (3, 37) and np.abs(got - want).max() <= 2e-5 * max(np.abs(want).max(), 1e-3)
        r = ops.reward(torch.from_numpy(logits.reshape(-1, C)).to(DEV), lbl, "normal").cpu().numpy()
        rr = coalition.reward(torch.from_numpy(logits.reshape(-1, C)), lbl, "normal").numpy()
        assert np.abs(r - rr).max() <= 1e-5 * np.abs(rr).max()


@pytest.mark.parametrize("name", ["dgcnn", "pointnet2"])
def test_single_cloud_and_empty_batch(name):
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict(name))
    x = torch.from_numpy(masked_clouds()[5:12]).to(DEV)
    full = model.forward_point_major(x).cpu().numpy()
    one = model.forward_point_major(x[3:4].contiguous()).cpu().numpy()
    assert np.abs(one - full[3:4]).max() <= 1e-6 * np.abs(full).max()      # a cloud's logits do not depend on its batch
    empty = model.forward_point_major(x[:0].contiguous())
    assert tuple(empty.shape) == (0, 10)


@pytest.mark.parametrize("name", ["pointnet2", "pointconv"])
def test_gathered_a_gemm_is_bitwise_the_two_kernel_route(name, monkeypatch):
    """Grouped-MLP layers 1+2 fused in the tcgen05 GEMM (gathered A) against group_sub_act + GEMM through HBM."""
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict(name))
    x = torch.from_numpy(masked_clouds()[3:17]).to(DEV)
    fused = model.forward_point_major(x).clone()
    monkeypatch.setenv("IQ_TC_NO_GATHER", "1")
    unfused = model.forward_point_major(x).clone()
    assert torch.equal(fused, unfused)


@pytest.mark.parametrize("k", [5, 12])
def test_dgcnn_other_neighbourhood_sizes(k):
    """args.k other than the default 20 (models/dgcnn.py:55): fused kNN, re-rank and gather take k at run time."""
    a = types.SimpleNamespace(model="dgcnn", k=k, dataset="shapenet", device=DEV)
    sd = synthetic.make_state_dict("dgcnn")
    model = final_util.build_model(a, sd)
    x = masked_clouds()[::5]                                            # 7 clouds, from fully masked to untouched
    got = model.forward_point_major(torch.from_numpy(x).to(DEV)).cpu().numpy()
    want = nets.forward("dgcnn", torch.from_numpy(x).permute(0, 2, 1).contiguous(),
                        {kk: torch.from_numpy(np.asarray(v)) for kk, v in sd.items()}, k=k).numpy()
    assert np.abs(got - want).max() <= 1e-3 * np.abs(want).max()


GUARD = 1 << 20


def guarded(nbytes):
    """A uint8 buffer of nbytes between two 1 MiB bands of 0xA5; returns (whole, interior view)."""
    nbytes = (int(nbytes) + 255) // 256 * 256
    whole = torch.full((GUARD + nbytes + GUARD,), 0xA5, dtype=torch.uint8, device=DEV)
    return whole, whole[GUARD:GUARD + nbytes]


def intact(whole):
    return bool((whole[:GUARD] == 0xA5).all()) and bool((whole[-GUARD:] == 0xA5).all())


@pytest.mark.parametrize("name", ["dgcnn", "gcnn", "pointnet", "pointnet2", "pointconv"])
def test_forward_stays_inside_workspace_logits_and_input(name):
    """compute-sanitizer is not available on the GPU pool, so out-of-bounds writes are looked for with guard bands:
    the forward gets EXACTLY iq_model_workspace_bytes of scratch, its logits and its input between 0xA5 bands, for a
    batch that is not a multiple of the internal chunk; bands must survive and the result must equal the normal call."""
    from interpret_quality_b200 import _lib
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict(name))
    model.set_chunk(8)
    x = torch.from_numpy(masked_clouds()[:19]).to(DEV)                    # 19 clouds: chunks of 8, 8, 3
    B, N = x.shape[0], x.shape[1]
    want = model.forward_point_major(x)
    lib = _lib.load()
    h = model._get_handle()
    need = lib.iq_model_workspace_bytes(h, B, N)
    assert need > 0
    ws_whole, ws = guarded(need)
    out_whole, out = guarded(B * model.output_channels * 4)
    in_whole, xin = guarded(x.numel() * 4)
    xin[:x.numel() * 4].copy_(x.reshape(-1).view(torch.uint8))
    with torch.cuda.device(DEV):
        _lib.check(lib.iq_model_forward(h, xin.data_ptr(), 1, B, N, out.data_ptr(), ws.data_ptr(), int(need), 0, 0,
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert intact(ws_whole) and intact(out_whole) and intact(in_whole)
    got = out[:B * model.output_channels * 4].view(torch.float32).reshape(B, model.output_channels)
    assert torch.equal(got, want)
    assert torch.equal(xin[:x.numel() * 4].view(torch.float32).reshape(x.shape), x)          # the input is read-only
    # the coalition entry (collapsed clouds, other chunking) under the same guard bands
    center = torch.from_numpy(coalition.center_of(synthetic.make_cloud(1024))).to(DEV)
    out.fill_(0)
    with torch.cuda.device(DEV):
        _lib.check(lib.iq_model_forward_coalitions(h, xin.data_ptr(), 1, B, N, center.data_ptr(), out.data_ptr(),
                                                   ws.data_ptr(), int(need), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert intact(ws_whole) and intact(out_whole) and intact(in_whole)
    got = out[:B * model.output_channels * 4].view(torch.float32).reshape(B, model.output_channels)
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())
    # iq_model_workspace_bytes covers both entries; a workspace too small for either must be refused, not overrun
    short = int(need) // 3 // 256 * 256
    with torch.cuda.device(DEV):
        rc = lib.iq_model_forward(h, xin.data_ptr(), 1, B, N, out.data_ptr(), ws.data_ptr(), short, 0, 0,
                                  torch.cuda.current_stream().cuda_stream)
        assert rc != 0 and "workspace" in _lib.last_error()
        rc = lib.iq_model_forward_coalitions(h, xin.data_ptr(), 1, B, N, center.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                             short, torch.cuda.current_stream().cuda_stream)
        assert rc != 0 and "workspace" in _lib.last_error()
    torch.cuda.synchronize()
    assert intact(ws_whole)


def test_coalition_kernels_stay_inside_their_outputs():
    """mask_shapley writes exactly its ((R+1)*bs, N, 3) output: guard bands either side survive."""
    from oracle import geom
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, R)[0])
    d = torch.from_numpy(data[0]).to(DEV)
    center = torch.from_numpy(coalition.center_of(data)).to(DEV)      # the same fp32 centre on both sides
    orders = ops.to_dev_i64(synthetic.make_orders(3, R), DEV)
    region = ops.to_dev_i64(rid, DEV)
    rows = 3 * (R + 1)
    whole, buf = guarded(rows * 1024 * 3 * 4)
    out = buf[:rows * 1024 * 3 * 4].view(torch.float32).reshape(rows, 1024, 3)
    ops.mask_shapley(d, center, orders, region, out=out)
    torch.cuda.synchronize()
    assert intact(whole)
    want = geom.mask_shapley(data[0], coalition.center_of(data), synthetic.make_orders(3, R), rid)
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("name", ["dgcnn", "pointnet2"])
def test_lanes_do_not_change_the_result(name):
    """Chunks in flight (iq_model_set_lanes) is scheduling only: 1, 2, 3 and 4 lanes give bit-identical logits, also
    when the caller's stream is not the default one and has work queued before and after the forward."""
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict(name))
    model.set_chunk(4)
    x = torch.from_numpy(masked_clouds()[:23]).to(DEV)                    # 6 chunks, the last one ragged
    model.set_lanes(1)
    want = model.forward_point_major(x).clone()
    side = torch.cuda.Stream(device=DEV)
    for lanes in (2, 3, 4):
        model.set_lanes(lanes)
        assert torch.equal(model.forward_point_major(x), want)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            xs = x * 1.0                                                    # produced on the caller's stream just before
            got = model.forward_point_major(xs)
            total = got.sum()                                               # consumed on it right after
        side.synchronize()
        assert torch.equal(got, want) and torch.isfinite(total)
    with pytest.raises(Exception):
        model.set_lanes(5)


@pytest.mark.parametrize("name,N", [("dgcnn", 1000), ("dgcnn", 200), ("gcnn", 1000), ("pointnet", 1000), ("pointnet", 77)])
def test_any_number_of_points(name, N):
    """The reference takes N from the tensor's shape (models/dgcnn.py:12-18, models/pointnet.py:77-88); the tensor-core
    kernels tile clouds in 128-point blocks.  Other N run the exact fp32 route (DGCNN / GCNN) or are padded with copies of
    the first point (PointNet: max pools only) -- against the oracle on masked and unmasked clouds, with and without the
    coalition hint, and PointNet's critical-point indices against the oracle's."""
    from oracle import geom, nets
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=DEV)
    sd = synthetic.make_state_dict(name)
    model = final_util.build_model(a, sd)
    data = synthetic.make_cloud(N)
    regions = 8
    rid = geom.region_id(data[0], geom.fps(data, regions)[0])
    center = coalition.center_of(data)
    masked = geom.mask_shapley(data[0], center, synthetic.make_orders(1, regions), rid)          # (9, N, 3)
    x = torch.from_numpy(masked).permute(0, 2, 1).contiguous()
    want = nets.forward(name, x, sd).numpy()
    got = model(x.to(DEV))
    got_l = (got[0] if isinstance(got, tuple) else got).cpu().numpy()
    scale = np.abs(want).max()
    err = np.abs(got_l - want).max(1) / scale
    print("%s N=%d: per-cloud error vs oracle max %.2e median %.2e" % (name, N, err.max(), np.median(err)))
    assert np.median(err) <= 1e-5 and err.max() <= 1e-3
    hinted = model.forward_point_major(torch.from_numpy(masked).to(DEV), masked_to=torch.from_numpy(center).to(DEV))
    assert np.array_equal(hinted.cpu().numpy(), got_l)                   # no collapse at this N: the plain route, bit for bit
    if name == "pointnet":
        want_crt = nets.pointnet(x[-1:], sd)[2].numpy()                  # the unmasked cloud: no duplicated points, no ties
        assert (got[2][-1:].cpu().numpy() == want_crt).mean() > 0.99


@pytest.mark.parametrize("name", ["dgcnn", "pointnet2"])
def test_same_handle_on_two_streams_back_to_back(name):
    """A model handle owns ONE workspace: forwards issued back to back on two different streams must not overlap on it.
    The library orders them (an event at the end of every forward); both results must equal the serial ones."""
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", feature_transform=True, device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict(name))
    x1 = torch.from_numpy(masked_clouds(2)[:40]).to(DEV)
    x2 = torch.from_numpy(masked_clouds(2)[20:60]).to(DEV).contiguous()
    center = torch.from_numpy(coalition.center_of(synthetic.make_cloud(1024))).to(DEV)
    want1 = model.forward_point_major(x1, masked_to=center).clone()
    want2 = model.forward_point_major(x2, masked_to=center).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(device=DEV), torch.cuda.Stream(device=DEV)
    for _ in range(3):
        with torch.cuda.stream(s1):
            got1 = model.forward_point_major(x1, masked_to=center)
        with torch.cuda.stream(s2):
            got2 = model.forward_point_major(x2, masked_to=center)      # same handle, other stream, no host sync in between
        s1.synchronize()
        s2.synchronize()
        assert torch.equal(got1, want1) and torch.equal(got2, want2)
