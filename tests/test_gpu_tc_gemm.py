"""Unit parity of the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) against float64 torch and the exact fp32 SIMT GEMM."""
import numpy as np
import pytest
import torch

from interpret_quality_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 64), (256, 128, 128), (1024, 256, 64), (4096, 512, 128),
                                   (2048, 1024, 512), (128 * 149, 128, 96), (1, 128, 64), (200, 256, 260), (5440, 1024, 2048),
                                   (128 * 3 + 127, 320, 132)])
@pytest.mark.parametrize("act", [0, 2])
def test_store_epilogue_vs_float64(M, N, K, act):
    """Includes row counts that are not multiples of 128 (TMA zero-fills the ragged A tile and clips the stores) and K
    that is not a multiple of 32 (zero-filled K tail): the shapes of the classifier heads and the per-point products."""
    rs = np.random.RandomState(M + N + K)
    x = (rs.normal(size=(M, K)) * rs.uniform(0.1, 4.0, size=(1, K))).astype(np.float32)
    w = rs.normal(size=(N, K)).astype(np.float32)
    b = rs.normal(size=(N,)).astype(np.float32)
    want = torch.from_numpy(x).double() @ torch.from_numpy(w).double().T + torch.from_numpy(b).double()
    if act == 2:
        want = torch.nn.functional.leaky_relu(want, 0.2)
    want = want.numpy()
    got_tc = ops.linear(cu(x), cu(w), cu(b), act=act, engine=1).cpu().numpy()
    got_fp32 = ops.linear(cu(x), cu(w), cu(b), act=act, engine=0).cpu().numpy()
    scale = np.abs(want).max()
    e_tc, e_fp32 = np.abs(got_tc - want).max() / scale, np.abs(got_fp32 - want).max() / scale
    assert e_fp32 <= 2e-6
    # 3xTF32 sits at fp32 noise; single-pass TF32 would be ~5e-4.  The dropped Alo*Blo terms add up with sqrt(K):
    # 1.4e-5 at K = 2048 (products that long keep the fourth term in the models: TcGemm::four_terms)
    assert e_tc <= (1e-5 if K <= 512 else 2e-5), (e_tc, e_fp32)


def test_single_pass_tf32_would_fail_this_bar():
    """Sanity of the bar above: rounding the operands to TF32 once is two orders of magnitude worse."""
    rs = np.random.RandomState(0)
    x = rs.normal(size=(256, 128)).astype(np.float32)
    w = rs.normal(size=(128, 128)).astype(np.float32)
    rnd = lambda a: ((a.view(np.uint32) + 0x1000) & 0xffffe000).view(np.float32)
    want = x.astype(np.float64) @ w.astype(np.float64).T
    tf32 = rnd(x.copy()).astype(np.float64) @ rnd(w.copy()).astype(np.float64).T
    assert np.abs(tf32 - want).max() / np.abs(want).max() > 1e-4


@pytest.mark.parametrize("clouds,points,N,K", [(1, 128, 128, 32), (3, 1024, 1024, 512), (5, 2048, 256, 128),
                                               (160, 1024, 128, 64)])
@pytest.mark.parametrize("engine", [0, 1])
def test_pool_epilogue_vs_float64(clouds, points, N, K, engine):
    rs = np.random.RandomState(clouds + points + N + K)
    x = rs.normal(size=(clouds * points, K)).astype(np.float32)
    x[points // 2:points // 2 + 40] = x[0]                       # duplicated points -> argmax ties
    w = rs.normal(size=(N, K)).astype(np.float32)
    b = rs.normal(size=(N,)).astype(np.float32)
    y = torch.nn.functional.leaky_relu(torch.from_numpy(x).double() @ torch.from_numpy(w).double().T
                                       + torch.from_numpy(b).double(), 0.2).view(clouds, points, N)
    mx, mean, arg = ops.linear_pool(cu(x), cu(w), cu(b), clouds, points, act=2, engine=engine, want_arg=True)
    scale = float(y.abs().max())
    assert np.abs(mx.cpu().numpy() - y.max(1)[0].numpy()).max() / scale <= 1e-5
    assert np.abs(mean.cpu().numpy() - y.mean(1).numpy()).max() / scale <= 1e-5
    # the reported arg-max point reaches the max value (ties / fp32 noise may move the index itself)
    picked = torch.gather(y, 1, arg.cpu().view(clouds, 1, N)).squeeze(1).numpy()
    assert np.abs(picked - y.max(1)[0].numpy()).max() / scale <= 1e-5
    assert int(arg.min()) >= 0 and int(arg.max()) < points


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 8), (1024, 256, 64), (4096, 512, 128), (2048, 1024, 512),
                                   (128 * 149, 128, 96), (200, 256, 264), (1, 128, 64), (128 * 3 + 127, 384, 136)])
@pytest.mark.parametrize("act", [0, 2])
def test_f16_split_store_vs_float64(M, N, K, act):
    """kind::f16 MMAs on two-term fp16 splits of the scaled operands (engine 2, the form conv5 runs in): the same
    fp32-noise bar as 3xTF32, including ragged row tiles and K tails of the 64-element ring stage."""
    rs = np.random.RandomState(M + N + K)
    x = (rs.normal(size=(M, K)) * rs.uniform(0.1, 4.0, size=(1, K))).astype(np.float32)
    w = rs.normal(size=(N, K)).astype(np.float32)
    b = rs.normal(size=(N,)).astype(np.float32)
    want = torch.from_numpy(x).double() @ torch.from_numpy(w).double().T + torch.from_numpy(b).double()
    if act == 2:
        want = torch.nn.functional.leaky_relu(want, 0.2)
    want = want.numpy()
    got = ops.linear(cu(x), cu(w), cu(b), act=act, engine=2).cpu().numpy()
    got_tf32 = ops.linear(cu(x), cu(w), cu(b), act=act, engine=1).cpu().numpy()
    scale = np.abs(want).max()
    e16, e32 = np.abs(got - want).max() / scale, np.abs(got_tf32 - want).max() / scale
    print("f16x2 %.2e  3xtf32 %.2e of scale (M %d N %d K %d)" % (e16, e32, M, N, K))
    assert e16 <= 1e-5, (e16, e32)


@pytest.mark.parametrize("magnitude,bar", [(1e-3, 1e-5), (1e-5, 1e-3), (1500.0, 1e-5), (3500.0, 2e-3)])
def test_f16_split_range(magnitude, bar):
    """Range behaviour of the fp16 split (activations are stored as fp16 pairs of 8 x): small values lose only their low
    term to the fp16 subnormal spacing (absolute error 4e-9), values beyond 8188 saturate the high term and degrade
    gracefully up to 16376 -- never an infinity."""
    rs = np.random.RandomState(7)
    x = (rs.normal(size=(256, 128)) * magnitude).astype(np.float32)
    x = np.clip(x, -16000.0, 16000.0)
    w = rs.normal(size=(128, 128)).astype(np.float32)
    want = x.astype(np.float64) @ w.astype(np.float64).T
    got = ops.linear(cu(x), cu(w), None, engine=2).cpu().numpy()
    assert np.isfinite(got).all()
    err = np.abs(got - want).max() / np.abs(want).max()
    print("|x| ~ %g: %.2e of scale" % (magnitude, err))
    assert err <= bar


@pytest.mark.parametrize("clouds,points,N,K", [(1, 128, 128, 32), (3, 1024, 1024, 512), (5, 2048, 256, 128),
                                               (160, 1024, 128, 64), (7, 384, 1000, 512)])
def test_f16_split_pool_vs_float64(clouds, points, N, K):
    rs = np.random.RandomState(clouds + points + N + K)
    x = rs.normal(size=(clouds * points, K)).astype(np.float32)
    x[points // 2:points // 2 + 40] = x[0]
    w = rs.normal(size=(N, K)).astype(np.float32)
    b = rs.normal(size=(N,)).astype(np.float32)
    y = torch.nn.functional.leaky_relu(torch.from_numpy(x).double() @ torch.from_numpy(w).double().T
                                       + torch.from_numpy(b).double(), 0.2).view(clouds, points, N)
    mx, mean, arg = ops.linear_pool(cu(x), cu(w), cu(b), clouds, points, act=2, engine=2, want_arg=True)
    scale = float(y.abs().max())
    assert np.abs(mx.cpu().numpy() - y.max(1)[0].numpy()).max() / scale <= 1e-5
    assert np.abs(mean.cpu().numpy() - y.mean(1).numpy()).max() / scale <= 1e-5
    picked = torch.gather(y, 1, arg.cpu().view(clouds, 1, N)).squeeze(1).numpy()
    assert np.abs(picked - y.max(1)[0].numpy()).max() / scale <= 1e-5


@pytest.mark.parametrize("name", ["dgcnn", "gcnn"])
def test_f16_paths_equal_3xtf32(name):
    """The kind::f16 forms of conv5, of the tcgen05 EdgeConv products and of the Gram kNN nomination (IQ_F16_CONV5 /
    IQ_F16_STORE / IQ_F16_GRAM) against the all-3xTF32 forward on masked clouds, plain and collapsed.  Nothing upstream of a
    kNN decision changes (the nomination is re-ranked exactly either way), so the logits agree to fp32 noise."""
    import os
    import types
    from interpret_quality_b200 import _lib, synthetic
    from interpret_quality_b200.tools import final_util
    from oracle import coalition, geom
    a = types.SimpleNamespace(model=name, k=20, dataset="shapenet", device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict(name))
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, 32)[0])
    center = coalition.center_of(data)
    masked = geom.mask_shapley(data[0], center, synthetic.make_orders(2, 32), rid)
    x = cu(masked)
    keys = ("IQ_F16_CONV5", "IQ_F16_STORE", "IQ_F16_GRAM")

    def run(bits, collapsed):
        before = {key: os.environ.get(key) for key in keys}
        for i, key in enumerate(keys):
            os.environ[key] = "1" if bits & (1 << i) else "0"
        _lib.load().iq_debug_reload_env()
        try:
            assert _lib.f16_paths() == bits
            masked_to = cu(np.asarray(center, dtype=np.float32).reshape(3)) if collapsed else None
            return model.forward_point_major(x, masked_to=masked_to).cpu().numpy()
        finally:
            for key, val in before.items():
                if val is None:
                    del os.environ[key]
                else:
                    os.environ[key] = val
            _lib.load().iq_debug_reload_env()

    for collapsed in (False, True):
        base = run(0, collapsed)
        scale = np.abs(base).max()
        for bits in (1, 3, 5, 7):
            err = np.abs(run(bits, collapsed) - base).max() / scale
            print("%s %s f16 paths %d vs 3xtf32: %.2e of scale" % (name, "collapsed" if collapsed else "plain", bits, err))
            assert err <= 1e-5, (bits, err)                      # two fp32-noise evaluations: measured 3e-6 ... 4e-6


def test_batched_gram_keys_through_dgcnn_engine_switch():
    """Same DGCNN forward through both engines: logits agree to fp32 noise, so the kNN graphs agree."""
    import types
    from interpret_quality_b200 import synthetic
    from interpret_quality_b200.tools import final_util
    from oracle import coalition, geom
    a = types.SimpleNamespace(model="dgcnn", k=20, dataset="shapenet", device=DEV)
    model = final_util.build_model(a, synthetic.make_state_dict("dgcnn"))
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, 32)[0])
    masked = geom.mask_shapley(data[0], coalition.center_of(data), synthetic.make_orders(2, 32), rid)
    x = cu(masked)
    model.set_engine("3xtf32")
    tc = model.forward_point_major(x).cpu().numpy()
    model.set_engine("fp32")
    fp = model.forward_point_major(x).cpu().numpy()
    per_cloud = np.abs(tc - fp).max(1) / np.abs(fp).max()
    # a near-tie neighbour may be decided differently by the two kNN paths on a few clouds
    assert np.median(per_cloud) <= 1e-5 and per_cloud.max() <= 1e-3


def test_pointnet2_chain_kernel_equals_two_kernel_route():
    """chain_tc.cu (layers 1-2-3 + group max in one kernel, H2 kept in TMEM) against the round-1 route (gathered-A STORE
    -> H2 in HBM -> POOL): the same 3xTF32 products, only the order of the two cross terms differs."""
    import os
    import types
    from interpret_quality_b200 import _lib, synthetic
    from interpret_quality_b200.tools import final_util
    from oracle import coalition, geom
    a = types.SimpleNamespace(model="pointnet2", k=20, dataset="shapenet", feature_transform=True, device="cuda:0")
    model = final_util.build_model(a, synthetic.make_state_dict("pointnet2"))
    data = synthetic.make_cloud(1024)
    rid = geom.region_id(data[0], geom.fps(data, 32)[0])
    masked = geom.mask_shapley(data[0], coalition.center_of(data), synthetic.make_orders(1, 32), rid)[::2]
    x = torch.from_numpy(masked).to("cuda:0")
    chained = model.forward_point_major(x).cpu().numpy()
    os.environ["IQ_TC_NO_CHAIN"] = "1"
    _lib.load().iq_debug_reload_env()
    try:
        two_kernels = model.forward_point_major(x).cpu().numpy()
    finally:
        del os.environ["IQ_TC_NO_CHAIN"]
        _lib.load().iq_debug_reload_env()
    err = np.abs(chained - two_kernels).max() / np.abs(two_kernels).max()
    print("pointnet2 chain vs two-kernel route: %.2e of scale" % err)
    assert err <= 2e-6


@pytest.mark.parametrize("C1,C2,C3,K", [(32, 32, 64, 16), (64, 64, 128, 32), (64, 96, 128, 128), (64, 64, 128, 32),
                                        (128, 128, 256, 64), (128, 128, 256, 128), (32, 32, 64, 128), (128, 128, 256, 16)])
def test_chained_grouped_mlp_kernel(C1, C2, C3, K):
    """csrc/chain_tc.cu alone on random data against a plain fp32 evaluation (torch, float64 accumulate): every
    (widths, group size) combination PointNet++ MSG uses plus the corners of the supported range; several tiles per CTA."""
    g = torch.Generator().manual_seed(C1 + C2 + C3 + K)
    clouds, nsrc = 3, 200
    S = 128 * 160 // K                                          # 160 tiles of 128 rows per cloud: > 148 CTAs, 3+ tiles each
    U = torch.randn((clouds * nsrc, C1), generator=g).cuda()
    V = torch.randn((clouds * S, C1), generator=g).cuda() * 0.5
    b1 = torch.randn((C1,), generator=g).cuda() * 0.1
    idx = torch.randint(0, nsrc, (clouds * S * K,), generator=g, dtype=torch.int32).cuda()
    W2 = (torch.randn((C2, C1), generator=g) / C1 ** 0.5).cuda()
    b2 = torch.randn((C2,), generator=g).cuda() * 0.1
    W3 = (torch.randn((C3, C2), generator=g) / C2 ** 0.5).cuda()
    b3 = torch.randn((C3,), generator=g).cuda() * 0.1
    got = ops.grouped_mlp_max(U, V, b1, idx, clouds, S, K, W2, b2, W3, b3)
    cloud_of = torch.arange(clouds * S * K, device="cuda") // (S * K)
    h1 = torch.relu(U[cloud_of * nsrc + idx.long()] - V.repeat_interleave(K, 0) + b1)
    h2 = torch.relu((h1.double() @ W2.double().T).float() + b2)
    h3 = torch.relu((h2.double() @ W3.double().T).float() + b3)
    want = h3.reshape(clouds * S, K, C3).max(1)[0]
    err = float((got - want).abs().max() / want.abs().max())
    assert err <= 2e-5, err
