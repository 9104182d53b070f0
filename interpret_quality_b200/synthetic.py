"""Deterministic synthetic inputs for parity tests and the benchmark.

Datasets and trained checkpoints of the reference are external downloads
(README.md:38-42,68 of the reference) and unavailable offline, so every test and
bench line runs on synthetic clouds and seeded "trained-like" weights
(SURVEY.md section 8d).  Everything here is numpy-only and uses *uniform* draws
from the frozen MT19937 legacy stream, so the same bytes come out on any host.

Nothing in this module touches the GPU or the oracle.
"""
import zlib

import numpy as np

__all__ = ["make_cloud", "make_state_dict", "make_orders", "make_pairs_and_contexts",
           "MODEL_SPECS", "state_dict_spec"]


def make_cloud(num_points=1024, seed=1234):
    """Surface-style cloud: half the points on a sphere shell, half on box faces.

    Returns float32 (1, num_points, 3), centred and scaled to max-norm 1 the way
    the reference's ShapeNet loader normalises (final_data_shapley.py:155-157).
    """
    rs = np.random.RandomState(seed)
    n_s = num_points // 2
    n_b = num_points - n_s
    # sphere shell radius 0.6, offset +0.2 in x (uniform on sphere via z/phi)
    z = rs.uniform(-1.0, 1.0, n_s)
    phi = rs.uniform(0.0, 2.0 * np.pi, n_s)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    sph = 0.6 * np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    sph[:, 0] += 0.2
    # faces of a 1.2 x 0.6 x 0.8 box, offset -0.3 in x
    half = np.array([0.6, 0.3, 0.4])
    pts = rs.uniform(-1.0, 1.0, (n_b, 3)) * half
    face = rs.randint(0, 6, n_b)
    axis = face // 2
    sign = np.where(face % 2 == 0, -1.0, 1.0)
    pts[np.arange(n_b), axis] = sign * half[axis]
    pts[:, 0] -= 0.3
    cloud = np.concatenate([sph, pts], axis=0)
    cloud = cloud[rs.permutation(num_points)]
    cloud = cloud - cloud.mean(axis=0, keepdims=True)
    cloud = cloud / np.sqrt((cloud ** 2).sum(axis=1)).max()
    return np.ascontiguousarray(cloud[None].astype(np.float32))


# ---------------------------------------------------------------------------
# state-dict layouts of the five reference classifiers (names and shapes are
# the checkpoint format of the reference: models/*.py; verified key-for-key by
# tests/golden/make_golden.py against the reference's own state_dict()).
# An entry is (prefix, kind, shape...) with kind in
#   conv  : weight (out,in,*ones) [+ bias]
#   lin   : weight (out,in) [+ bias]
#   bn    : weight,bias,running_mean,running_var (c,), num_batches_tracked ()
# ---------------------------------------------------------------------------
def _conv(prefix, cout, cin, nd, bias=True):
    return (prefix, "conv", (cout, cin) + (1,) * nd, bias)


def _lin(prefix, cout, cin, bias=True):
    return (prefix, "lin", (cout, cin), bias)


def _bn(prefix, c):
    return (prefix, "bn", (c,), True)


def _edgeconv_family():
    spec = []
    chans = [(6, 64), (128, 64), (128, 128), (256, 256)]
    for i, (ci, co) in enumerate(chans, 1):
        spec.append(_bn("bn%d" % i, co))
    spec.append(_bn("bn5", 1024))
    for i, (ci, co) in enumerate(chans, 1):
        spec.append(_conv("conv%d.0" % i, co, ci, 2, bias=False))
        spec.append(_bn("conv%d.1" % i, co))          # alias of bn<i> in the reference
    spec.append(_conv("conv5.0", 1024, 512, 1, bias=False))
    spec.append(_bn("conv5.1", 1024))
    spec.append(_lin("linear1", 512, 2048, bias=False))
    spec.append(_bn("bn6", 512))
    spec.append(_lin("linear2", 256, 512))
    spec.append(_bn("bn7", 256))
    spec.append(_lin("linear3", 10, 256))
    return spec


def _stn(prefix, k):
    return [_conv(prefix + "conv1", 64, k, 1), _conv(prefix + "conv2", 128, 64, 1),
            _conv(prefix + "conv3", 1024, 128, 1), _lin(prefix + "fc1", 512, 1024),
            _lin(prefix + "fc2", 256, 512), _lin(prefix + "fc3", k * k, 256),
            _bn(prefix + "bn1", 64), _bn(prefix + "bn2", 128), _bn(prefix + "bn3", 1024),
            _bn(prefix + "bn4", 512), _bn(prefix + "bn5", 256)]


def _pointnet():
    spec = _stn("feat.stn.", 3)
    spec += [_conv("feat.conv1", 64, 3, 1), _conv("feat.conv2", 128, 64, 1),
             _conv("feat.conv3", 1024, 128, 1), _bn("feat.bn1", 64), _bn("feat.bn2", 128),
             _bn("feat.bn3", 1024)]
    spec += _stn("feat.fstn.", 64)
    spec += [_lin("fc1", 512, 1024), _lin("fc2", 256, 512), _lin("fc3", 10, 256),
             _bn("bn1", 512), _bn("bn2", 256)]
    return spec


def _pointnet2():
    spec = []

    def msg(prefix, cin, mlps):
        for b, mlp in enumerate(mlps):
            last = cin + 3
            for j, co in enumerate(mlp):
                spec.append(_conv("%s.conv_blocks.%d.%d" % (prefix, b, j), co, last, 2))
                last = co
        for b, mlp in enumerate(mlps):
            for j, co in enumerate(mlp):
                spec.append(_bn("%s.bn_blocks.%d.%d" % (prefix, b, j), co))

    msg("sa1", 0, [[32, 32, 64], [64, 64, 128], [64, 96, 128]])
    msg("sa2", 320, [[64, 64, 128], [128, 128, 256], [128, 128, 256]])
    last = 643
    for j, co in enumerate([256, 512, 1024]):
        spec.append(_conv("sa3.mlp_convs.%d" % j, co, last, 2))
        last = co
    for j, co in enumerate([256, 512, 1024]):
        spec.append(_bn("sa3.mlp_bns.%d" % j, co))
    spec += [_lin("fc1", 512, 1024), _bn("bn1", 512), _lin("fc2", 256, 512), _bn("bn2", 256),
             _lin("fc3", 10, 256)]
    return spec


def _pointconv():
    spec = []

    def sa(prefix, cin, mlp):
        last = cin
        for j, co in enumerate(mlp):
            spec.append(_conv("%s.mlp_convs.%d" % (prefix, j), co, last, 2))
            last = co
        for j, co in enumerate(mlp):
            spec.append(_bn("%s.mlp_bns.%d" % (prefix, j), co))
        wn = [(8, 3), (8, 8), (16, 8)]
        for j, (co, ci) in enumerate(wn):
            spec.append(_conv("%s.weightnet.mlp_convs.%d" % (prefix, j), co, ci, 2))
        for j, (co, ci) in enumerate(wn):
            spec.append(_bn("%s.weightnet.mlp_bns.%d" % (prefix, j), co))
        spec.append(_lin("%s.linear" % prefix, mlp[-1], 16 * mlp[-1]))
        spec.append(_bn("%s.bn_linear" % prefix, mlp[-1]))
        dn = [(16, 1), (8, 16), (1, 8)]
        for j, (co, ci) in enumerate(dn):
            spec.append(_conv("%s.densitynet.mlp_convs.%d" % (prefix, j), co, ci, 2))
        for j, (co, ci) in enumerate(dn):
            spec.append(_bn("%s.densitynet.mlp_bns.%d" % (prefix, j), co))

    sa("sa1", 3, [64, 64, 128])
    sa("sa2", 131, [128, 128, 256])
    sa("sa3", 259, [256, 512, 1024])
    spec += [_lin("fc1", 512, 1024), _bn("bn1", 512), _lin("fc2", 256, 512), _bn("bn2", 256),
             _lin("fc3", 10, 256)]
    return spec


MODEL_SPECS = {
    "dgcnn": _edgeconv_family,
    "gcnn": _edgeconv_family,
    "gcnn_adv": _edgeconv_family,
    "pointnet": _pointnet,
    "pointnet2": _pointnet2,
    "pointconv": _pointconv,
}

# in the DGCNN/GCNN checkpoints "convK.1.*" and "bnK.*" are the same module
_ALIASES = {"conv%d.1" % i: "bn%d" % i for i in range(1, 6)}


def state_dict_spec(model, num_classes=10):
    """[(key, shape, dtype)] in checkpoint order for `model`."""
    out = []
    for prefix, kind, shape, bias in MODEL_SPECS[model]():
        if prefix in ("linear3", "fc3") and kind == "lin":
            shape = (num_classes, shape[1])
        if kind in ("conv", "lin"):
            out.append((prefix + ".weight", shape, np.float32))
            if bias:
                out.append((prefix + ".bias", (shape[0],), np.float32))
        else:
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                out.append((prefix + "." + leaf, shape, np.float32))
            out.append((prefix + ".num_batches_tracked", (), np.int64))
    return out


def _rs_for(key, seed):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def make_state_dict(model, seed=7, num_classes=10):
    """Seeded "trained-like" weights as {key: ndarray} (SURVEY.md section 7.2 recipe).

    conv/linear weights ~ U(+-2/sqrt(fan_in)), biases ~ U(+-1/sqrt(fan_in)),
    BN gamma ~ U(0.5,1.5), beta and running_mean ~ U(+-0.17), running_var ~ U(0.5,1.5);
    PointConv's DensityNet output BN gets bias 1.0 so its always-on ReLU
    (models/pointconv.py:230-233 of the reference) is not dead.
    """
    sd = {}
    for key, shape, dtype in state_dict_spec(model, num_classes):
        prefix, leaf = key.rsplit(".", 1)
        src_prefix = _ALIASES.get(prefix, prefix) if model in ("dgcnn", "gcnn", "gcnn_adv") else prefix
        rs = _rs_for(src_prefix + "." + leaf, seed)
        if leaf == "num_batches_tracked":
            val = np.array(100, dtype=np.int64)
        elif len(shape) >= 2:                      # conv / linear weight
            fan_in = int(np.prod(shape[1:]))
            b = 2.0 / np.sqrt(fan_in)
            val = rs.uniform(-b, b, shape).astype(np.float32)
        elif leaf == "bias" and not _is_bn(model, prefix):
            fan_in = _fan_in_of(model, prefix, num_classes)
            b = 1.0 / np.sqrt(fan_in)
            val = rs.uniform(-b, b, shape).astype(np.float32)
        elif leaf in ("weight", "running_var"):
            val = rs.uniform(0.5, 1.5, shape).astype(np.float32)
        else:                                      # BN beta / running_mean
            val = rs.uniform(-0.17, 0.17, shape).astype(np.float32)
        sd[key] = val
    if model == "pointconv":
        for s in ("sa1", "sa2", "sa3"):
            sd["%s.densitynet.mlp_bns.2.bias" % s] = np.ones((1,), np.float32)
    return sd


def _kinds(model):
    return {p: (k, s) for p, k, s, _ in MODEL_SPECS[model]()}


def _is_bn(model, prefix):
    return _kinds(model)[prefix][0] == "bn"


def _fan_in_of(model, prefix, num_classes):
    return int(np.prod(_kinds(model)[prefix][1][1:]))


def make_orders(num_samples=1000, num_regions=32, seed=1):
    """Seed-replayed permutations, the call sequence of generate_all_orders
    (final_shapley_value.py:59-72 of the reference) after set_random(seed):
    one legacy np.random.permutation(arange(R)) per sample."""
    rs = np.random.RandomState(seed)
    return np.stack([rs.permutation(np.arange(0, num_regions, 1)) for _ in range(num_samples)]).astype(np.int64)


def make_pairs_and_contexts(num_pairs, num_regions=32, orders_m=(0, 1, 2, 3, 6, 9, 12, 15, 18, 21, 24, 27, 30),
                            max_contexts=100, seed=1):
    """Region pairs and contexts with the shapes of final_gen_pair.py:18-43,288-300 of the
    reference: pairs (P,2) with j>i, and per order m a context array (P, ctx, m) where
    ctx = min(C(R-2,m), max_contexts).  Sampling uses the legacy numpy stream."""
    from itertools import combinations
    from math import comb
    rs = np.random.RandomState(seed)
    all_pairs = np.array([[i, j] for i in range(num_regions) for j in range(num_regions) if j > i])
    pairs = all_pairs[rs.choice(all_pairs.shape[0], size=num_pairs, replace=False)].astype(np.int64)
    contexts = {}
    for m in orders_m:
        per_pair = []
        for (ri, rj) in pairs:
            rest = [r for r in range(num_regions) if r != ri and r != rj]
            if comb(len(rest), m) > max_contexts:
                per_pair.append([rs.choice(rest, m, replace=False) for _ in range(max_contexts)])
            else:
                per_pair.append(list(combinations(rest, m)))
        contexts[m] = np.array(per_pair)
        if m == 0:
            contexts[m] = contexts[m].reshape(num_pairs, 1, 0)
    return pairs, contexts
