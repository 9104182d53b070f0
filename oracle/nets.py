"""ORACLE (test infrastructure, not product code).

torch fp32 CPU restatement of the five classifiers of the reference, written as
plain functions over a checkpoint-format state dict {key: tensor}.  Evaluation
mode only (running BN statistics, dropout = identity), which is the only mode
the coalition path uses (tools/final_util.py:261 of the reference).

Each function cites the reference lines it restates.  Discrete geometry
(in-model FPS, ball query, K=3 squared distances) goes through geom_oracle.c so
the decisions do not depend on the host BLAS.

Pinned against the reference's own nn.Modules by tests/golden/make_golden.py.
"""
import torch
import torch.nn.functional as F

from . import geom

EPS = 1e-5


def _t(sd, key):
    v = sd[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(v)


def _bn(x, sd, p):
    return F.batch_norm(x, _t(sd, p + ".running_mean"), _t(sd, p + ".running_var"), _t(sd, p + ".weight"),
                        _t(sd, p + ".bias"), False, 0.0, EPS)


def _conv(x, sd, p):
    w = _t(sd, p + ".weight")
    b = _t(sd, p + ".bias") if (p + ".bias") in sd else None
    return F.conv1d(x, w, b) if w.dim() == 3 else F.conv2d(x, w, b)


def _lin(x, sd, p):
    return F.linear(x, _t(sd, p + ".weight"), _t(sd, p + ".bias") if (p + ".bias") in sd else None)


# --------------------------------------------------------------------------- DGCNN / GCNN
def knn_indices(x, k):
    """models/dgcnn.py:12-18: top-k of -|xi|^2 + 2 xi.xj - |xj|^2 over j.  x (B,C,N)."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    return pd.topk(k=k, dim=-1)[1]


def edge_features(x, idx):
    """models/dgcnn.py:21-47: (B,C,N),(B,N,k) -> (B,2C,N,k) = [x_j - x_i ; x_i]."""
    B, C, N = x.shape
    k = idx.shape[-1]
    pts = x.transpose(2, 1).contiguous()                              # (B,N,C)
    flat = (idx + torch.arange(B).view(-1, 1, 1) * N).reshape(-1)
    nbr = pts.reshape(B * N, C)[flat].view(B, N, k, C)
    ctr = pts.view(B, N, 1, C).expand(B, N, k, C)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2)


def edgeconv_net(x, sd, k=20, dynamic=True):
    """models/dgcnn.py:89-120 (dynamic graph) and :159-194 (graph fixed on xyz).  x (B,3,N)."""
    B = x.shape[0]
    idx0 = knn_indices(x, k)
    feats = []
    h = x
    for i in range(1, 5):
        idx = knn_indices(h, k) if (dynamic and i > 1) else idx0
        e = edge_features(h, idx)
        e = F.leaky_relu(_bn(_conv(e, sd, "conv%d.0" % i), sd, "bn%d" % i), 0.2)
        h = e.max(dim=-1)[0]
        feats.append(h)
    h = torch.cat(feats, dim=1)
    h = F.leaky_relu(_bn(_conv(h, sd, "conv5.0"), sd, "bn5"), 0.2)
    g = torch.cat((F.adaptive_max_pool1d(h, 1).view(B, -1), F.adaptive_avg_pool1d(h, 1).view(B, -1)), 1)
    g = F.leaky_relu(_bn(_lin(g, sd, "linear1"), sd, "bn6"), 0.2)
    g = F.leaky_relu(_bn(_lin(g, sd, "linear2"), sd, "bn7"), 0.2)
    return _lin(g, sd, "linear3")


# --------------------------------------------------------------------------- PointNet
def _stn(x, sd, p, k):
    """models/pointnet.py:29-47: T-Net returning (B,k,k) = fc3(...) + I."""
    B = x.shape[0]
    h = F.relu(_bn(_conv(x, sd, p + "conv1"), sd, p + "bn1"))
    h = F.relu(_bn(_conv(h, sd, p + "conv2"), sd, p + "bn2"))
    h = F.relu(_bn(_conv(h, sd, p + "conv3"), sd, p + "bn3"))
    h = torch.max(h, 2, keepdim=True)[0].view(-1, 1024)
    h = F.relu(_bn(_lin(h, sd, p + "fc1"), sd, p + "bn4"))
    h = F.relu(_bn(_lin(h, sd, p + "fc2"), sd, p + "bn5"))
    h = _lin(h, sd, p + "fc3")
    h = h + torch.eye(k, dtype=torch.float32).flatten().view(1, k * k).repeat(B, 1)
    return h.view(-1, k, k)


def pointnet(x, sd):
    """models/pointnet.py:66-115 with feature_transform=True.  x (B,3,N) ->
    (logits (B,C), trans_feat (B,64,64), crt_points (B,1024) int64)."""
    trans = _stn(x, sd, "feat.stn.", 3)
    h = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1)
    h = F.relu(_bn(_conv(h, sd, "feat.conv1"), sd, "feat.bn1"))
    trans_feat = _stn(h, sd, "feat.fstn.", 64)
    h = torch.bmm(h.transpose(2, 1), trans_feat).transpose(2, 1)
    h = F.relu(_bn(_conv(h, sd, "feat.conv2"), sd, "feat.bn2"))
    h = _bn(_conv(h, sd, "feat.conv3"), sd, "feat.bn3")
    g, crt = torch.max(h, 2)
    g = g.view(-1, 1024)
    g = F.relu(_bn(_lin(g, sd, "fc1"), sd, "bn1"))
    g = F.relu(_bn(_lin(g, sd, "fc2"), sd, "bn2"))
    return _lin(g, sd, "fc3"), trans_feat, crt


# --------------------------------------------------------------------------- PointNet++ (MSG)
def _take(points, idx):
    """models/pointnet2.py:27-43: points (B,N,C), idx (B,...) -> (B,...,C)."""
    B = points.shape[0]
    bidx = torch.arange(B).view([B] + [1] * (idx.dim() - 1)).expand_as(idx)
    return points[bidx, idx, :]


def _fps(xyz, npoint):
    return torch.from_numpy(geom.fps(xyz.contiguous().numpy(), npoint))


def _sa_msg(xyz, feats, sd, p, npoint, radii, nsamples):
    """models/pointnet2.py:196-240.  xyz (B,3,N), feats (B,D,N)|None -> ((B,3,S), (B,D',S))."""
    pts = xyz.permute(0, 2, 1).contiguous()
    f = feats.permute(0, 2, 1) if feats is not None else None
    B, N, C = pts.shape
    new_xyz = _take(pts, _fps(pts, npoint))
    outs = []
    for b, (radius, K) in enumerate(zip(radii, nsamples)):
        gi = torch.from_numpy(geom.ball_query(radius, K, pts.numpy(), new_xyz.contiguous().numpy()))
        gx = _take(pts, gi) - new_xyz.view(B, npoint, 1, C)
        g = torch.cat([_take(f, gi), gx], dim=-1) if f is not None else gx
        g = g.permute(0, 3, 2, 1)
        j = 0
        while ("%s.conv_blocks.%d.%d.weight" % (p, b, j)) in sd:
            g = F.relu(_bn(_conv(g, sd, "%s.conv_blocks.%d.%d" % (p, b, j)), sd, "%s.bn_blocks.%d.%d" % (p, b, j)))
            j += 1
        outs.append(torch.max(g, 2)[0])
    return new_xyz.permute(0, 2, 1), torch.cat(outs, dim=1)


def pointnet2_msg(x, sd):
    """models/pointnet2.py:265-276 with sa3 = group-all (:120-136,156-178).  x (B,3,N)."""
    B = x.shape[0]
    l1_xyz, l1 = _sa_msg(x, None, sd, "sa1", 512, [0.1, 0.2, 0.4], [16, 32, 128])
    l2_xyz, l2 = _sa_msg(l1_xyz, l1, sd, "sa2", 128, [0.2, 0.4, 0.8], [32, 64, 128])
    g = torch.cat([l2_xyz.permute(0, 2, 1), l2.permute(0, 2, 1)], dim=-1)        # (B,128,643) xyz first
    g = g.unsqueeze(1).permute(0, 3, 2, 1)                                        # (B,643,128,1)
    for j in range(3):
        g = F.relu(_bn(_conv(g, sd, "sa3.mlp_convs.%d" % j), sd, "sa3.mlp_bns.%d" % j))
    g = torch.max(g, 2)[0].view(B, 1024)
    g = F.relu(_bn(_lin(g, sd, "fc1"), sd, "bn1"))
    g = F.relu(_bn(_lin(g, sd, "fc2"), sd, "bn2"))
    return _lin(g, sd, "fc3")


# --------------------------------------------------------------------------- PointConv
def _sqdist3(a, b):
    # geometry is always evaluated in fp32 (the reference's dtype); cast back for the float64 yardstick runs
    return torch.from_numpy(geom.square_distance3(a.contiguous().numpy(), b.contiguous().numpy())).to(a.dtype)


def _pointconv_sa(xyz, feats, sd, p, npoint, nsample, bandwidth, group_all):
    """models/pointconv.py:341-391.  xyz (B,3,N), feats (B,D,N)|None."""
    B, _, N = xyz.shape
    pts = xyz.permute(0, 2, 1).contiguous()
    f = feats.permute(0, 2, 1) if feats is not None else None
    # compute_density, models/pointconv.py:199-209
    dens = (torch.exp(-_sqdist3(pts, pts) / (2.0 * bandwidth * bandwidth)) / (2.5 * bandwidth)).mean(dim=-1)
    inv = (1.0 / dens).view(B, N, 1)
    if group_all:                                  # models/pointconv.py:148-170
        new_xyz = pts.mean(dim=1, keepdim=True)
        gx = pts.view(B, 1, N, 3) - new_xyz.view(B, 1, 1, 3)
        g = torch.cat([gx, f.view(B, 1, N, -1)], dim=-1) if f is not None else gx
        gd = inv.view(B, 1, N, 1)
        S = 1
    else:                                          # models/pointconv.py:117-145
        S = npoint
        new_xyz = _take(pts, _fps(pts, npoint))
        idx = torch.topk(_sqdist3(new_xyz, pts), nsample, dim=-1, largest=False, sorted=False)[1]
        gx = _take(pts, idx) - new_xyz.view(B, S, 1, 3)
        g = torch.cat([gx, _take(f, idx)], dim=-1) if f is not None else gx
        gd = _take(inv, idx)
    g = g.permute(0, 3, 2, 1)                                                  # (B,C+D,K,S)
    for j in range(3):
        g = F.relu(_bn(_conv(g, sd, "%s.mlp_convs.%d" % (p, j)), sd, "%s.mlp_bns.%d" % (p, j)))
    ds = (gd / gd.max(dim=2, keepdim=True)[0]).permute(0, 3, 2, 1)             # (B,1,K,S)
    for j in range(3):    # DensityNet applies ReLU after every layer (models/pointconv.py:226-235)
        ds = F.relu(_bn(_conv(ds, sd, "%s.densitynet.mlp_convs.%d" % (p, j)), sd, "%s.densitynet.mlp_bns.%d" % (p, j)))
    g = g * ds
    w = gx.permute(0, 3, 2, 1)                                                 # (B,3,K,S)
    for j in range(3):    # WeightNet, models/pointconv.py:257-265
        w = F.relu(_bn(_conv(w, sd, "%s.weightnet.mlp_convs.%d" % (p, j)), sd, "%s.weightnet.mlp_bns.%d" % (p, j)))
    agg = torch.matmul(g.permute(0, 3, 1, 2), w.permute(0, 3, 2, 1)).view(B, S, -1)
    out = _lin(agg, sd, p + ".linear")
    out = F.relu(_bn(out.permute(0, 2, 1), sd, p + ".bn_linear"))
    return new_xyz.permute(0, 2, 1), out


def pointconv(x, sd):
    """models/pointconv.py:413-424.  x (B,3,N)."""
    B = x.shape[0]
    l1_xyz, l1 = _pointconv_sa(x, None, sd, "sa1", 512, 32, 0.1, False)
    l2_xyz, l2 = _pointconv_sa(l1_xyz, l1, sd, "sa2", 128, 64, 0.2, False)
    _, l3 = _pointconv_sa(l2_xyz, l2, sd, "sa3", 1, None, 0.4, True)
    g = l3.view(B, 1024)
    g = F.relu(_bn(_lin(g, sd, "fc1"), sd, "bn1"))
    g = F.relu(_bn(_lin(g, sd, "fc2"), sd, "bn2"))
    return _lin(g, sd, "fc3")


def forward(model, x, sd, k=20):
    """Logits (B,C) of `model` in {pointnet, pointnet2, pointconv, dgcnn, gcnn, gcnn_adv}; x (B,3,N) fp32 CPU."""
    with torch.no_grad():
        x = x.contiguous()
        if model == "dgcnn":
            return edgeconv_net(x, sd, k, dynamic=True)
        if model in ("gcnn", "gcnn_adv"):
            return edgeconv_net(x, sd, k, dynamic=False)
        if model == "pointnet":
            return pointnet(x, sd)[0]
        if model == "pointnet2":
            return pointnet2_msg(x, sd)
        if model == "pointconv":
            return pointconv(x, sd)
        raise ValueError("unknown model %r" % model)
