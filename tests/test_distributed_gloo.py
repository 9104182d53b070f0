"""Multi-rank host logic on CPU (gloo, world_size 2): permutation / pair sharding + the single allreduce.
The per-rank partial sums come from the reference's golden logits (no GPU needed); the GPU ranks plug
tools.final_common.shapley_partial_sums into the same functions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from interpret_quality_b200 import distributed as iqd
from interpret_quality_b200 import synthetic

R, LBL = 32, 3
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _reward(logits):
    others = [c for c in range(logits.shape[1]) if c != LBL]
    return logits[:, LBL] - torch.logsumexp(logits[:, others], dim=1)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(GOLDEN, "pointnet.npz"))
        logits = torch.from_numpy(g["shapley_logits"])
        nperm = int(g["shapley_nperm"])
        orders = synthetic.make_orders(16, R)[:nperm]
        v = _reward(logits).numpy().reshape(nperm, R + 1)

        def partial(order_slice):
            lo, _ = iqd.shard_range(nperm, rank, world)
            phi = np.zeros(R)
            for i, order in enumerate(order_slice):
                phi[order] += v[lo + i][1:] - v[lo + i][:-1]
            return torch.from_numpy(phi)

        phi = iqd.shapley_values_sharded(partial, orders, nperm, 1)
        # interactions: pairs sharded, disjoint slabs combined by one allreduce
        il = torch.from_numpy(g["inter_logits_m3"])
        P, rows, C = il.shape

        def logits_fn(sl):
            z = torch.zeros_like(il)
            z[sl[0]:sl[1]] = il[sl[0]:sl[1]]
            return z

        def reduce_fn(z):
            vv = _reward(z.reshape(-1, C)).reshape(P, rows // 4, 4).double()
            return vv[..., 0] + vv[..., 3] - vv[..., 1] - vv[..., 2]

        inter = iqd.interactions_sharded(logits_fn, reduce_fn, P)
        if rank == 0:
            np.savez(out, phi=phi, inter=inter)
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 100, 101):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = iqd.shard_range(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_world_size_one_needs_no_process_group():
    phi = iqd.shapley_values_sharded(lambda o: torch.ones(R, dtype=torch.float64) * len(o), np.zeros((10, R)), 10, 5)
    assert np.allclose(phi, 1.0)


def test_two_ranks_gloo_match_the_unsharded_reference(tmp_path):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    g = np.load(os.path.join(GOLDEN, "pointnet.npz"))
    assert np.abs(res["phi"] - g["shapley_phi"]).max() <= 1e-6 * np.abs(g["shapley_phi"]).max()
    assert np.abs(res["inter"] - g["inter_m3"]).max() <= 1e-5 * max(np.abs(g["inter_m3"]).max(), 1e-3)


def _pose_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seen = []

        def pose_fn(i):
            seen.append(i)
            return np.arange(R, dtype=np.float64) * (i + 1), torch.full((6, 4), float(i + 1))

        logged = []
        shap, logits = iqd.poses_sharded(pose_fn, 7, R, 6, 4, torch.device("cpu"),
                                         on_pose=lambda i, phi: logged.append(i))
        assert seen == list(range(rank, 7, world)) == logged        # round-robin deal, whole poses per rank
        assert shap.dtype == torch.float64 and logits.dtype == torch.float32
        np.savez(out % rank, shap=shap.numpy(), logits=logits.numpy())
    finally:
        dist.destroy_process_group()


def test_poses_dealt_round_robin_and_complete_on_every_rank(tmp_path):
    out = str(tmp_path / "pose%d.npz")
    mp.spawn(_pose_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    want_shap = np.arange(R)[None, :] * np.arange(1, 8)[:, None]
    for rank in range(2):
        res = np.load(out % rank)
        assert np.array_equal(res["shap"], want_shap)
        assert np.array_equal(res["logits"], np.broadcast_to(np.arange(1, 8, dtype=np.float32)[:, None, None], (7, 6, 4)))


def test_poses_world_size_one_is_the_serial_loop():
    shap, logits = iqd.poses_sharded(lambda i: (np.full(R, i, dtype=np.float64), torch.zeros(2, 3)), 3, R, 2, 3,
                                     torch.device("cpu"))
    assert np.array_equal(shap.numpy()[:, 0], [0.0, 1.0, 2.0]) and logits.shape == (3, 2, 3)
