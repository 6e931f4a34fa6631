"""CPU tests of the factor-sharded multi-GPU mode's host logic (world_size 2, gloo): the landmark / plane partition is a
bijection, and the shards' reduced systems — linearised independently (here by the oracle) and summed with ONE all-reduce —
equal the reduced system of the whole window, which is exactly what the GPU path relies on."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    from gf2_loader import load
    gf2 = load()
    synth = importlib.import_module("gf2_b200.synth"); shard = importlib.import_module("gf2_b200.shard")
    import gf2_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.make_windows(2, config_id=4, n_landmarks=150, wheel=True, n_planes=400)
    orc.imu_preintegrate(w); orc.wheel_preintegrate(w)
    opts = gf2.abi.default_opts()
    mine = shard.shard_windows(w, rank, world)
    empty = shard.shard_windows(w, 0, 10 ** 6)  # no landmarks beyond l = 0 ... build a truly empty shard below
    empty["n_landmarks"][:] = 0
    if "n_planes" in empty:
        empty["n_planes"][:] = 0
    err = 0.0
    for i in range(2):
        S_r, g_r, c_r, _, _ = orc.linearize_window(mine, i, opts)
        S_nv, g_nv, c_nv, _, _ = orc.linearize_window(empty, i, opts)     # IMU + wheel + prior only
        vis = torch.from_numpy(np.concatenate([(S_r - S_nv).ravel(), g_r - g_nv, [c_r - c_nv]]))
        dist.all_reduce(vis)                                               # the one collective per linearisation
        S_full, g_full, c_full, _, _ = orc.linearize_window(w, i, opts)
        D = S_full.shape[0]
        S_sum = vis[:D * D].numpy().reshape(D, D) + S_nv; g_sum = vis[D * D:D * D + D].numpy() + g_nv; c_sum = float(vis[-1]) + c_nv
        err = max(err, np.abs(S_sum - S_full).max() / np.abs(S_full).max(), np.abs(g_sum - g_full).max() / np.abs(g_full).max(), abs(c_sum - c_full) / c_full)
    counts = torch.tensor([int(mine["n_landmarks"].sum()), int(mine["n_planes"].sum())])
    dist.all_reduce(counts)
    if rank == 0:
        q.put((err, counts.tolist(), int(w["n_landmarks"].sum()), int(w["n_planes"].sum())))
    dist.destroy_process_group()


def test_shards_sum_to_the_full_reduced_system_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, counts, nl, npl = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert counts == [nl, npl]          # partition covers every landmark / plane exactly once
    assert err < 1e-11, err


def test_partition_is_a_bijection_and_keeps_observations(gf2):
    synth = importlib.import_module("gf2_b200.synth"); shard = importlib.import_module("gf2_b200.shard")
    w = synth.make_windows(1, n_landmarks=101)
    parts = [shard.shard_windows(w, r, 3) for r in range(3)]
    assert sum(int(p["n_landmarks"][0]) for p in parts) == 101
    lam = shard.gather_landmarks(w["n_landmarks"], [p["inv_depth"] for p in parts], 3)
    assert np.array_equal(lam[0, :101], w["inv_depth"][0, :101])
    # observation multiset preserved
    allobs = np.concatenate([p["obs"][0][:int(p["track_len"][0].sum())] for p in parts])
    ref = w["obs"][0][:int(w["track_len"][0].sum())]
    assert sorted(map(tuple, allobs.tolist())) == sorted(map(tuple, ref.tolist()))
