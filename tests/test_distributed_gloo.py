"""CPU, world_size 2 over gloo: the sharding + statistics-gather logic of the multi-GPU path.
Each rank steps its shard of environments with the host build of the device core (test-only) and
the gathered shards must equal a single-process run bit for bit (results depend only on
(seed, global env id, step))."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_narde_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, steps, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import support
    hs = support.HostSim()
    base, n = D.shard_range(total, rank, world)
    lo, hi = hs.reset(n, env_base=base, seed=seed, step=0)
    stats = np.zeros(8, np.int64)
    for t in range(1, steps + 1):
        o = hs.step_full(lo, hi, env_base=base, seed=seed, step=t, cap=16, flags=2, want_obs=False)
        s = o["stats"]
        stats[[0, 1, 2, 3, 4, 5, 7]] += s[[0, 1, 2, 3, 4, 5, 7]]
        stats[6] = max(stats[6], s[6])
    # policy-weight broadcast from rank 0 (NCCL on GPUs): every rank ends with rank 0's tensors
    w = [torch.full((1000,), float(rank + 1)), torch.arange(16, dtype=torch.int16) * (rank + 1)]
    D.broadcast_policy(w, src=0)
    assert (w[0] == 1.0).all() and (w[1] == torch.arange(16, dtype=torch.int16)).all()
    allst = D.gather_stats(torch.from_numpy(stats))
    tmax = D.max_over_ranks(torch.tensor([float(rank + 1)], dtype=torch.float64))
    planes = [torch.zeros((D.shard_range(total, r, world)[1], 32), dtype=torch.uint8) for r in range(world)]
    dist.all_gather(planes, torch.from_numpy(np.concatenate([lo, hi], axis=1)))
    if rank == 0:
        np.save(os.path.join(out_dir, "planes.npy"), torch.cat(planes).numpy())
        np.save(os.path.join(out_dir, "stats.npy"), allst.numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), tmax.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_tiles_the_env_ids():
    for total, world in ((1048576, 8), (10, 3), (7, 8), (4096, 1)):
        nxt = 0
        for r in range(world):
            base, n = D.shard_range(total, r, world)
            assert base == nxt
            nxt += n
        assert nxt == total
    assert D.shard_range(1048576, 3, 8) == (3 * 131072, 131072)
    with pytest.raises(ValueError):
        D.shard_range(8, 2, 2)


def test_two_rank_sharded_selfplay_equals_single_process(tmp_path):
    import support
    total, steps, seed = 300, 40, 0x5EED
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, steps, seed, str(tmp_path)), nprocs=2, join=True)
    planes = np.load(tmp_path / "planes.npy")
    allst = np.load(tmp_path / "stats.npy")
    hs = support.HostSim()
    lo, hi = hs.reset(total, env_base=0, seed=seed, step=0)
    stats = np.zeros(8, np.int64)
    for t in range(1, steps + 1):
        o = hs.step_full(lo, hi, env_base=0, seed=seed, step=t, cap=16, flags=2, want_obs=False)
        stats[[0, 1, 2, 3, 4, 5, 7]] += o["stats"][[0, 1, 2, 3, 4, 5, 7]]
        stats[6] = max(stats[6], o["stats"][6])
    assert (planes == np.concatenate([lo, hi], axis=1)).all()
    merged = allst.sum(0)
    merged[6] = allst[:, 6].max()
    assert (merged == stats).all()
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0
