"""Multi-GPU host logic on CPU (gloo, world_size 2): bodies are sharded in contiguous blocks with no data-path
collective; the solved poses are all-gathered once per frame. Each rank runs its shard (the oracle stands
in for the device here - this tests the sharding / gather plumbing, not the kernels) and the gathered result must equal
the single-process run bit for bit, because a body's data depends only on (seed, global body index)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, out_path):
    import importlib
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_py
    pkg = importlib.import_module("3dobjecttracking_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = pkg.sharding.shard_bounds(n_total, rank, world)
    wl = pkg.synth.make_workload("c2", n_bodies=count, n_divides=2, seed=7, first_body=first)
    trk = oracle_py.OracleTracker(wl)
    trk.start_modalities(0)
    trk.tracking_step(0)
    local = torch.from_numpy(trk.get_poses().copy())
    counts = [pkg.sharding.shard_bounds(n_total, r, world)[1] for r in range(world)]
    gathered = pkg.sharding.all_gather_poses(local, counts=counts)
    if rank == 0:
        np.save(out_path, gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything(pkg):
    sb = pkg.sharding.shard_bounds
    for n in (0, 1, 5, 128, 1024, 1027):
        for world in (1, 2, 3, 8):
            blocks = [sb(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        sb(4, 2, 2)


@pytest.mark.parametrize("n_total", [4, 5])
def test_two_rank_gloo_run_equals_single_process(pkg, oracle, tmp_path, n_total):
    import torch.multiprocessing as mp
    out = str(tmp_path / "gathered.npy")
    port = 29500 + (os.getpid() % 500) + n_total
    mp.spawn(_worker, args=(2, port, n_total, out), nprocs=2, join=True)
    gathered = np.load(out)
    wl = pkg.synth.make_workload("c2", n_bodies=n_total, n_divides=2, seed=7)
    trk = oracle.OracleTracker(wl)
    trk.start_modalities(0)
    trk.tracking_step(0)
    assert gathered.shape == (n_total, 3, 4)
    assert np.array_equal(gathered.view(np.uint32), trk.get_poses().view(np.uint32))
