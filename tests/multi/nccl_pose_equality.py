"""Launched by tests/test_gpu_multi.py under torch.distributed.run with N ranks (one per GPU, NCCL):
every rank tracks its contiguous shard of a weak-scaled job through the C ABI, the solved poses are all-gathered
with NCCL (sharding.all_gather_poses), and rank 0 compares the gathered array bit for bit with a single-context run
of the whole job on its own GPU. Exit code 0 = equal."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("3dobjecttracking_b200")
    capi = importlib.import_module("3dobjecttracking_b200.capi")
    per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    which = sys.argv[2] if len(sys.argv) > 2 else "c2"
    wl = pkg.synth.make_workload(which, n_bodies=per_rank, n_divides=3, seed=9, first_body=rank * per_rank)
    ctx = capi.context_from_workload(wl, device=local)
    ctx.start_modalities(0)
    ctx.tracking_step(0, wl.n_corr_iterations, wl.n_update_iterations)
    poses = torch.from_numpy(ctx.get_poses().copy()).cuda(local)
    gathered = pkg.sharding.all_gather_poses(poses).cpu().numpy()
    ctx.close()
    ok = True
    if rank == 0:
        full = pkg.synth.make_workload(which, n_bodies=per_rank * world, n_divides=3, seed=9)
        c1 = capi.context_from_workload(full, device=local)
        c1.start_modalities(0)
        c1.tracking_step(0, full.n_corr_iterations, full.n_update_iterations)
        ref = c1.get_poses()
        c1.close()
        ok = gathered.shape == ref.shape and np.array_equal(gathered.view(np.uint32), ref.view(np.uint32))
        moved = float(np.abs(ref - full.start_body2world).max())
        print(f"nccl_pose_equality: world {world}, {per_rank} bodies/rank ({which}): "
              f"{'bit-identical' if ok else 'MISMATCH'}; max pose change vs start {moved:.3e}", flush=True)
        ok = ok and moved > 1e-4
    flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
