"""Diagnostic for the end-to-end loop at N > 1 GPUs (launch with torchrun): per-rank step times of the e2e variants
(prefetch / synchronous ingest, with / without the per-frame NCCL pose all-gather), optionally with the process bound to
the CPUs next to its GPU before the pinned frames are allocated (NUMA=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib, time
import numpy as np


def bind_to_gpu_numa(local_rank):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
        cpus = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = ids & allowed
        if use:
            os.sched_setaffinity(0, use)
        return f"gpu {local_rank} bdf {bdf} numa {node} cpus {cpus} -> bound to {len(use)} cpus"
    except Exception as e:
        return f"numa binding failed: {e!r}"


def main():
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    note = bind_to_gpu_numa(local_rank) if os.environ.get("NUMA") == "1" else "no binding"
    import torch, torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pkg = importlib.import_module("3dobjecttracking_b200")
    capi = importlib.import_module("3dobjecttracking_b200.capi")
    nb = 128
    wl = pkg.synth.make_workload("c4", n_bodies=nb, n_divides=4, seed=0, first_body=rank * nb)
    stream = torch.cuda.current_stream(dev)
    ctx = capi.context_from_workload(wl, device=local_rank, stream=stream.cuda_stream)
    ctx.start_modalities(0); ctx.synchronize()
    h_color = torch.from_numpy(wl.color_frames).pin_memory()
    h_depth = torch.from_numpy(wl.depth_frames.view(np.uint8).reshape(nb, wl.depth_frames.shape[1], -1)).pin_memory()
    poses_np = torch.from_numpy(np.ascontiguousarray(wl.start_body2world.reshape(nb, 12))).pin_memory().numpy()
    out_poses = torch.empty((nb, 12), dtype=torch.float32).pin_memory()
    poses_dev = torch.empty((nb, 12), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * nb, 12), dtype=torch.float32, device=dev)
    n_corr, n_upd = wl.n_corr_iterations, wl.n_update_iterations

    def hand_over(prefetch):
        ctx.upload_batch_ptr(True, 0, nb, h_color.data_ptr(), h_color.stride(0), h_color.stride(1))
        ctx.upload_batch_ptr(False, 0, nb, h_depth.data_ptr(), h_depth.stride(0), h_depth.stride(1))
        if prefetch:
            ctx.prefetch_frames()

    def run(prefetch, gather, steps=20):
        times = []
        if prefetch:
            hand_over(True)
        for i in range(steps + 3):
            t0 = time.perf_counter()
            if not prefetch:
                hand_over(False)
            ctx.set_poses(poses_np)
            ctx.tracking_step(0, n_corr, n_upd)
            if prefetch:
                hand_over(True)
            ctx._ck(ctx.L.m3tb_get_poses(ctx.h, 0, nb, capi._p(out_poses.numpy())))
            if gather and world > 1:
                poses_dev.copy_(out_poses, non_blocking=True)
                dist.all_gather_into_tensor(gathered, poses_dev)
            if i >= 3:
                times.append(time.perf_counter() - t0)
        ctx.synchronize(); torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        return np.array(times) * 1e3

    def ingest_only(steps=10):
        times = []
        for i in range(steps):
            hand_over(False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.region_correspondences(0, 0)   # first consumer of the new frames: k_ingest + one small k_track launch
            e1.record(stream)
            torch.cuda.synchronize(dev)
            times.append(e0.elapsed_time(e1))
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        return np.array(times)

    res = {}
    res["ingest+corr0 (device ms)"] = ingest_only()
    for name, (p, g) in {"prefetch+gather": (True, True), "prefetch": (True, False), "sync ingest": (False, False), "sync ingest+gather": (False, True)}.items():
        res[name] = run(p, g)
    lines = [f"rank {rank} [{note}]"] + [f"   {k:28s} median {np.median(v):8.3f} ms  max {v.max():8.3f}" for k, v in res.items()]
    for r in range(world):
        if r == rank:
            print("\n".join(lines), flush=True)
        if world > 1:
            dist.barrier(device_ids=[local_rank])
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
