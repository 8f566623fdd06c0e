"""Times one config-5 tracking step (32 chains x 8 links x (300 + 300)) on the cluster-fused path and on the general
multi-launch path (default) - M3TB_CLUSTER=1 selects the former, CUDA events, device-resident frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib, os, sys
import numpy as np
import torch
synth = importlib.import_module("3dobjecttracking_b200.synth")
capi = importlib.import_module("3dobjecttracking_b200.capi")
n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for variant in ("projected", "constrained"):
    wl = synth.make_chain_workload(n_chains=n_chains, n_links=8, n_lines=300, n_points=300, n_divides=4, variant=variant, seed=0)
    poses = np.ascontiguousarray(wl.start_body2world.reshape(-1, 12))
    for no_cluster in ("0", "1"):
        os.environ["M3TB_CLUSTER"] = "0" if no_cluster == "1" else "1"
        stream = torch.cuda.current_stream()
        ctx = capi.context_from_workload(wl, stream=stream.cuda_stream)
        ctx.start_modalities(0)
        ts = []
        for i in range(13):
            ctx.set_poses(poses); ctx.reset_joint_poses()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.tracking_step(0, wl.n_corr_iterations, wl.n_update_iterations); e1.record(stream)
            torch.cuda.synchronize()
            if i >= 3: ts.append(e0.elapsed_time(e1))
        print(f"{variant:12s} {'multi-launch' if no_cluster == '1' else 'cluster-fused'}: {np.mean(ts):.3f} ms/step (min {np.min(ts):.3f})", flush=True)
        ctx.close()
