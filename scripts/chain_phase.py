"""Phase clocks of the cluster-fused chain path (M3TB_TIMING=1): per corr iteration views / region / depth, per update
accumulate / barrier / cluster solve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib, os, sys
import numpy as np
os.environ["M3TB_TIMING"] = "1"
synth = importlib.import_module("3dobjecttracking_b200.synth")
capi = importlib.import_module("3dobjecttracking_b200.capi")
variant = sys.argv[1] if len(sys.argv) > 1 else "projected"
n_chains = int(sys.argv[2]) if len(sys.argv) > 2 else 16
wl = synth.make_chain_workload(n_chains=n_chains, n_links=8, n_lines=300, n_points=300, n_divides=4, variant=variant, seed=0)
ctx = capi.context_from_workload(wl)
ctx.start_modalities(0)
for _ in range(3):
    ctx.set_poses(wl.start_body2world); ctx.reset_joint_poses()
    ctx.tracking_step(0, wl.n_corr_iterations, wl.n_update_iterations)
ctx.synchronize()
for body in (0, 3, 8 * (n_chains - 1)):
    c = ctx.phase_clocks(body, 256)
    c = c[c > 0]
    d = np.diff(c)
    print(f"{variant} body {body}: total {c[-1]-c[0]} cycles; prologue {d[0]}; n stamps {len(c)}")
    per = d[1:1 + 9 * wl.n_corr_iterations].reshape(-1, 9)
    print("   " + " ".join(f"{l:>8}" for l in ["views", "region", "depth", "acc0", "bar0", "solve0", "acc1", "bar1", "solve1"]))
    for row in per:
        print("   " + " ".join(f"{v:8d}" for v in row))
