"""Diagnostic: free-running divergence GPU vs oracle on config-5 chains, per correspondence iteration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as oracle
from helpers import pose_error
synth = importlib.import_module("3dobjecttracking_b200.synth")
capi = importlib.import_module("3dobjecttracking_b200.capi")
variant = sys.argv[1] if len(sys.argv) > 1 else "projected"
wl = synth.make_chain_workload(n_chains=4, n_links=8, n_lines=300, n_points=300, n_divides=4, seed=6, variant=variant)
ctx = capi.context_from_workload(wl)
orc = oracle.OracleTracker(wl)
orc.start_modalities(0); ctx.start_modalities(0)
np.set_printoptions(precision=2, linewidth=200)
for it in range(2):
    for corr in range(wl.n_corr_iterations):
        ctx.corr_iteration(it, corr, wl.n_update_iterations)
        orc.tracking_step(it, n_corr=corr + 1, corr_begin=corr)
        dt, dr = pose_error(ctx.get_poses(), orc.get_poses())
        et, er = pose_error(orc.get_poses(), wl.gt_body2world)
        print(it, corr, "gpu-vs-oracle dt max %.2e med %.2e dr max %.2e med %.2e | oracle-vs-gt t %.2e r %.2e" % (dt.max(), np.median(dt), dr.max(), np.median(dr), np.median(et), np.median(er)))
    ctx.calculate_results(it); orc.calculate_results(it)
print("views equal:", [ctx.get_closest_views(b) == (orc.bodies[b].region_view, orc.bodies[b].depth_view) for b in range(8)])
