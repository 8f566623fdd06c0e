"""Debug: cycle stamps inside StructureSolveBlock (library built with M3TB_EXTRA_NVCC_FLAGS=-DM3TB_STRUCT_STAMPS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib, os, sys
import numpy as np
synth = importlib.import_module("3dobjecttracking_b200.synth")
capi = importlib.import_module("3dobjecttracking_b200.capi")
labels = ["jac+cons", "assembly", "pivots", "factor", "subst", "update"]
for variant in ("projected", "constrained"):
    wl = synth.make_chain_workload(n_chains=16, n_links=8, n_lines=300, n_points=300, n_divides=4, variant=variant, seed=0)
    for no_cluster in ("0", "1"):
        os.environ["M3TB_CLUSTER"] = "0" if no_cluster == "1" else "1"
        ctx = capi.context_from_workload(wl)
        ctx.start_modalities(0)
        for _ in range(2):
            ctx.set_poses(wl.start_body2world); ctx.reset_joint_poses()
            ctx.tracking_step(0, wl.n_corr_iterations, wl.n_update_iterations)
        th, _ = ctx.get_structure_theta(0, 128)
        raw = np.zeros(128, np.float32)
        import ctypes as C
        n, u = C.c_int(0), C.c_int(0)
        ctx._ck(ctx.L.m3tb_get_structure_theta(ctx.h, 0, capi._p(raw), 128, C.byref(n), C.byref(u)))
        st = raw[112:118]
        d = np.diff(np.concatenate([[0], st]))
        print(variant, "multi-launch" if no_cluster == "1" else "cluster", " ".join(f"{l}={int(v)}" for l, v in zip(labels, d)), "total", int(st[-1]), flush=True)
        ctx.close()
