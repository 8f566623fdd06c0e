"""Regenerates tests/golden/triangle_tracker_views.npz: every template view the oracle visits while it replays
TrackerTest.OptimizePoseMatrix and RefinerTest.OptimizePoseMatrix (M3T/test/tracker_test.cpp:164-179,
refiner_test.cpp:96-105) on the reference's frame pair. The reference generates its 2562-view models at test time with
OpenGL; here only the views that are actually selected are regenerated (reference_rig.py), on demand: run the loop, and
whenever GetClosestView picks a view that is not there yet, generate it and start over. Needs cv2 (build container)."""
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import reference_rig as rr   # noqa: E402
import oracle_py as oracle   # noqa: E402
from replay import ReferenceReplay  # noqa: E402


NEAR = 2e-4


def main():
    base = np.load(os.path.join(HERE, "triangle_test_view.npz"))
    views = {"region": {int(base["region_view"]): (base["region_points"], float(base["region_scalar"]))},
             "depth": {int(base["depth_view"]): (base["depth_points"], float(base["depth_scalar"]))}}
    pts = rr.geodesic_points()
    # the reference's arithmetic first, then the mirror arithmetic of the CUDA path with the near-tie neighbours of every
    # view on its way (NEAR: dot-product margin; neighbouring views are about 6e-3 apart at their closest)
    for scenario, mirror, near in (("tracker", False, 0.0), ("refiner", False, 0.0), ("tracker", True, NEAR), ("refiner", True, NEAR)):
        while True:
            rep = ReferenceReplay(oracle, views)
            missing = rep.run(scenario, mirror, near)
            if not missing:
                break
            for kind, v in missing:
                print(f"{scenario} mirror={mirror}: generating {kind} view {v}", flush=True)
                c2b = rr.camera2body_from_point(pts[v])
                p, s = (rr.region_view_points if kind == "region" else rr.depth_view_points)(c2b)
                views[kind][v] = (p, float(s))
        print(scenario, "final pose\n", rep.pose())
    out = {"orientations": base["orientations"]}
    for kind in ("region", "depth"):
        ids = sorted(views[kind])
        out[f"{kind}_ids"] = np.array(ids, np.int32)
        out[f"{kind}_points"] = np.stack([views[kind][i][0] for i in ids])
        out[f"{kind}_scalars"] = np.array([views[kind][i][1] for i in ids], np.float32)
    np.savez_compressed(os.path.join(HERE, "triangle_tracker_views.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
