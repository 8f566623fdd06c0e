"""Replays the reference's TrackerTest / RefinerTest pose optimisation on the oracle with a sparse set of regenerated
template views (shared by make_tracker_views.py and tests/test_reference_goldens.py)."""
import ctypes as C
import os
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class ReferenceReplay:
    def __init__(self, oracle, views):
        """views: {"region": {view_id: (points[200,38], contour_length)}, "depth": {view_id: (points[200,36], area)}}"""
        import reference_rig as rr
        self.oracle, self.rr = oracle, rr
        rig, ka = rr.rig(), rr.KA
        self.rig, self.ka = rig, ka
        ori = np.ascontiguousarray((-rr.geodesic_points()).astype(np.float32))
        nv = ori.shape[0]
        self.available = {k: set(v) for k, v in views.items()}
        self.near = 0.0
        self.keep = []

        def model(kind, fl):
            pts = np.zeros((nv, 200, fl), np.float32)
            scal = np.zeros(nv, np.float32)
            for vid, (p, s) in views[kind].items():
                pts[vid], scal[vid] = p, s
            m = SimpleNamespace(n_views=nv, n_points=200, orientations=ori, view_scalars=scal, points=pts,
                                stride_depth_offset=0.002, max_radius_depth_offset=0.05)
            self.keep.append(m)
            return oracle.make_model(m)

        self.rmodel, self.dmodel = model("region", 38), model("depth", 36)
        cc, dc = ka["color_camera"], ka["depth_camera"]
        self.cf = oracle.ColorFrame()
        self.cf.intrinsics = oracle.Intrinsics(cc["fu"], cc["fv"], cc["ppu"], cc["ppv"], cc["width"], cc["height"])
        self.cf.world2camera[:] = rig["color_w2c"][:3].astype(np.float32).reshape(12).tolist()
        self.cf.bgr, self.cf.pitch = rig["color"].ctypes.data, rig["color"].strides[0]
        self.df = oracle.DepthFrame()
        self.df.intrinsics = oracle.Intrinsics(dc["fu"], dc["fv"], dc["ppu"], dc["ppv"], dc["width"], dc["height"])
        self.df.world2camera[:] = rig["depth_w2c"][:3].astype(np.float32).reshape(12).tolist()
        self.df.depth, self.df.pitch, self.df.depth_scale = rig["depth"].ctypes.data, rig["depth"].strides[0], dc["depth_scale"]
        # both modalities measure occlusions in these tests (MeasureOcclusions); with n_unoccluded_iterations = 10 the
        # handling stays inactive at iteration 0, exactly as in the reference
        self.rp, self.dp = oracle.region_params(None), oracle.depth_params(None)
        self.rp.measure_occlusions = 1
        self.dp.measure_occlusions = 1
        n3 = 16 ** 3
        self.hf, self.hb = np.full(n3, 1.0 / n3, np.float32), np.full(n3, 1.0 / n3, np.float32)
        self.lines = np.zeros(200, oracle.REGION_LINE_DTYPE)
        self.points = np.zeros(200, oracle.DEPTH_POINT_DTYPE)
        self.body = (oracle.Body * 1)()
        B = self.body[0]
        B.body2world[:] = rig["body2world"][:3].astype(np.float32).reshape(12).tolist()
        B.region, B.region_model, B.color = C.pointer(self.rp), C.pointer(self.rmodel), C.pointer(self.cf)
        B.depth, B.depth_model, B.depth_frame = C.pointer(self.dp), C.pointer(self.dmodel), C.pointer(self.df)
        B.histogram_f, B.histogram_b = oracle.ptr(self.hf), oracle.ptr(self.hb)
        B.tikhonov_rotation, B.tikhonov_translation = 1000.0, 30000.0   # optimizer.h:52-53 (triangle_optimizer defaults)
        B.lines = self.lines.ctypes.data_as(C.POINTER(oracle.RegionLine))
        B.points = self.points.ctypes.data_as(C.POINTER(oracle.DepthPoint))
        B.region_occlusion_frame = C.pointer(self.df)
        self.L = oracle.lib()

    def pose(self):
        m = np.eye(4, dtype=np.float32)
        m[:3] = np.array(list(self.body[0].body2world), np.float32).reshape(3, 4)
        return m

    def _closest(self):
        """The views GetClosestView would select for the current pose, plus every view whose orientation dot product
        is within self.near of the winner's (a free-running replay on another arithmetic may take the runner-up)."""
        b2w = np.eye(4); b2w[:3] = self.pose()[:3]
        out = []
        ori = -self.rr.geodesic_points()
        for kind, w2c in (("region", self.rig["color_w2c"]), ("depth", self.rig["depth_w2c"])):
            b2c = w2c @ b2w
            t = b2c[:3, 3]
            dots = ori @ (b2c[:3, :3].T @ (t / np.linalg.norm(t)))
            best = int(np.argmax(dots))
            out.append((kind, best))
            for v in np.nonzero(dots >= dots[best] - self.near)[0]:
                if int(v) != best:
                    out.append((kind, int(v)))
        return out

    def _missing(self):
        return [(k, v) for k, v in self._closest() if v not in self.available[k]]

    def run(self, scenario, mirror=False, near=0.0):
        """Returns the list of (kind, view) that were needed but not available (empty: the replay is complete).
        mirror: the oracle's LINEAR / RODRIGUES arithmetic (what the CUDA path computes) instead of the reference's
        POLAR / PADE; near: also ask for the views within this dot-product margin of each selected view."""
        L, o = self.L, self.oracle
        self.near = near
        mode, exp = (o.ROTATION_LINEAR, o.EXP_RODRIGUES) if mirror else (o.ROTATION_POLAR, o.EXP_PADE)
        if scenario == "tracker":   # StartModalities(0); ExecuteTrackingStep(0): 7 x 2 (tracker_test.cpp:164-179)
            miss = self._missing()
            if miss:
                return miss
            L.orc_start_modalities(self.body, 1, 0, mode, 1)
            for corr in range(7):
                miss = self._missing()
                if miss:
                    return miss
                L.orc_tracking_step(self.body, 1, 0, corr, corr + 1, 2, mode, exp, 1, None)
                assert (self.body[0].region_view in self.available["region"]) and (self.body[0].depth_view in self.available["depth"])
            L.orc_calculate_results(self.body, 1, 0, mode, 1)
        else:                       # Refiner::ExecuteRefinementStep: 7 x (StartModalities ; correspondences ; 3 updates)
            for corr in range(7):
                miss = self._missing()
                if miss:
                    return miss
                L.orc_start_modalities(self.body, 1, 0, mode, 1)
                L.orc_tracking_step(self.body, 1, 0, corr, corr + 1, 3, mode, exp, 1, None)
                assert (self.body[0].region_view in self.available["region"]) and (self.body[0].depth_view in self.available["depth"])
        return []
