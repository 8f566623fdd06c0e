"""Measured occlusion handling (region_modality.cpp:1343-1389, depth_modality.cpp:736-776, the two-pass
logic of CalculateCorrespondences) on the oracle: behaviour checks on a synthetic occluder."""
import copy
import importlib

import numpy as np
import pytest

synth = importlib.import_module("3dobjecttracking_b200.synth")


def occluded_workload(n_bodies=2, seed=9, region_only=False, **kw):
    wl = synth.make_workload("c2", n_bodies=n_bodies, n_divides=3, seed=seed)
    synth.fill_depth_offsets(wl.region_model, seed)
    synth.fill_depth_offsets(wl.depth_model, seed)
    for b in range(n_bodies):
        synth.add_occluder(wl, b, side="left" if b % 2 == 0 else "top", seed=seed)
    wl.region.measure_occlusions = True
    wl.region.n_unoccluded_iterations = 0
    wl.depth.measure_occlusions = True
    wl.depth.n_unoccluded_iterations = 0
    for k, v in kw.items():
        setattr(wl.region if hasattr(wl.region, k) else wl.depth, k, v)
    if region_only:
        wl.depth = None
        wl.depth_model = None
    return wl


def test_occluded_lines_and_points_are_dropped(oracle):
    wl = occluded_workload()
    base = copy.deepcopy(wl)
    base.region.measure_occlusions = False
    base.depth.measure_occlusions = False
    t_occ, t_base = oracle.OracleTracker(wl), oracle.OracleTracker(base)
    for t in (t_occ, t_base):
        t.set_poses(wl.gt_body2world)
        t.start_modalities(0)
    for b in range(wl.n_bodies):
        n0, _ = t_base.region_correspondences(b, 0, 0)
        n1, _ = t_occ.region_correspondences(b, 0, 0)
        v0, v1 = t_base.lines[b]["valid"][:n0], t_occ.lines[b]["valid"][:n1]
        assert n0 == n1 and v1.sum() < 0.85 * v0.sum() and v1.sum() > 0.2 * v0.sum(), (v0.sum(), v1.sum())
        assert not (v1 & ~v0).any()          # occlusion handling only removes lines
        # the dropped lines sit under the occluder: their centres project into its image rectangle
        dropped = (v0 == 1) & (v1 == 0)
        cu, cv = t_base.lines[b]["center_u"][:n0][dropped], t_base.lines[b]["center_v"][:n0][dropped]
        kept_cu = t_base.lines[b]["center_u"][:n0][v1 == 1]
        kept_cv = t_base.lines[b]["center_v"][:n0][v1 == 1]
        if b % 2 == 0:
            assert cu.max() < kept_cu.max() and np.median(cu) < np.median(kept_cu)
        else:
            assert cv.max() < kept_cv.max() and np.median(cv) < np.median(kept_cv)
        m0, _ = t_base.depth_correspondences(b, 0, 0)
        m1, _ = t_occ.depth_correspondences(b, 0, 0)
        p0, p1 = t_base.points[b]["valid"][:m0], t_occ.points[b]["valid"][:m1]
        assert p1.sum() < 0.9 * p0.sum() and not (p1 & ~p0).any()
    # histograms: the occluder's colour stays out of the foreground histogram when occlusions are handled
    nb = wl.region.n_histogram_bins
    occ_bin = (60 >> 4) * nb * nb + (170 >> 4) * nb + (70 >> 4)
    assert t_occ.hist_f[0][occ_bin] < 0.25 * t_base.hist_f[0][occ_bin]


def test_unoccluded_iterations_and_fallback_pass(oracle):
    """Before n_unoccluded_iterations have passed, and when fewer than min_n_unoccluded lines survive, the result is
    exactly the one without occlusion handling (region_modality.cpp:435-463)."""
    base = occluded_workload()
    base.region.measure_occlusions = False
    base.depth.measure_occlusions = False
    t_base = oracle.OracleTracker(base)
    t_base.set_poses(base.gt_body2world)
    t_base.start_modalities(0)
    t_base.region_correspondences(0, 0, 0)
    t_base.depth_correspondences(0, 0, 0)
    for kw in (dict(n_unoccluded_iterations=10), dict(min_n_unoccluded_lines=10000, min_n_unoccluded_points=10000)):
        wl = occluded_workload()
        for k, v in kw.items():
            if hasattr(wl.region, k):
                setattr(wl.region, k, v)
            if hasattr(wl.depth, k):
                setattr(wl.depth, k, v)
        t = oracle.OracleTracker(wl)
        t.set_poses(wl.gt_body2world)
        t.hist_f[:] = t_base.hist_f
        t.hist_b[:] = t_base.hist_b
        t.region_correspondences(0, 0, 0)
        t.depth_correspondences(0, 0, 0)
        assert np.array_equal(t.lines[0].view(np.uint8), t_base.lines[0].view(np.uint8)), kw
        assert np.array_equal(t.points[0].view(np.uint8), t_base.points[0].view(np.uint8)), kw
    # ... and from iteration 10 on the default n_unoccluded_iterations = 10 switches the handling on
    wl = occluded_workload(n_unoccluded_iterations=10)
    wl.depth.n_unoccluded_iterations = 10
    t = oracle.OracleTracker(wl)
    t.set_poses(wl.gt_body2world)
    t.start_modalities(0)
    n, _ = t.region_correspondences(0, 10, 0)
    assert t.lines[0]["valid"][:n].sum() < 0.85 * t_base.lines[0]["valid"][:n].sum()


def test_tracking_with_occluder_converges(oracle):
    """A third of the body hidden: the tracker still pulls the pose in (translation clearly, rotation is weakly
    constrained by what remains visible of the prism)."""
    from helpers import pose_error
    wl = synth.make_workload("c2", n_bodies=8, n_divides=3, seed=9)
    synth.fill_depth_offsets(wl.region_model, 9)
    synth.fill_depth_offsets(wl.depth_model, 9)
    for b in range(8):
        synth.add_occluder(wl, b, side="left" if b % 2 == 0 else "top", cover=0.3, seed=9)
    for m in (wl.region, wl.depth):
        m.measure_occlusions = True
        m.n_unoccluded_iterations = 0
    t = oracle.OracleTracker(wl)
    t.start_modalities(0)
    e0t, e0r = pose_error(wl.start_body2world, wl.gt_body2world)
    for it in range(3):
        t.tracking_step(it)
        t.calculate_results(it)
    e1t, e1r = pose_error(t.get_poses(), wl.gt_body2world)
    assert np.median(e1t) < 0.5 * np.median(e0t) and np.median(e1r) < 0.8 * np.median(e0r), (e0t, e1t, e0r, e1r)
