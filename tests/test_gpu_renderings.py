"""SURVEY f4 on the device: modeled occlusion handling, region checking and silhouette checking (the checks that read
FocusedDepthRenderer / FocusedSilhouetteRenderer images, handed over with m3tb_upload_*_rendering) must give the
oracle's per-line / per-point records bit for bit, the same histograms, and poses within the per-iteration gate."""
import copy

import numpy as np
import pytest

from helpers import assert_lines_bit_equal, assert_points_bit_equal, pose_error

pytestmark = pytest.mark.gpu


def _workload(synth, which):
    wl = synth.make_workload("c2", n_bodies=5, n_divides=3, seed=19)
    synth.fill_depth_offsets(wl.region_model)
    synth.fill_depth_offsets(wl.depth_model)
    synth.add_renderings(wl, occluder_bodies=(1, 3))
    wl.region, wl.depth = copy.copy(wl.region), copy.copy(wl.depth)
    if which in ("all", "region"):
        wl.region.model_occlusions = True
        wl.region.use_region_checking = True
        wl.region.n_unoccluded_iterations = 0
    if which in ("all", "depth"):
        wl.depth.model_occlusions = True
        wl.depth.use_silhouette_checking = True
        wl.depth.n_unoccluded_iterations = 0
    if which == "fallback":  # too few survivors: the second pass without occlusion handling
        wl.region.model_occlusions = True
        wl.region.n_unoccluded_iterations = 0
        wl.region.min_n_unoccluded_lines = 150
        wl.depth.model_occlusions = True
        wl.depth.n_unoccluded_iterations = 0
        wl.depth.min_n_unoccluded_points = 190
    return wl


@pytest.mark.parametrize("which", ["all", "region", "depth", "fallback"])
def test_renderer_image_checks_bit_exact(capi, oracle, synth, which):
    wl = _workload(synth, which)
    ctx = capi.context_from_workload(wl)
    orc = oracle.OracleTracker(wl, rotation_mode=oracle.ROTATION_LINEAR, exp_mode=oracle.EXP_RODRIGUES)
    base = oracle.OracleTracker(synth.make_workload("c2", n_bodies=5, n_divides=3, seed=19),
                                rotation_mode=oracle.ROTATION_LINEAR, exp_mode=oracle.EXP_RODRIGUES)
    orc.start_modalities(0)
    ctx.start_modalities(0)
    nb = wl.region.n_histogram_bins
    for b in range(wl.n_bodies):
        hf, hb = ctx.get_histograms(b, nb)
        assert np.array_equal(hf.view(np.uint32), orc.hist_f[b].view(np.uint32)), (which, b)
        assert np.array_equal(hb.view(np.uint32), orc.hist_b[b].view(np.uint32)), (which, b)
    rejected = 0
    for corr in range(wl.n_corr_iterations):
        ctx.set_poses(orc.get_poses())
        base.set_poses(orc.get_poses())
        ctx.corr_iteration(0, corr, wl.n_update_iterations)
        gpu = ctx.get_poses()
        for b in range(wl.n_bodies):
            n, view = orc.region_correspondences(b, 0, corr)
            assert ctx.get_closest_views(b)[0] == view
            assert_lines_bit_equal(ctx.get_region_lines(b, wl.lines_per_body), orc.lines[b][:n])
            m, view = orc.depth_correspondences(b, 0, corr)
            assert ctx.get_closest_views(b)[1] == view
            assert_points_bit_equal(ctx.get_depth_points(b, wl.points_per_body), orc.points[b][:m])
            n0, _ = base.region_correspondences(b, 0, corr)
            m0, _ = base.depth_correspondences(b, 0, corr)
            rejected += int(base.lines[b][:n0]["valid"].sum() - orc.lines[b][:n]["valid"].sum())
            rejected += int(base.points[b][:m0]["valid"].sum() - orc.points[b][:m]["valid"].sum())
        orc.tracking_step(0, n_corr=corr + 1, corr_begin=corr)
        dt, dr = pose_error(gpu, orc.get_poses())
        assert dt.max() < 1e-5 and dr.max() < 1e-5, (which, corr, dt, dr)
    if which != "fallback":
        assert rejected > 100, rejected     # the checks really rejected lines / points on the occluded bodies
    ctx.calculate_results(0)
    orc.calculate_results(0)
    for b in range(wl.n_bodies):
        hf, hb = ctx.get_histograms(b, nb)
        assert np.array_equal(hf.view(np.uint32), orc.hist_f[b].view(np.uint32)), (which, b)
        assert np.array_equal(hb.view(np.uint32), orc.hist_b[b].view(np.uint32)), (which, b)
    ctx.close()


def test_rendering_upload_errors(capi, synth):
    wl = synth.make_workload("c2", n_bodies=1, n_divides=2, seed=1)
    synth.add_renderings(wl)
    ctx = capi.context_from_workload(wl)
    r = copy.copy(wl.renderings[0]["region_depth"])
    r.scale = 0.0
    with pytest.raises(capi.M3TBError):
        ctx.upload_rendering(0, "region_depth", r)
    with pytest.raises(capi.M3TBError):
        ctx.upload_rendering(5, "region_depth", wl.renderings[0]["region_depth"])
    ctx.close()
