"""GPU parity for measured occlusion handling: the occlusion window scans inside k_track (region lines and
depth points, two-pass fallback), k_histogram and the ROI ingest, against the oracle on frames with a synthetic
occluder. Integer control flow (which lines / points survive) and the per-line state are bit-exact."""
import copy

import numpy as np
import pytest

from helpers import assert_lines_bit_equal, assert_points_bit_equal, pose_error
from test_oracle_occlusion import occluded_workload

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", ["both", "region_only", "fallback", "late_start", "depth_scaling"])
def test_correspondences_with_occlusion_bit_exact(capi, oracle, variant):
    kw = {}
    if variant == "fallback":      # too few survivors: everything recomputed without occlusion handling
        kw = dict(min_n_unoccluded_lines=150, min_n_unoccluded_points=190)
    if variant == "late_start":
        kw = dict(n_unoccluded_iterations=3)
    wl = occluded_workload(n_bodies=4, region_only=(variant == "region_only"), **kw)
    if variant == "late_start":
        wl.depth.n_unoccluded_iterations = 3
    if variant == "depth_scaling":
        wl.depth.use_depth_scaling = True
        wl.depth.considered_distances = (0.08, 0.04, 0.02)
    ctx = capi.context_from_workload(wl)
    orc = oracle.OracleTracker(wl, rotation_mode=oracle.ROTATION_LINEAR, exp_mode=oracle.EXP_RODRIGUES)
    orc.start_modalities(0)
    ctx.start_modalities(0)
    nbins = wl.region.n_histogram_bins
    for b in range(wl.n_bodies):  # StartModality's histograms (occlusion handling active when n_unoccluded_iterations == 0)
        hf, hb = ctx.get_histograms(b, nbins)
        assert np.array_equal(hf.view(np.uint32), orc.hist_f[b].view(np.uint32))
        assert np.array_equal(hb.view(np.uint32), orc.hist_b[b].view(np.uint32))
    dropped = 0
    for iteration in (0, 3):
        for corr in (0, 2):
            ctx.set_poses(orc.get_poses())
            ctx.region_correspondences(iteration, corr)
            if wl.depth:
                ctx.depth_correspondences(iteration, corr)
            for b in range(wl.n_bodies):
                n, view = orc.region_correspondences(b, iteration, corr)
                lines = ctx.get_region_lines(b, wl.lines_per_body)
                assert ctx.get_closest_views(b)[0] == view
                assert_lines_bit_equal(lines, orc.lines[b][:n])
                dropped += int((lines["valid"] == 0).sum())
                if wl.depth:
                    m, _ = orc.depth_correspondences(b, iteration, corr)
                    assert_points_bit_equal(ctx.get_depth_points(b, wl.points_per_body), orc.points[b][:m])
    assert dropped > 0
    # CalculateResults' histogram update with occlusion handling
    orc.calculate_results(3)
    ctx.calculate_results(3)
    for b in range(wl.n_bodies):
        hf, hb = ctx.get_histograms(b, nbins)
        assert np.array_equal(hf.view(np.uint32), orc.hist_f[b].view(np.uint32))
        assert np.array_equal(hb.view(np.uint32), orc.hist_b[b].view(np.uint32))
    ctx.close()


def test_pose_parity_per_iteration_with_occlusion(capi, oracle):
    wl = occluded_workload(n_bodies=4)
    ctx = capi.context_from_workload(wl)
    orc = oracle.OracleTracker(wl, rotation_mode=oracle.ROTATION_POLAR, exp_mode=oracle.EXP_PADE)
    orc.start_modalities(0)
    ctx.start_modalities(0)
    for corr in range(wl.n_corr_iterations):
        ctx.set_poses(orc.get_poses())
        ctx.corr_iteration(0, corr, wl.n_update_iterations)
        orc.tracking_step(0, n_corr=corr + 1, corr_begin=corr)
        dt, dr = pose_error(ctx.get_poses(), orc.get_poses())
        assert dt.max() < 1e-4 and dr.max() < 1e-4, (corr, dt, dr)
    ctx.close()


def test_occlusion_with_pinned_roi_ingest(capi, oracle):
    """Pinned host frames + ROI ingest: a region-only body that measures occlusions needs its depth ROI too."""
    import torch
    wl = occluded_workload(n_bodies=3, region_only=True)
    ctx_a = capi.context_from_workload(wl)
    ctx_b = capi.context_from_workload(wl, upload_frames=False)
    hc = torch.from_numpy(wl.color_frames).pin_memory()
    hd = torch.from_numpy(wl.depth_frames.view(np.uint8).reshape(wl.n_bodies, wl.depth_frames.shape[1], -1)).pin_memory()
    ctx_b.upload_batch_ptr(True, 0, wl.n_bodies, hc.data_ptr(), hc.stride(0), hc.stride(1))
    ctx_b.upload_batch_ptr(False, 0, wl.n_bodies, hd.data_ptr(), hd.stride(0), hd.stride(1))
    for c in (ctx_a, ctx_b):
        c.start_modalities(0)
        c.tracking_step(0, wl.n_corr_iterations, wl.n_update_iterations)
    pa, pb = ctx_a.get_poses(), ctx_b.get_poses()
    assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32))
    assert ctx_b.last_ingest_bytes() > 0
    ctx_a.close()
    ctx_b.close()
