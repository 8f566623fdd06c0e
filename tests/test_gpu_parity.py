"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star / SURVEY §8c):
  * per-line / per-point state: bit-exact (integer control flow AND float payload) against the oracle in
    ROTATION_LINEAR mode (same expressions, -fmad=false vs -ffp-contract=off);
  * gradient / Hessian: <= 1e-5 of max|H| (summation order differs: warp tree vs serial);
  * pose after every correspondence iteration: <= 1e-4 rad and <= 1e-4 m against the reference-faithful oracle
    (polar rotation(), Pade exp).
"""
import numpy as np
import pytest

from helpers import assert_lines_bit_equal, assert_points_bit_equal, pose_error, rel_to_max

pytestmark = pytest.mark.gpu

TOL_POSE_M = 1e-4
TOL_POSE_RAD = 1e-4


@pytest.fixture(scope="module")
def wl_c2(synth):
    return synth.make_workload("c2", n_bodies=4, n_divides=4, seed=3)


@pytest.fixture(scope="module")
def wl_c3(synth):
    return synth.make_workload("c3", n_bodies=6, n_divides=4, seed=5)


def _setup(capi, oracle, wl, rotation_mode, exp_mode):
    ctx = capi.context_from_workload(wl)
    orc = oracle.OracleTracker(wl, rotation_mode=rotation_mode, exp_mode=exp_mode)
    return ctx, orc


def test_histograms_start_and_update_exact(capi, oracle, wl_c2):
    """StartModality / CalculateResults histogram side: counts are integers, blend is two roundings -> exact."""
    ctx, orc = _setup(capi, oracle, wl_c2, oracle.ROTATION_LINEAR, oracle.EXP_RODRIGUES)
    orc.start_modalities(0)
    ctx.start_modalities(0)
    nb = wl_c2.region.n_histogram_bins
    for b in range(wl_c2.n_bodies):
        hf, hb = ctx.get_histograms(b, nb)
        assert np.array_equal(hf.view(np.uint32), orc.hist_f[b].view(np.uint32))
        assert np.array_equal(hb.view(np.uint32), orc.hist_b[b].view(np.uint32))
        assert hf.sum() > 0.99 and hb.sum() > 0.99
    orc.calculate_results(0)
    ctx.calculate_results(0)
    for b in range(wl_c2.n_bodies):
        hf, hb = ctx.get_histograms(b, nb)
        assert np.array_equal(hf.view(np.uint32), orc.hist_f[b].view(np.uint32))
        assert np.array_equal(hb.view(np.uint32), orc.hist_b[b].view(np.uint32))
    ctx.close()


@pytest.mark.parametrize("which", ["c2", "c3"])
def test_fine_grained_calls_match_oracle(capi, oracle, wl_c2, wl_c3, which):
    """Modality-level API, call by call, for every corr / update iteration of one tracking step."""
    wl = wl_c2 if which == "c2" else wl_c3
    ctx, orc = _setup(capi, oracle, wl, oracle.ROTATION_LINEAR, oracle.EXP_RODRIGUES)
    orc.start_modalities(0)
    for b in range(wl.n_bodies):
        ctx.set_histograms(b, orc.hist_f[b], orc.hist_b[b])
    nl, npnt = wl.lines_per_body, wl.points_per_body
    for corr in range(wl.n_corr_iterations):
        # both sides start the iteration from the same poses (the oracle's)
        ctx.set_poses(orc.get_poses())
        if wl.region:
            ctx.region_correspondences(0, corr)
        if wl.depth:
            ctx.depth_correspondences(0, corr)
        for b in range(wl.n_bodies):
            if wl.region:
                n, view = orc.region_correspondences(b, 0, corr)
                assert ctx.get_closest_views(b)[0] == view
                assert_lines_bit_equal(ctx.get_region_lines(b, nl), orc.lines[b][:n])
            if wl.depth:
                n, view = orc.depth_correspondences(b, 0, corr)
                assert ctx.get_closest_views(b)[1] == view
                assert_points_bit_equal(ctx.get_depth_points(b, npnt), orc.points[b][:n])
        for upd in range(wl.n_update_iterations):
            ctx.set_poses(orc.get_poses())
            g_r = H_r = g_d = H_d = None
            if wl.region:
                g_r, H_r = ctx.region_gradient_hessian(0, corr, upd)
            if wl.depth:
                g_d, H_d = ctx.depth_gradient_hessian(0, corr, upd)
            ctx.calculate_optimization(0, corr, upd)
            gpu_poses = ctx.get_poses()
            for b in range(wl.n_bodies):
                g = np.zeros(6, np.float32)
                H = np.zeros((6, 6), np.float32)
                if wl.region:
                    og, oH = orc.region_gradient_hessian(b, corr, upd)
                    assert rel_to_max(H_r[b], oH) < 1e-5 and rel_to_max(g_r[b], og) < 1e-4
                    assert np.array_equal(H_r[b], H_r[b].T)
                    g, H = g + og, H + oH
                if wl.depth:
                    og, oH = orc.depth_gradient_hessian(b, corr)
                    assert rel_to_max(H_d[b], oH) < 1e-5 and rel_to_max(g_d[b], og) < 1e-4
                    g, H = g + og, H + oH
                ok, theta = orc.optimize(b, g, H)
                assert ok
            dt, dr = pose_error(gpu_poses, orc.get_poses())
            assert dt.max() < 1e-5 and dr.max() < 1e-5, (corr, upd, dt.max(), dr.max())  # logf ulp differences in local mode
    ctx.close()


@pytest.mark.parametrize("which", ["c2", "c3"])
def test_pose_parity_per_iteration(capi, oracle, wl_c2, wl_c3, which):
    """The contract gate (BASELINE.json north_star): pose after each correspondence iteration within
    1e-4 rad / 1e-4 m of the reference-faithful oracle (polar rotation(), Pade exp) ON IDENTICAL INPUTS:
    both sides enter every iteration with the oracle's pose. (Free-running trajectories are compared in
    the next test with a robust statistic: the path contains discrete events - closest-view switches,
    int() truncations of line coordinates - that a 1e-7 difference in summation order can flip.)"""
    wl = wl_c2 if which == "c2" else wl_c3
    ctx, orc = _setup(capi, oracle, wl, oracle.ROTATION_POLAR, oracle.EXP_PADE)
    orc.start_modalities(0)
    ctx.start_modalities(0)
    for corr in range(wl.n_corr_iterations):
        ctx.set_poses(orc.get_poses())
        ctx.corr_iteration(0, corr, wl.n_update_iterations)
        orc.tracking_step(0, n_corr=corr + 1, corr_begin=corr)
        dt, dr = pose_error(ctx.get_poses(), orc.get_poses())
        assert dt.max() < TOL_POSE_M and dr.max() < TOL_POSE_RAD, (corr, dt, dr)
    ctx.close()


@pytest.mark.parametrize("which", ["c2", "c3"])
def test_free_running_trajectories(capi, oracle, wl_c2, wl_c3, which):
    """Both sides free-run a whole tracking step from the same start. Almost all bodies must stay within the
    per-iteration tolerance for the whole step; a body that hits a discrete event may deviate, but only by the
    size of one flipped line / view switch (<< the 5 mm / 3 deg start perturbation)."""
    wl = wl_c2 if which == "c2" else wl_c3
    ctx, orc = _setup(capi, oracle, wl, oracle.ROTATION_POLAR, oracle.EXP_PADE)
    orc.start_modalities(0)
    ctx.start_modalities(0)
    worst_t = np.zeros(wl.n_bodies)
    worst_r = np.zeros(wl.n_bodies)
    for corr in range(wl.n_corr_iterations):
        ctx.corr_iteration(0, corr, wl.n_update_iterations)
        orc.tracking_step(0, n_corr=corr + 1, corr_begin=corr)
        dt, dr = pose_error(ctx.get_poses(), orc.get_poses())
        worst_t, worst_r = np.maximum(worst_t, dt), np.maximum(worst_r, dr)
    within = (worst_t < TOL_POSE_M) & (worst_r < TOL_POSE_RAD)
    assert within.mean() >= 0.5, (worst_t, worst_r)
    assert worst_t.max() < 3e-3 and worst_r.max() < 2e-2, (worst_t, worst_r)
    ctx.close()


def test_fused_step_equals_iteration_by_iteration(capi, oracle, wl_c2):
    """m3tb_tracking_step (one launch) == 7 x m3tb_corr_iteration, bit for bit; and it moves towards ground truth."""
    ctx_a = capi.context_from_workload(wl_c2)
    ctx_b = capi.context_from_workload(wl_c2)
    for c in (ctx_a, ctx_b):
        c.start_modalities(0)
    ctx_a.tracking_step(0, wl_c2.n_corr_iterations, wl_c2.n_update_iterations)
    for corr in range(wl_c2.n_corr_iterations):
        ctx_b.corr_iteration(0, corr, wl_c2.n_update_iterations)
    pa, pb = ctx_a.get_poses(), ctx_b.get_poses()
    assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32))
    dt0, dr0 = pose_error(wl_c2.start_body2world, wl_c2.gt_body2world)
    dt1, dr1 = pose_error(pa, wl_c2.gt_body2world)
    assert (dt1 < dt0).all() and (dr1 < dr0).all(), (dt0, dt1, dr0, dr1)
    ctx_a.close()
    ctx_b.close()


def test_full_cycle_two_frames(capi, oracle, wl_c2):
    """StartModalities + 2 x (tracking step + CalculateResults): poses within tolerance of the oracle."""
    ctx, orc = _setup(capi, oracle, wl_c2, oracle.ROTATION_POLAR, oracle.EXP_PADE)
    orc.start_modalities(0)
    ctx.start_modalities(0)
    for it in range(2):
        ctx.tracking_step(it, wl_c2.n_corr_iterations, wl_c2.n_update_iterations)
        ctx.calculate_results(it)
        orc.tracking_step(it)
        orc.calculate_results(it)
        dt, dr = pose_error(ctx.get_poses(), orc.get_poses())
        assert np.median(dt) < TOL_POSE_M and np.median(dr) < TOL_POSE_RAD, (it, dt, dr)
        assert dt.max() < 3e-3 and dr.max() < 2e-2, (it, dt, dr)  # discrete events, see test_free_running_trajectories
    ctx.close()


def test_errors_are_loud(capi):
    """No silent fallbacks: bad use returns an error status with a message."""
    ctx = capi.Context(0, 2, 2, 1)
    with pytest.raises(capi.M3TBError):
        ctx.tracking_step(0, 1, 1)  # nothing set up
    rp = capi.region_params()
    rp.function_length = 6
    with pytest.raises(capi.M3TBError):
        ctx.set_body(0, rp, None, None)
    ctx.close()


@pytest.mark.parametrize("rot_deg,trans_m", [(3.0, 0.005), (9.0, 0.03)])
def test_tiles_do_not_change_results(capi, synth, monkeypatch, rot_deg, trans_m):
    """Shared-memory ROI tiles are a pure staging optimisation: with tiles on / off (M3TB_NO_TILES=1) every pose
    is bit-identical, including when the pose moves so far (3 cm ~ 30 px) that samples leave the tile and are
    served by the global-memory fallback."""
    wl = synth.make_workload("c2", n_bodies=6, n_divides=4, seed=11, rot_deg=rot_deg, trans_m=trans_m)
    poses = []
    for no_tiles in ("0", "1"):
        monkeypatch.setenv("M3TB_NO_TILES", no_tiles)
        ctx = capi.context_from_workload(wl)
        ctx.start_modalities(0)
        ctx.tracking_step(0, wl.n_corr_iterations, wl.n_update_iterations)
        poses.append(ctx.get_poses())
        ctx.close()
    assert np.array_equal(poses[0].view(np.uint32), poses[1].view(np.uint32))


@pytest.mark.parametrize("rot_deg,trans_m", [(3.0, 0.005), (9.0, 0.03)])
def test_pinned_roi_ingest_matches_full_upload(capi, synth, rot_deg, trans_m):
    """Frame ingest: frames handed over in pinned memory are fetched ROI-only (zero-copy, k_ingest)
    instead of being copied in full. Results must be bit-identical to the full-copy path, also when the pose moves
    far enough (3 cm ~ 30 px) that samples leave the ROI and are read from the pinned frame directly, and over a
    whole cycle (StartModalities, tracking step, CalculateResults, next frame)."""
    import torch
    wl = synth.make_workload("c2", n_bodies=5, n_divides=4, seed=21, rot_deg=rot_deg, trans_m=trans_m)
    pin_c = torch.from_numpy(wl.color_frames).pin_memory()
    pin_d = torch.from_numpy(wl.depth_frames.view(np.uint8).reshape(wl.n_bodies, wl.depth_frames.shape[1], -1)).pin_memory()
    results = []
    for pinned in (False, True):
        ctx = capi.context_from_workload(wl, upload_frames=not pinned)
        moved = 0
        out = []
        for it in range(2):
            if pinned:  # a new frame pair per iteration (same pixels), handed over from pinned memory
                ctx.upload_batch_ptr(True, 0, wl.n_bodies, pin_c.data_ptr(), pin_c.stride(0), pin_c.stride(1))
                ctx.upload_batch_ptr(False, 0, wl.n_bodies, pin_d.data_ptr(), pin_d.stride(0), pin_d.stride(1))
            if it == 0:
                ctx.start_modalities(0)
            ctx.tracking_step(it, wl.n_corr_iterations, wl.n_update_iterations)
            ctx.calculate_results(it)
            out.append(ctx.get_poses())
            if pinned:
                moved = ctx.last_ingest_bytes()
        if pinned:
            full = wl.color_frames[0].nbytes + wl.depth_frames[0].nbytes
            assert 0 < moved < 0.5 * full * wl.n_bodies, (moved, full * wl.n_bodies)
        hf, hb = ctx.get_histograms(0, wl.region.n_histogram_bins)
        results.append((np.stack(out), hf, hb))
        ctx.close()
    assert np.array_equal(results[0][0].view(np.uint32), results[1][0].view(np.uint32))
    assert np.array_equal(results[0][1], results[1][1]) and np.array_equal(results[0][2], results[1][2])


@pytest.mark.parametrize("mode", ["reference_length", "max_length", "depth_scaling"])
def test_adaptive_coverage_and_depth_scaling_bit_exact(capi, oracle, synth, mode):
    """use_adaptive_coverage (region_modality.cpp:415-425, depth_modality.cpp:280-287: the number of lines / points
    follows the contour length / surface area of the closest view, against a reference value or the model's maximum)
    and use_depth_scaling (depth_modality.cpp:829-830: search radius proportional to the point's depth): per-line /
    per-point state bit-exact against the oracle, and fewer items than n_max are processed."""
    import copy
    wl = copy.deepcopy(synth.make_workload("c2", n_bodies=4, n_divides=3, seed=12))
    if mode == "reference_length":
        wl.region.use_adaptive_coverage = True
        wl.region.reference_contour_length = float(np.median(wl.region_model.view_scalars)) * 1.3
        wl.depth.use_adaptive_coverage = True
        wl.depth.reference_surface_area = float(np.median(wl.depth_model.view_scalars)) * 1.3
    elif mode == "max_length":
        wl.region.use_adaptive_coverage = True
        wl.depth.use_adaptive_coverage = True
    else:
        wl.depth.use_depth_scaling = True
        wl.depth.considered_distances = (0.08, 0.04, 0.02)
    ctx, orc = _setup(capi, oracle, wl, oracle.ROTATION_LINEAR, oracle.EXP_RODRIGUES)
    orc.start_modalities(0)
    ctx.start_modalities(0)
    fewer = 0
    for corr in (0, 1, 3):
        ctx.set_poses(orc.get_poses())
        ctx.region_correspondences(0, corr)
        ctx.depth_correspondences(0, corr)
        for b in range(wl.n_bodies):
            n, _ = orc.region_correspondences(b, 0, corr)
            lines = ctx.get_region_lines(b, wl.lines_per_body)
            assert len(lines) == n
            assert_lines_bit_equal(lines, orc.lines[b][:n])
            m, _ = orc.depth_correspondences(b, 0, corr)
            pts = ctx.get_depth_points(b, wl.points_per_body)
            assert len(pts) == m
            assert_points_bit_equal(pts, orc.points[b][:m])
            fewer += int(n < wl.lines_per_body) + int(m < wl.points_per_body)
        ctx.corr_iteration(0, corr, wl.n_update_iterations)
        orc.tracking_step(0, n_corr=corr + 1, corr_begin=corr)
        dt, dr = pose_error(ctx.get_poses(), orc.get_poses())
        assert dt.max() < TOL_POSE_M and dr.max() < TOL_POSE_RAD, (mode, corr, dt, dr)
    if mode != "depth_scaling":
        assert fewer > 0
    ctx.close()


def test_frame_prefetch_matches_synchronous_ingest(capi, synth):
    """m3tb_prefetch_frames: frames handed over one step ahead and ingested on a side stream into the alternate buffers
    while the previous step is tracking; three consecutive frames (different images per step) give bit-identical
    poses to the synchronous path."""
    import torch
    wls = [synth.make_workload("c2", n_bodies=6, n_divides=3, seed=20 + k) for k in range(3)]
    wl = wls[0]
    hc = [torch.from_numpy(w.color_frames).pin_memory() for w in wls]
    hd = [torch.from_numpy(w.depth_frames.view(np.uint8).reshape(w.n_bodies, w.depth_frames.shape[1], -1)).pin_memory() for w in wls]

    def hand_over(ctx, k, prefetch):
        ctx.upload_batch_ptr(True, 0, wl.n_bodies, hc[k].data_ptr(), hc[k].stride(0), hc[k].stride(1))
        ctx.upload_batch_ptr(False, 0, wl.n_bodies, hd[k].data_ptr(), hd[k].stride(0), hd[k].stride(1))
        if prefetch:
            ctx.prefetch_frames()

    out = []
    for prefetch in (False, True):
        ctx = capi.context_from_workload(wl, upload_frames=False)
        hand_over(ctx, 0, prefetch)
        ctx.start_modalities(0)
        poses = []
        for k in range(3):
            ctx.set_poses(wls[k].start_body2world)
            ctx.tracking_step(k, wl.n_corr_iterations, wl.n_update_iterations)
            ctx.calculate_results(k)
            if k < 2:
                hand_over(ctx, k + 1, prefetch)      # while step k may still be running
            poses.append(ctx.get_poses())
        assert ctx.last_ingest_bytes() > 0
        out.append(np.stack(poses))
        ctx.close()
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))
    # and the frames really differed from step to step
    assert np.abs(out[0][0] - out[0][1]).max() > 1e-3
