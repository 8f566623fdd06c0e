"""Finite-difference pins of the oracle's gradients / Hessians: the analytic g, H of
RegionModality::CalculateGradientAndHessian (global mode) and DepthModality::CalculateGradientAndHessian are compared
with numerical derivatives of the energies they are derived from, evaluated independently in float64 numpy from
the per-line / per-point records only (sign conventions, weights, the [rot | trans] body-frame parametrisation)."""
import numpy as np
import pytest
from scipy.linalg import expm


@pytest.fixture(scope="module")
def rig(synth, oracle):
    wl = synth.make_workload("c2", n_bodies=1, n_divides=2, seed=4)
    orc = oracle.OracleTracker(wl, rotation_mode=oracle.ROTATION_LINEAR, exp_mode=oracle.EXP_RODRIGUES)
    orc.start_modalities(0)
    return wl, orc


def _variation(theta):
    K = np.array([[0, -theta[2], theta[1]], [theta[2], 0, -theta[0]], [-theta[1], theta[0], 0]])
    V = np.eye(4)
    V[:3, :3] = expm(K)
    V[:3, 3] = theta[3:]
    return V


def _h(p):
    P = np.eye(4)
    P[:3] = np.asarray(p, np.float64).reshape(3, 4)
    return P


def _central(f, n=6, eps=1e-5):
    out = []
    for k in range(n):
        d = np.zeros(n); d[k] = eps
        out.append((f(d) - f(-d)) / (2 * eps))
    return np.array(out)


@pytest.mark.parametrize("corr", [0, 2, 3])
def test_region_gradient_hessian_match_finite_differences(rig, oracle, corr):
    wl, orc = rig
    orc.set_poses(wl.start_body2world)
    n, _ = orc.region_correspondences(0, 0, corr)
    lines = orc.lines[0][:n]
    lines = lines[lines["valid"] != 0]
    assert len(lines) > 30
    g, H = orc.region_gradient_hessian(0, corr, 0)  # opt_iteration 0 < n_global_iterations: global mode
    ci = wl.color_intrinsics
    b2c0 = _h(wl.color_world2camera) @ _h(wl.start_body2world[0])
    sd = wl.region.standard_deviations[min(corr, len(wl.region.standard_deviations) - 1)]
    mev = max(1.0 / (2.0 * np.arctanh(2 * wl.region.function_amplitude) ** 2), wl.region.function_slope)
    Xb = np.concatenate([lines["center_f_body"].astype(np.float64), np.ones((len(lines), 1))], 1)
    ncts = lines["normal_component_to_scale"].astype(np.float64)
    w = mev / (ncts ** 2 * sd ** 2)
    var = lines["measured_variance"].astype(np.float64)
    mean = lines["mean"].astype(np.float64)

    def delta(theta):
        Xc = (b2c0 @ _variation(theta) @ Xb.T).T
        u = Xc[:, 0] * ci.fu / Xc[:, 2] + ci.ppu
        v = Xc[:, 1] * ci.fv / Xc[:, 2] + ci.ppv
        return (lines["normal_u"] * (u - lines["center_u"]) + lines["normal_v"] * (v - lines["center_v"]) -
                lines["delta_r"]) * ncts

    energy = lambda th: float(np.sum(-w * (mean - delta(th)) ** 2 / (2 * var)))
    g_fd = _central(energy)
    J = _central(delta).T                       # [n_lines, 6]
    H_gn = -(J * (w / var)[:, None]).T @ J
    assert np.allclose(g, g_fd, rtol=2e-3, atol=2e-3 * np.abs(g_fd).max())
    assert np.allclose(H, H_gn, rtol=2e-3, atol=2e-3 * np.abs(H_gn).max())
    assert np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H.astype(np.float64)) <= 1e-3)  # negative semi-definite


def test_depth_gradient_hessian_match_finite_differences(rig, oracle):
    wl, orc = rig
    orc.set_poses(wl.start_body2world)
    n, _ = orc.depth_correspondences(0, 0, 0)
    pts = orc.points[0][:n]
    pts = pts[pts["valid"] != 0]
    assert len(pts) > 30
    g, H = orc.depth_gradient_hessian(0, 0)
    b2c0 = _h(wl.depth_world2camera) @ _h(wl.start_body2world[0])
    sd = wl.depth.standard_deviations[0]
    Yc = np.concatenate([pts["correspondence_center_f_camera"].astype(np.float64), np.ones((len(pts), 1))], 1)
    Xb = pts["center_f_body"].astype(np.float64)
    nb = pts["normal_f_body"].astype(np.float64)
    w = 1.0 / (sd * Yc[:, 2])

    def eps(theta):
        Yb = (np.linalg.inv(b2c0 @ _variation(theta)) @ Yc.T).T[:, :3]
        return np.sum(nb * (Xb - Yb), axis=1)

    energy = lambda th: float(-0.5 * np.sum((w * eps(th)) ** 2))
    g_fd = _central(energy)
    J = _central(eps).T
    H_gn = -(J * (w ** 2)[:, None]).T @ J
    assert np.allclose(g, g_fd, rtol=2e-3, atol=2e-3 * np.abs(g_fd).max())
    assert np.allclose(H, H_gn, rtol=2e-3, atol=2e-3 * np.abs(H_gn).max())


def test_local_mode_uses_distribution_log_ratio(rig, oracle):
    """opt_iteration >= n_global_iterations: d loglik / d delta = (ln dist[i_u] - ln dist[i_u - 1]) * learning_rate / var
    with i_u = int(delta + 6.5); lines with i_u <= 0 or >= 12 drop out of both g and H (region_modality.cpp:519-530)."""
    wl, orc = rig
    orc.set_poses(wl.start_body2world)
    n, _ = orc.region_correspondences(0, 0, 1)
    g_glob, H_glob = orc.region_gradient_hessian(0, 1, 0)
    g_loc, H_loc = orc.region_gradient_hessian(0, 1, 1)
    # same lines, same Jacobians: H of the local pass can only lose (negative semi-definite) contributions
    d = (H_loc - H_glob).astype(np.float64)
    assert np.all(np.linalg.eigvalsh(d) >= -1e-2 * np.abs(H_glob).max())
    assert not np.allclose(g_loc, g_glob)


def test_tracking_converges_towards_ground_truth(synth, oracle):
    """End-to-end sanity of the restated loop (Tracker::ExecuteTrackingStep): with region + depth a 5 mm / 3 deg
    perturbation is reduced by > 3x in translation and > 2x in rotation within two frames."""
    from helpers import pose_error
    wl = synth.make_workload("c2", n_bodies=4, n_divides=3, seed=9)
    orc = oracle.OracleTracker(wl)
    orc.start_modalities(0)
    dt0, dr0 = pose_error(wl.start_body2world, wl.gt_body2world)
    for it in range(2):
        orc.tracking_step(it)
        orc.calculate_results(it)
    dt, dr = pose_error(orc.get_poses(), wl.gt_body2world)
    assert np.median(dt) < np.median(dt0) / 3 and np.median(dr) < np.median(dr0) / 2, (dt, dr)
