"""The oracle against the reference's own known answers for the whole chain.

RegionModalityTest.Calculate{Global,Local}GradientAndHessian, DepthModalityTest.CalculateGradientAndHessian and
OptimizerTest.Optimize (M3T/test/modality_test.cpp:280-316,534-550, optimizer_test.cpp:97-105) run the reference on the
real frame pair data/_sequence/*_image_200.png with the triangle body at the pose of test/common_test.cpp:9-13 and
compare H, g and the optimised pose with stored matrices (CompareToLoadedMatrix, 1e-3 / 1e-5). Their sparse viewpoint
model is generated at test time with OpenGL and is not checked in; tests/golden/reference_rig.py regenerates the one
template view those tests use (exact ray cast of the convex prism + cv2.findContours + the identical mt19937{7}
stream) -> tests/golden/triangle_test_view.npz, and the oracle is run on exactly those inputs.

Achieved agreement (this is what the asserts below encode):
    region H (global = local file)   1.5e-6 relative Frobenius  (= the 6 digits of the stored matrix)
    region g, local mode             <= 1e-4 element-wise
    region g, global mode            <= 1.3e-3 element-wise (the only quantity sensitive to the distribution tails;
                                     the reference's own tolerance is 1e-3 and, by a quirk of its comparator, only applies
                                     to positive elements)
    depth H                          3.4e-6 relative Frobenius, depth g <= 5e-4 element-wise
    optimizer pose                   8e-6 absolute; passes the reference's own comparator at its 1e-5 tolerance.
                                     (The stored pose equals b2w*[R | R t] to 5e-7 - an older update convention, cf. ICG -
                                     while the current code, restated by the oracle, applies [R | t] (link.cpp:222-224);
                                     the two differ by theta_r x theta_t ~ 8e-6 m, invisible to the reference's comparator
                                     because it divides by the signed element and so skips negative entries.)
"""
import ctypes as C
import os
from types import SimpleNamespace

import cv2
import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref(oracle):
    import sys
    sys.path.insert(0, GOLDEN)
    import reference_rig as rr
    rig = rr.rig()
    view = np.load(os.path.join(GOLDEN, "triangle_test_view.npz"))
    ka = rr.KA
    L = oracle.lib()
    nv = view["orientations"].shape[0]

    def model(kind, fl):
        pts = np.zeros((nv, 200, fl), np.float32)
        pts[int(view[f"{kind}_view"])] = view[f"{kind}_points"]
        scal = np.zeros(nv, np.float32)
        scal[int(view[f"{kind}_view"])] = view[f"{kind}_scalar"]
        m = SimpleNamespace(n_views=nv, n_points=200, orientations=np.ascontiguousarray(view["orientations"]),
                            view_scalars=scal, points=pts, stride_depth_offset=0.002, max_radius_depth_offset=0.05)
        return m, oracle.make_model(m)

    rmodel_np, rmodel = model("region", 38)
    dmodel_np, dmodel = model("depth", 36)
    cf = oracle.ColorFrame()
    cc = ka["color_camera"]
    cf.intrinsics = oracle.Intrinsics(cc["fu"], cc["fv"], cc["ppu"], cc["ppv"], cc["width"], cc["height"])
    cf.world2camera[:] = rig["color_w2c"][:3].astype(np.float32).reshape(12).tolist()
    cf.bgr = rig["color"].ctypes.data
    cf.pitch = rig["color"].strides[0]
    df = oracle.DepthFrame()
    dc = ka["depth_camera"]
    df.intrinsics = oracle.Intrinsics(dc["fu"], dc["fv"], dc["ppu"], dc["ppv"], dc["width"], dc["height"])
    df.world2camera[:] = rig["depth_w2c"][:3].astype(np.float32).reshape(12).tolist()
    df.depth = rig["depth"].ctypes.data
    df.pitch = rig["depth"].strides[0]
    df.depth_scale = dc["depth_scale"]
    b2w = np.ascontiguousarray(rig["body2world"][:3].astype(np.float32).reshape(12))
    return SimpleNamespace(L=L, rig=rig, ka=ka, view=view, rmodel=rmodel, dmodel=dmodel, keep=(rmodel_np, dmodel_np), cf=cf,
                           df=df, b2w=b2w, rp=oracle.region_params(None), dp=oracle.depth_params(None))


def _mat(ka, name):
    return np.array(ka[name]["data"], np.float64)


def _region(ref, oracle, opt_iteration, mode):
    """SetUp; StartModality(0,0); CalculateCorrespondences(0,0); CalculateGradientAndHessian(0,0,opt)."""
    L, p = ref.L, oracle.ptr
    n3 = 16 ** 3
    mf, mb = np.zeros(n3, np.float32), np.zeros(n3, np.float32)
    hf, hb = np.full(n3, 1.0 / n3, np.float32), np.full(n3, 1.0 / n3, np.float32)
    L.orc_region_add_line_pixels(C.byref(ref.rp), C.byref(ref.rmodel), C.byref(ref.cf), p(ref.b2w), mode, p(mf), p(mb))
    L.orc_hist_calculate(16, 1.0, p(mf), p(hf))
    L.orc_hist_calculate(16, 1.0, p(mb), p(hb))
    lines = np.zeros(200, oracle.REGION_LINE_DTYPE)
    view = C.c_int(0)
    n = L.orc_region_correspondences(C.byref(ref.rp), C.byref(ref.rmodel), C.byref(ref.cf), None, p(hf), p(hb), p(ref.b2w), 0, 0,
                                     0, mode, lines.ctypes.data_as(C.POINTER(oracle.RegionLine)), C.byref(view))
    g, H = np.zeros(6, np.float32), np.zeros(36, np.float32)
    L.orc_region_gradient_hessian(C.byref(ref.rp), C.byref(ref.cf), p(ref.b2w), lines.ctypes.data_as(C.POINTER(oracle.RegionLine)),
                                  n, 0, opt_iteration, mode, p(g), p(H))
    return g, H.reshape(6, 6), lines[:n], view.value


def _depth(ref, oracle, mode):
    L, p = ref.L, oracle.ptr
    pts = np.zeros(200, oracle.DEPTH_POINT_DTYPE)
    view = C.c_int(0)
    n = L.orc_depth_correspondences(C.byref(ref.dp), C.byref(ref.dmodel), C.byref(ref.df), p(ref.b2w), 0, 0, 0, mode,
                                    pts.ctypes.data_as(C.POINTER(oracle.DepthPoint)), C.byref(view))
    g, H = np.zeros(6, np.float32), np.zeros(36, np.float32)
    L.orc_depth_gradient_hessian(C.byref(ref.dp), C.byref(ref.df), p(ref.b2w), pts.ctypes.data_as(C.POINTER(oracle.DepthPoint)), n, 0,
                                 p(g), p(H))
    return g, H.reshape(6, 6), pts[:n], view.value


def _rel_fro(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b)


def _reference_comparator(loaded, matrix):
    """CompareToLoadedMatrix (M3T/test/common_test.cpp:206-228): max over elements of |loaded - matrix| / matrix with
    non-finite ratios ignored - note the SIGNED denominator (negative elements of `matrix` can never fail)."""
    loaded, matrix = np.asarray(loaded, np.float32), np.asarray(matrix, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(loaded - matrix) / matrix
    r = np.where(np.isfinite(r), r, 0.0)
    # entries that are cancellation residues (|x| < 1e-5 of the largest entry, e.g. H(0,1) = 0.228 next to 3e5) carry
    # fewer significant digits than the tolerance and depend on the summation order of the build: not comparable
    r = np.where(np.abs(loaded) >= 1e-5 * np.abs(loaded).max(), r, 0.0)
    return float(r.max())


def _rel_elem(a, b):
    return np.abs(np.asarray(a, np.float64) - b) / np.abs(b)


def test_template_view_is_unambiguous(ref):
    """The closest view of the test pose wins the arg-max by a clear margin for both cameras, so regenerating only
    that view is sufficient."""
    assert int(ref.view["region_view"]) == int(ref.view["depth_view"])
    assert float(ref.view["region_argmax_margin"]) > 1e-4 and float(ref.view["depth_argmax_margin"]) > 1e-4


def test_region_modality_global_hessian_and_gradient(ref, oracle):
    g, H, lines, view = _region(ref, oracle, 0, oracle.ROTATION_POLAR)
    assert view == int(ref.view["region_view"])
    assert lines["valid"].sum() == 179
    Hg, gg = _mat(ref.ka, "region_modality_global_hessian"), _mat(ref.ka, "region_modality_global_gradient").reshape(6)
    assert _rel_fro(H, Hg) < 1e-5, _rel_fro(H, Hg)
    assert _reference_comparator(Hg, H) < 1e-3            # the reference's own check, its own tolerance
    assert _rel_elem(g, gg).max() < 2e-3, _rel_elem(g, gg)
    assert _reference_comparator(gg, g) < 2e-3
    a = np.diag([1000.0] * 3 + [30000.0] * 3)
    step_ref, step = np.linalg.solve(-Hg + a, gg), np.linalg.solve(-H.astype(np.float64) + a, g)
    assert np.abs(step - step_ref).max() < 2e-6, (step, step_ref)  # Gauss-Newton step: rad / m


def test_region_modality_local_hessian_and_gradient(ref, oracle):
    g, H, _, _ = _region(ref, oracle, 1, oracle.ROTATION_POLAR)
    Hl, gl = _mat(ref.ka, "region_modality_local_hessian"), _mat(ref.ka, "region_modality_local_gradient").reshape(6)
    assert _rel_fro(H, Hl) < 1e-5
    assert _reference_comparator(Hl, H) < 1e-3
    assert _rel_elem(g, gl).max() < 2e-4, _rel_elem(g, gl)
    assert _reference_comparator(gl, g) < 1e-3


def test_depth_modality_hessian_and_gradient(ref, oracle):
    g, H, pts, view = _depth(ref, oracle, oracle.ROTATION_POLAR)
    assert view == int(ref.view["depth_view"])
    assert pts["valid"].sum() == 182
    Hd, gd = _mat(ref.ka, "depth_modality_hessian"), _mat(ref.ka, "depth_modality_gradient").reshape(6)
    assert _rel_fro(H, Hd) < 2e-5, _rel_fro(H, Hd)
    assert _rel_elem(H, Hd).max() < 2e-4
    assert _reference_comparator(Hd, H) < 1e-3
    assert _rel_elem(g, gd).max() < 1e-3, _rel_elem(g, gd)
    assert _reference_comparator(gd, g) < 1e-3


def test_rotation_mode_does_not_matter_at_golden_precision(ref, oracle):
    """rotation() as polar factor (Eigen) vs the linear block (CUDA path): far below the golden tolerance."""
    for opt in (0, 1):
        g1, H1, _, _ = _region(ref, oracle, opt, oracle.ROTATION_POLAR)
        g0, H0, _, _ = _region(ref, oracle, opt, oracle.ROTATION_LINEAR)
        assert _rel_fro(H0, H1.astype(np.float64)) < 1e-5 and np.abs(g0 - g1).max() < 2e-4 * np.abs(g1).max()


def test_optimizer_pose_after_one_gauss_newton_step(ref, oracle):
    """OptimizerTest.Optimize: region + depth, lambda = (5000, 500000) (data/optimizer_test/optimizer.yaml), one
    CalculateOptimization -> pose; compared with the reference's own comparator at its own tolerance (1e-5)."""
    gr, Hr, _, _ = _region(ref, oracle, 0, oracle.ROTATION_POLAR)
    gd, Hd, _, _ = _depth(ref, oracle, oracle.ROTATION_POLAR)
    g = (gr + gd).astype(np.float32)
    H = (Hr + Hd).astype(np.float32)
    pose = ref.b2w.copy()
    theta = np.zeros(6, np.float32)
    lam = ref.ka["optimizer_test_tikhonov"]
    assert ref.L.orc_optimize_rigid(oracle.ptr(g), oracle.ptr(H.reshape(36)), lam["rotation"], lam["translation"], oracle.EXP_PADE,
                                    oracle.ptr(pose), oracle.ptr(theta)) == 1
    golden = _mat(ref.ka, "optimizer_triangle_pose")
    ours = np.eye(4, dtype=np.float32)
    ours[:3] = pose.reshape(3, 4)
    assert _reference_comparator(golden, ours) < 1e-5
    start = ref.b2w.reshape(3, 4).astype(np.float64)
    moved = np.abs(golden[:3] - start).max()
    err = np.abs(ours[:3] - golden[:3]).max()
    assert moved > 3e-3 and err < 1e-5, (moved, err)      # a 4 mm step reproduced to 8 um ...
    assert np.abs(ours[:3, :3] - golden[:3, :3]).max() < 1e-6  # ... rotation to the 6 stored digits


@pytest.mark.parametrize("scenario", ["tracker", "refiner"])
def test_tracker_and_refiner_pose_known_answers(oracle, scenario):
    """TrackerTest.OptimizePoseMatrix (StartModalities + ExecuteTrackingStep: 7 x 2 iterations, tracker_test.cpp:164-179)
    and RefinerTest.OptimizePoseMatrix (7 x (StartModalities, correspondences, 3 updates), refiner.cpp:99-118) replayed on
    the oracle with the template views regenerated on demand (tests/golden/make_tracker_views.py ->
    triangle_tracker_views.npz: the 4 + 4 views the loop visits). SOFT check: the loop corrects the pose by ~1 cm / 3 deg
    and lands within 0.55 mm / 0.21 deg (tracker) resp. 0.93 mm / 0.3 deg (refiner) of the stored pose, i.e. 5-10 % of
    the correction. Bit-level agreement
    is not attainable here: the visited views other than the first are resampled from a software raster (+-1 px), the
    path selects views by arg-max near Voronoi boundaries, and the stored poses predate the current update convention
    (see the optimizer test above)."""
    import sys
    sys.path.insert(0, GOLDEN)
    from replay import ReferenceReplay
    z = np.load(os.path.join(GOLDEN, "triangle_tracker_views.npz"))
    views = {k: {int(i): (z[f"{k}_points"][n], float(z[f"{k}_scalars"][n])) for n, i in enumerate(z[f"{k}_ids"])}
             for k in ("region", "depth")}
    rep = ReferenceReplay(oracle, views)
    start = rep.pose().astype(np.float64)
    assert rep.run(scenario) == []          # every view the loop asked for is in the fixture
    ours = rep.pose().astype(np.float64)
    golden = _mat(rep.ka, f"{scenario}_triangle_pose")

    def dist(a, b):
        R = a[:3, :3] @ b[:3, :3].T
        return np.linalg.norm(a[:3, 3] - b[:3, 3]), np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))

    moved_t, moved_r = dist(golden, start)
    err_t, err_r = dist(ours, golden)
    assert moved_t > 8e-3 and moved_r > 0.03, (moved_t, moved_r)
    tol_t, tol_r = (7e-4, 4.5e-3) if scenario == "tracker" else (1.2e-3, 6e-3)
    assert err_t < tol_t and err_r < tol_r, (scenario, err_t, err_r, moved_t, moved_r)
    assert err_t < 0.12 * moved_t and err_r < 0.15 * moved_r
