"""The oracle against the reference's own known answers / fixtures that are reproducible without OpenGL
(SURVEY §4, §8c): exact ColorHistograms values, the checked-in .bin sparse-viewpoint models, parameter defaults."""
import ctypes as C
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_color_histograms_known_answers(oracle):
    """ColorHistogramsTest.TestHistogramCalculation (M3T/test/color_histograms_test.cpp:71-104): n_bins 32,
    learning rates 0.5, colours (122,154,63) and (78,64,187); every expected value is exact."""
    L = oracle.lib()
    nb = 32
    n3 = nb ** 3
    c1 = (C.c_uint8 * 3)(122, 154, 63)
    c2 = (C.c_uint8 * 3)(78, 64, 187)
    hf = np.full(n3, 1.0 / n3, np.float32)   # SetUpHistograms: uniform
    hb = np.full(n3, 1.0 / n3, np.float32)
    mf = np.zeros(n3, np.float32)
    mb = np.zeros(n3, np.float32)
    p = oracle.ptr

    def get(c):
        a, b = C.c_float(0), C.c_float(0)
        L.orc_hist_get(nb, p(hf), p(hb), c, C.byref(a), C.byref(b))
        return a.value, b.value

    L.orc_hist_add(nb, p(mf), c1)
    L.orc_hist_add(nb, p(mb), c2)
    L.orc_hist_clear(nb, p(mf), p(mb))           # ClearMemory
    L.orc_hist_calculate(nb, 0.5, p(mf), p(hf))  # UpdateHistograms with empty memory: unchanged
    L.orc_hist_calculate(nb, 0.5, p(mb), p(hb))
    assert get(c1) == (np.float32(1.0) / np.float32(n3),) * 2
    L.orc_hist_add(nb, p(mf), c1); L.orc_hist_add(nb, p(mf), c2); L.orc_hist_add(nb, p(mb), c2)
    L.orc_hist_calculate(nb, 1.0, p(mf), p(hf))  # InitializeHistograms
    L.orc_hist_calculate(nb, 1.0, p(mb), p(hb))
    L.orc_hist_clear(nb, p(mf), p(mb))
    assert get(c1) == (0.5, 0.0)
    assert get(c2) == (0.5, 1.0)
    L.orc_hist_add(nb, p(mf), c2); L.orc_hist_add(nb, p(mb), c1)
    L.orc_hist_calculate(nb, 0.5, p(mf), p(hf))  # UpdateHistograms
    L.orc_hist_calculate(nb, 0.5, p(mb), p(hb))
    assert get(c1) == (0.25, 0.5)
    assert get(c2) == (0.75, 0.5)


def test_reference_bin_models_parse_and_round_trip(pkg, tmp_path):
    """The reference's checked-in models (data/model_test/{region,depth}_model.bin): sizes, header values, payload
    sanity (unit normals, view orientation = unit vectors) and a byte-exact write-back."""
    io = pkg.model_io
    for name, kind, fl in (("region_model.bin", "region", 38), ("depth_model.bin", "depth", 36)):
        path = os.path.join(GOLDEN, name)
        mf = io.read_model(path)
        assert mf.kind == kind and mf.version == (10 if kind == "region" else 9)
        assert mf.n_divides == 2 and mf.model.n_views == 162 and mf.n_points == 10
        assert abs(mf.sphere_radius - 0.4) < 1e-6 and mf.image_size == 500  # data/model_test/region_model.yaml
        assert mf.body.geometry_path.endswith(b"schauma.obj")
        m = mf.model
        assert m.points.shape == (162, 10, fl)
        assert np.allclose(np.linalg.norm(m.orientations, axis=1), 1.0, atol=1e-5)
        nrm = np.linalg.norm(m.points[:, :, 3:6], axis=2)
        assert np.all(nrm > 0.97) and np.all(nrm < 1.03)          # depth normals are 8-bit decoded, not renormalised
        assert np.all(np.linalg.norm(m.points[:, :, :3], axis=2) < 0.2)  # body-frame metres
        if kind == "region":
            assert np.all(m.points[:, :, 6] >= 0) and np.all(m.points[:, :, 7] > 1e30)  # background_distance = FLT_MAX
            assert np.all(m.view_scalars > 0)
        out = tmp_path / name
        io.write_model(out, mf)
        assert open(out, "rb").read() == open(path, "rb").read()


def test_closest_view_on_reference_model(pkg, oracle):
    """GetClosestView on the reference's own model: feeding a view's orientation back (body2camera whose
    R^T t/|t| equals that orientation) must select that view; zero translation selects view 0."""
    mf = pkg.model_io.read_model(os.path.join(GOLDEN, "region_model.bin"))
    om = oracle.make_model(mf.model)
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for v in rng.choice(mf.model.n_views, 25, replace=False):
        o = mf.model.orientations[v].astype(np.float64)
        # any rotation R, translation t = R o * 0.7
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        pose = np.zeros((3, 4), np.float32)
        pose[:, :3] = R
        pose[:, 3] = R @ o * 0.7
        for mode in (oracle.ROTATION_LINEAR, oracle.ROTATION_POLAR):
            assert L.orc_closest_view(C.byref(om), oracle.ptr(pose.reshape(12)), mode) == v
    zero = np.zeros(12, np.float32); zero[[0, 5, 10]] = 1
    assert L.orc_closest_view(C.byref(om), oracle.ptr(zero), oracle.ROTATION_POLAR) == 0


def test_defaults_match_reference_headers(oracle, pkg):
    """Parameter defaults = the reference's in-class initialisers (region_modality.h:411-443, depth_modality.h:302-321)."""
    rp, dp = oracle.region_params(None), oracle.depth_params(None)
    assert (rp.n_lines_max, rp.function_length, rp.distribution_length, rp.n_global_iterations) == (200, 8, 12, 1)
    assert list(rp.scales)[:4] == [6, 4, 2, 1] and np.allclose(list(rp.standard_deviations)[:4], [15, 5, 3.5, 1.5])
    assert np.isclose(rp.function_amplitude, 0.43) and np.isclose(rp.function_slope, 0.5) and np.isclose(rp.learning_rate, 1.3)
    assert rp.n_histogram_bins == 16 and np.isclose(rp.learning_rate_f, 0.2) and np.isclose(rp.min_continuous_distance, 3.0)
    assert dp.n_points_max == 200 and np.isclose(dp.stride_length, 0.005)
    assert np.allclose(list(dp.considered_distances)[:3], [0.05, 0.02, 0.01])
    assert np.allclose(list(dp.standard_deviations)[:3], [0.05, 0.03, 0.02])
    ka = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))
    assert ka["region_modality_global_hessian"]["rows"] == 6  # fixtures present (used by test_reference_goldens.py)


def test_function_lookup_and_min_expected_variance(oracle):
    """PrecalculateFunctionLookup / DistributionVariables (region_modality.cpp:910-936): tanh smoothed step,
    min_expected_variance = max(1/(2 atanh(2A)^2), slope): 0.5 for the defaults, 0.607 for the RBOT setting."""
    L = oracle.lib()
    rp = oracle.region_params(None)
    lf = np.zeros(8, np.float32); lb = np.zeros(8, np.float32); mev = C.c_float(0)
    L.orc_function_lookup(C.byref(rp), oracle.ptr(lf), oracle.ptr(lb), C.byref(mev))
    x = np.arange(8) - 3.5
    assert np.allclose(lf, 0.5 - 0.43 * np.tanh(x / 1.0), atol=1e-6) and np.allclose(lf + lb, 1.0)
    assert abs(mev.value - 0.5) < 1e-7
    rp.function_amplitude, rp.function_slope = 0.36, 0.0
    L.orc_function_lookup(C.byref(rp), oracle.ptr(lf), oracle.ptr(lb), C.byref(mev))
    assert np.allclose(lf, [0.86] * 4 + [0.14] * 4, atol=1e-6)
    assert abs(mev.value - 1.0 / (2.0 * np.arctanh(0.72) ** 2)) < 1e-5


def test_oracle_is_frozen(oracle):
    """The oracle checks every GPU parity test; its own outputs on small seeded workloads (rigid, chains with projected /
    constrained / soft joints, measured occlusion) are frozen in tests/golden/oracle_regression.npz
    (tests/golden/make_oracle_regression.py) so that an edit cannot shift them unnoticed."""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_oracle_regression as m
    ref = np.load(os.path.join(GOLDEN, "oracle_regression.npz"))
    for name, wl in m.cases():
        poses = m.run(oracle, wl)
        assert np.allclose(poses, ref[name], atol=2e-6), (name, np.abs(poses - ref[name]).max())
