"""Pruned GetClosestView (csrc/m3t_b200_views.cuh): the host restatement of the device search must return exactly
the view of the reference's full scan (region_model.cpp:105-130: start value -1, strict >, first maximum wins) for
every query - including ties between duplicated views, non-unit orientations / queries, queries that hit a view
exactly, arbitrary start views for the lower bound, tiny and ragged models - while looking at a small fraction of the
views. Runs without a GPU (the library's host code only)."""
import numpy as np
import pytest


def _check(capi, ori, queries, rng, max_fraction=None):
    prev = rng.integers(0, len(ori), size=len(queries)).astype(np.int32)
    scan, pruned, ev = capi.debug_closest_view(ori, queries, prev)
    # the scan itself against numpy in float32 (same expression, no fused multiply-add)
    o, q = ori.astype(np.float32), queries.astype(np.float32)
    dots = (q[:, None, 0] * o[None, :, 0] + q[:, None, 1] * o[None, :, 1]) + q[:, None, 2] * o[None, :, 2]
    ref = np.where(dots.max(1) > -1.0, dots.argmax(1), 0)
    assert np.array_equal(scan, ref)
    assert np.array_equal(pruned, scan), np.nonzero(pruned != scan)
    if max_fraction is not None:
        assert ev.mean() <= max_fraction * len(ori), (ev.mean(), len(ori))
    return ev


def test_geodesic_models(capi, synth):
    rng = np.random.default_rng(0)
    for n_div in (2, 3, 4):
        m = synth.generate_region_model(n_div, 4, 0.8, 0)
        ori = m.orientations
        q = rng.normal(size=(3000, 3)).astype(np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        ev = _check(capi, ori, q, rng)
        # tracking regime: the start view is the previous answer's neighbourhood -> a handful of clusters
        scan, _, _ = capi.debug_closest_view(ori, q)
        _, pruned2, ev2 = capi.debug_closest_view(ori, q, scan)
        assert np.array_equal(pruned2, scan)
        assert ev2.mean() < (0.12 if n_div == 4 else 0.6) * len(ori), (n_div, ev2.mean())
        # queries exactly on views (lower bound == maximum), scaled queries
        _check(capi, ori, ori[rng.integers(0, len(ori), 500)], rng)
        _check(capi, ori, (q[:500].T * rng.uniform(0.2, 5.0, 500).astype(np.float32)).T, rng)


def test_ties_duplicates_and_non_unit_views(capi):
    rng = np.random.default_rng(1)
    base = rng.normal(size=(300, 3)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    ori = np.concatenate([base, base[::-1], base[:50] * np.float32(0.9986), np.zeros((3, 3), np.float32)])
    q = np.concatenate([base[rng.integers(0, 300, 400)], rng.normal(size=(800, 3)).astype(np.float32)])
    _check(capi, ori, q, rng)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 65, 162, 1000, 5000])
def test_ragged_sizes(capi, n):
    rng = np.random.default_rng(n)
    ori = rng.normal(size=(n, 3)).astype(np.float32)
    ori /= np.linalg.norm(ori, axis=1, keepdims=True)
    if n > 100:  # clustered, non-uniform directions
        ori[: n // 2] = (ori[: n // 2] * 0.05 + np.array([0.3, -0.8, 0.52], np.float32))
        ori /= np.linalg.norm(ori, axis=1, keepdims=True)
    q = rng.normal(size=(500, 3)).astype(np.float32)
    _check(capi, ori, q, rng)
    _check(capi, ori, -ori[rng.integers(0, n, 50)] * np.float32(3.0), rng)   # maximum may be <= -1: view 0


def test_reference_bin_model_views(capi, pkg):
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    rng = np.random.default_rng(5)
    for f in ("region_model.bin", "depth_model.bin"):
        m = pkg.model_io.read_model(os.path.join(here, "golden", f)).model
        q = rng.normal(size=(2000, 3)).astype(np.float32)
        _check(capi, m.orientations, q, rng)
