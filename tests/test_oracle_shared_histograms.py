"""RegionModality::UseSharedColorHistograms on the oracle (region_modality.cpp:168-179, 382-386, 575-582;
tracker.cpp:435-443, 507-515): the members of a shared ColorHistograms object add their line pixels to ONE pair of count
arrays, which is initialised / updated once. Closed forms: (i) two identical bodies sharing an object get exactly the
histograms one of them gets alone (counts double, and lr / (2 s) = (lr / s) / 2 exactly); (ii) for different bodies the
shared histogram is the count-weighted mean of the private ones."""
import numpy as np
import pytest


def _workload(synth, n, seed, duplicate=False):
    wl = synth.make_workload("c2", n_bodies=n, n_lines=200, n_points=0, n_divides=2, seed=seed)
    if duplicate:  # body 1 := body 0 (pose, frame)
        wl.start_body2world[1] = wl.start_body2world[0]
        wl.gt_body2world[1] = wl.gt_body2world[0]
        wl.color_frames[1] = wl.color_frames[0]
    return wl


def test_identical_members_equal_the_private_result(oracle, synth):
    private = oracle.OracleTracker(_workload(synth, 2, 3, duplicate=True))
    wl = _workload(synth, 2, 3, duplicate=True)
    wl.histogram_owner = np.array([0, 0], np.int32)
    shared = oracle.OracleTracker(wl)
    for t in (private, shared):
        t.start_modalities(0)
    assert np.array_equal(private.hist_f[0].view(np.uint32), private.hist_f[1].view(np.uint32))
    for b in (0, 1):
        assert np.array_equal(shared.hist_f[b].view(np.uint32), private.hist_f[0].view(np.uint32))
        assert np.array_equal(shared.hist_b[b].view(np.uint32), private.hist_b[0].view(np.uint32))
    for t in (private, shared):
        t.tracking_step(0)
        t.calculate_results(0)
    assert np.array_equal(shared.get_poses().view(np.uint32), private.get_poses().view(np.uint32))
    for b in (0, 1):
        assert np.array_equal(shared.hist_f[b].view(np.uint32), private.hist_f[0].view(np.uint32))
        assert np.array_equal(shared.hist_b[b].view(np.uint32), private.hist_b[0].view(np.uint32))


def test_shared_histogram_is_the_count_weighted_mean(oracle, synth):
    n = 3
    private = oracle.OracleTracker(_workload(synth, n, 5))
    wl = _workload(synth, n, 5)
    wl.histogram_owner = np.array([0, 0, -1], np.int32)   # bodies 0 and 1 share, body 2 keeps its own
    shared = oracle.OracleTracker(wl)
    private.start_modalities(0)
    shared.start_modalities(0)
    # InitializeHistograms: h = counts / sum(counts); private h_b = c_b / s_b, so c_b is known up to s_b. The number of
    # pixels a body adds: every valid line adds at most 18 + 18 pixels; recover s_b from the smallest non-zero entry (= 1 / s_b).
    def counts(h):
        s = np.round(1.0 / h[h > 0].min())
        return np.round(h.astype(np.float64) * s), s
    for hs, hp in ((shared.hist_f, private.hist_f), (shared.hist_b, private.hist_b)):
        c0, s0 = counts(hp[0])
        c1, s1 = counts(hp[1])
        expect = (c0 + c1) / (s0 + s1)
        assert np.abs(hs[0] - expect).max() < 1e-7
        assert np.array_equal(hs[1].view(np.uint32), hs[0].view(np.uint32))          # members see the same object
        assert np.array_equal(hs[2].view(np.uint32), hp[2].view(np.uint32))          # the private body is untouched
    assert np.abs(shared.hist_f[0] - private.hist_f[0]).max() > 1e-5                 # and sharing does change something
