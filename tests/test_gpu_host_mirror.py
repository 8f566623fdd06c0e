"""The C++ mirror of the reference interface (m3t_b200::Body / Camera / Model / Modality / Link / Optimizer / Tracker):
an application written like an M3T application drives the CUDA path through it."""
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import pose_error

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_tracker_fused_and_object_wise_agree(pkg):
    exe = pkg._build.build_host_example()
    r = subprocess.run([exe, "4", "200", "200", "3", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().split("\n")[-1])
    assert out["refused_without_setup"] is True          # "Set up tracker ... first"
    fused = np.array(out["fused"], np.float32).reshape(-1, 3, 4)
    obj = np.array(out["object_wise"], np.float32).reshape(-1, 3, 4)
    start = np.array(out["start"], np.float32).reshape(-1, 3, 4)
    # Tracker::ExecuteTrackingStep = 1 fused launch (+ StartModalities + CalculateResults histogram launches);
    # the object-wise path issues one batched launch per phase, not one per object
    assert out["launches_fused"] in (3, 4)   # + k_bin (bin-index images of the freshly copied frames) on the k_track2 path
    assert out["launches_object_wise"] in (2 + 7 * (2 + 2 * 3), 3 + 7 * (2 + 2 * 3))
    dt, dr = pose_error(fused, obj)
    assert dt.max() < 1e-5 and dr.max() < 1e-5, (dt, dr)
    moved_t, moved_r = pose_error(fused, start)
    assert moved_t.min() > 5e-4 and moved_r.min() > 5e-3   # the 5 mm / 3 deg perturbation was corrected


def test_cpp_tracker_kinematic_chains(pkg):
    """The same application with m3t::Link trees: 2 chains x 4 links (root 6 DoF + revolute-x children), one Optimizer
    per chain; start poses of the children from Optimizer::CalculateConsistentPoses; fused Tracker path (k_track +
    k_structure per update) vs the object-wise Modality / Optimizer fan-out (m3tb_calculate_optimization)."""
    exe = pkg._build.build_host_example()
    r = subprocess.run([exe, "8", "200", "200", "3", "2", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().split("\n")[-1])
    assert out["links_per_structure"] == 4 and out["joints_ok"] is True
    fused = np.array(out["fused"], np.float32).reshape(-1, 3, 4)
    obj = np.array(out["object_wise"], np.float32).reshape(-1, 3, 4)
    start = np.array(out["start"], np.float32).reshape(-1, 3, 4)
    dt, dr = pose_error(fused, obj)
    # the two paths sum region + depth in a different order (per thread vs per modality): rounding-level agreement
    # amplified by the chain iteration, not bit equality
    assert np.median(dt) < 2e-5 and np.median(dr) < 2e-4, (dt, dr)
    assert dt.max() < 1e-3 and dr.max() < 1e-2, (dt, dr)
    moved_t, moved_r = pose_error(fused, start)
    assert np.median(moved_t) > 5e-4 and np.median(moved_r) > 5e-3
    # the chain stayed a chain: consecutive links are Tx(0.01) apart
    for c in range(2):
        for j in range(1, 4):
            a, b = fused[4 * c + j - 1], fused[4 * c + j]
            rel_t = a[:, :3].T @ (b[:, 3] - a[:, 3])
            assert np.allclose(rel_t, (0.01, 0, 0), atol=2e-5), rel_t
