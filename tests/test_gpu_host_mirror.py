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
    assert out["launches_fused"] == 3
    assert out["launches_object_wise"] == 2 + 7 * (2 + 2 * 3)
    dt, dr = pose_error(fused, obj)
    assert dt.max() < 1e-5 and dr.max() < 1e-5, (dt, dr)
    moved_t, moved_r = pose_error(fused, start)
    assert moved_t.min() > 5e-4 and moved_r.min() > 5e-3   # the 5 mm / 3 deg perturbation was corrected
