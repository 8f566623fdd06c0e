"""Shared helpers of the test-suite (pose errors, comparison of line / point records)."""
import numpy as np


def pose_error(p, q):
    """(translation distance [m], rotation angle [rad]) between [n,3,4] pose arrays."""
    p = np.asarray(p, np.float64).reshape(-1, 3, 4)
    q = np.asarray(q, np.float64).reshape(-1, 3, 4)
    dt = np.linalg.norm(p[:, :, 3] - q[:, :, 3], axis=1)
    R = np.einsum("bij,bkj->bik", p[:, :, :3], q[:, :, :3])
    c = np.clip((np.trace(R, axis1=1, axis2=2) - 1.0) / 2.0, -1.0, 1.0)
    # small-angle safe: use the skew part
    s = 0.5 * np.sqrt((R[:, 2, 1] - R[:, 1, 2]) ** 2 + (R[:, 0, 2] - R[:, 2, 0]) ** 2 + (R[:, 1, 0] - R[:, 0, 1]) ** 2)
    return dt, np.arctan2(s, c)


LINE_FIELDS = ["center_f_body", "center_u", "center_v", "normal_u", "normal_v"]
LINE_FIELDS_VALID = ["delta_r", "normal_component_to_scale", "distribution", "mean", "measured_variance"]


def assert_lines_bit_equal(gpu, ref):
    """Per-line records must agree bit for bit (float payloads compared as uint32)."""
    assert len(gpu) == len(ref)
    assert np.array_equal(gpu["valid"], ref["valid"]), np.nonzero(gpu["valid"] != ref["valid"])
    for f in LINE_FIELDS:
        a, b = np.ascontiguousarray(gpu[f]).view(np.uint32), np.ascontiguousarray(ref[f]).view(np.uint32)
        assert np.array_equal(a, b), f
    v = ref["valid"] != 0
    for f in LINE_FIELDS_VALID:
        a = np.ascontiguousarray(gpu[f][v]).view(np.uint32)
        b = np.ascontiguousarray(ref[f][v]).view(np.uint32)
        assert np.array_equal(a, b), (f, np.abs(gpu[f][v] - ref[f][v]).max())


def assert_points_bit_equal(gpu, ref):
    assert len(gpu) == len(ref)
    assert np.array_equal(gpu["valid"], ref["valid"]), np.nonzero(gpu["valid"] != ref["valid"])
    for f in ("center_f_body", "normal_f_body"):
        assert np.array_equal(np.ascontiguousarray(gpu[f]).view(np.uint32), np.ascontiguousarray(ref[f]).view(np.uint32)), f
    v = ref["valid"] != 0
    f = "correspondence_center_f_camera"
    assert np.array_equal(np.ascontiguousarray(gpu[f][v]).view(np.uint32), np.ascontiguousarray(ref[f][v]).view(np.uint32)), f


def rel_to_max(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
