"""The oracle's restated third-party pieces (Eigen LDLT / matrix exponential / Affine inverse / rotation()) against
independent double-precision references (numpy / scipy)."""
import numpy as np
import pytest
from scipy.linalg import expm, polar


def _rand_pose(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    p = np.zeros((3, 4), np.float32)
    p[:, :3] = R
    p[:, 3] = rng.normal(size=3) * 0.3
    return p


def test_pose_multiply_inverse_rotation(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(20):
        a, b = _rand_pose(rng), _rand_pose(rng)
        out = np.zeros(12, np.float32)
        L.orc_pose_multiply(oracle.ptr(a.reshape(12)), oracle.ptr(b.reshape(12)), oracle.ptr(out))
        A = np.eye(4); A[:3] = a; B = np.eye(4); B[:3] = b
        assert np.allclose(out.reshape(3, 4), (A @ B)[:3], atol=2e-6)
        inv = np.zeros(12, np.float32)
        L.orc_pose_inverse(oracle.ptr(a.reshape(12)), oracle.ptr(inv))
        assert np.allclose(inv.reshape(3, 4), np.linalg.inv(A)[:3], atol=2e-6)
        # rotation(): polar factor of a slightly non-orthogonal linear block
        noisy = a.copy(); noisy[:, :3] += rng.normal(size=(3, 3)).astype(np.float32) * 1e-4
        r = np.zeros(9, np.float32)
        L.orc_pose_rotation(oracle.ptr(noisy.reshape(12)), oracle.ROTATION_POLAR, oracle.ptr(r))
        u, _ = polar(noisy[:, :3].astype(np.float64))
        assert np.allclose(r.reshape(3, 3), u, atol=1e-6)
        L.orc_pose_rotation(oracle.ptr(noisy.reshape(12)), oracle.ROTATION_LINEAR, oracle.ptr(r))
        assert np.array_equal(r.reshape(3, 3), noisy[:, :3])


@pytest.mark.parametrize("scale", [1e-4, 1e-2, 0.09, 0.3, 1.5, 3.0])
def test_exp_skew_pade_and_rodrigues(oracle, scale):
    """Vector2Skewsymmetric(w).exp(): the Pade restatement (Eigen MatrixFunctions) and the closed form used by the
    CUDA path against scipy.linalg.expm; they agree with each other to ~1e-7."""
    L = oracle.lib()
    rng = np.random.default_rng(2)
    for _ in range(10):
        w = rng.normal(size=3); w = (w / np.linalg.norm(w) * scale).astype(np.float32)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], np.float64)
        ref = expm(K)
        rp = np.zeros(9, np.float32); rr = np.zeros(9, np.float32)
        L.orc_exp_skew(oracle.ptr(w), oracle.EXP_PADE, oracle.ptr(rp))
        L.orc_exp_skew(oracle.ptr(w), oracle.EXP_RODRIGUES, oracle.ptr(rr))
        assert np.allclose(rp.reshape(3, 3), ref, atol=5e-7 * max(1.0, scale))
        assert np.allclose(rr.reshape(3, 3), ref, atol=5e-7 * max(1.0, scale))


@pytest.mark.parametrize("n", [1, 6, 13])
def test_ldlt_solve_spd(oracle, n):
    """Eigen::LDLT<MatrixXf, Lower> restatement on SPD systems shaped like the normal equations (rotation block ~1e2,
    translation block ~1e5 + Tikhonov): residual and agreement with numpy."""
    L = oracle.lib()
    rng = np.random.default_rng(3 + n)
    for _ in range(20):
        J = rng.normal(size=(40, n)) * np.array([1.0] * (n // 2) + [30.0] * (n - n // 2))
        a = (J.T @ J + np.diag([1000.0] * (n // 2) + [30000.0] * (n - n // 2))).astype(np.float32)
        b = (rng.normal(size=n) * 100).astype(np.float32)
        x = np.zeros(n, np.float32)
        assert L.orc_ldlt_solve(n, oracle.ptr(np.tril(a).reshape(-1)), oracle.ptr(b), oracle.ptr(x)) == 1
        ref = np.linalg.solve(a.astype(np.float64), b.astype(np.float64))
        assert np.allclose(x, ref, rtol=2e-4, atol=1e-7)


def test_ldlt_pivot_order_and_semidefinite(oracle):
    """Diagonal pivoting: the largest diagonal entry is eliminated first; a zero matrix returns zeros (Eigen's
    "entire diagonal is zero" early-out), an indefinite KKT-style matrix is still solved (optimizer.cpp:152-163)."""
    L = oracle.lib()
    x = np.zeros(3, np.float32)
    z = np.zeros(9, np.float32); b = np.array([1, 2, 3], np.float32)
    L.orc_ldlt_solve(3, oracle.ptr(z), oracle.ptr(b), oracle.ptr(x))
    assert np.array_equal(x, np.zeros(3, np.float32))
    # saddle point [[A, -Jc^T], [-Jc, 0]] (lower part given), A SPD
    A = np.array([[4.0, 1.0], [1.0, 3.0]]); Jc = np.array([[1.0, 2.0]])
    kkt = np.block([[A, -Jc.T], [-Jc, np.zeros((1, 1))]]).astype(np.float32)
    rhs = np.array([1.0, -2.0, 0.5], np.float32)
    L.orc_ldlt_solve(3, oracle.ptr(np.tril(kkt).reshape(-1)), oracle.ptr(rhs), oracle.ptr(x))
    assert np.allclose(kkt.astype(np.float64) @ x, rhs, atol=1e-5)


def test_optimize_rigid_update_convention(oracle):
    """Link::UpdatePoses: body2world <- body2world * [exp(skew(theta_r)) | theta_t] (translate, then rotate;
    link.cpp:222-238) with theta = (-H + diag(lambda))^-1 g; NaN in theta leaves the pose untouched (optimizer.cpp:165)."""
    L = oracle.lib()
    rng = np.random.default_rng(5)
    pose = _rand_pose(rng)
    J = rng.normal(size=(50, 6)) * np.array([1, 1, 1, 30, 30, 30.0])
    H = (-(J.T @ J)).astype(np.float32)
    g = (rng.normal(size=6) * np.array([5, 5, 5, 300, 300, 300.0])).astype(np.float32)
    p = pose.reshape(12).copy(); theta = np.zeros(6, np.float32)
    assert L.orc_optimize_rigid(oracle.ptr(g), oracle.ptr(H.reshape(36)), 1000.0, 30000.0, oracle.EXP_PADE, oracle.ptr(p),
                                oracle.ptr(theta)) == 1
    a = -H.astype(np.float64) + np.diag([1000.0] * 3 + [30000.0] * 3)
    th = np.linalg.solve(a, g.astype(np.float64))
    assert np.allclose(theta, th, rtol=1e-4, atol=1e-8)
    K = np.array([[0, -th[2], th[1]], [th[2], 0, -th[0]], [-th[1], th[0], 0]])
    P = np.eye(4); P[:3] = pose
    V = np.eye(4); V[:3, :3] = expm(K); V[:3, 3] = th[3:]
    assert np.allclose(p.reshape(3, 4), (P @ V)[:3], atol=2e-6)
    g[2] = np.nan
    p2 = pose.reshape(12).copy()
    assert L.orc_optimize_rigid(oracle.ptr(g), oracle.ptr(H.reshape(36)), 1000.0, 30000.0, oracle.EXP_PADE, oracle.ptr(p2),
                                oracle.ptr(theta)) == 0
    assert np.array_equal(p2, pose.reshape(12))
