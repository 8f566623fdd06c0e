"""Kinematic-structure maths of the oracle (link.cpp, constraint.cpp, soft_constraint.cpp,
optimizer.cpp). The reference has no known answers for this part (parity unpinned), so it is checked through closed
forms: scipy for Eigen's AngleAxis conversion, finite differences for every Jacobian, the one-step convergence of
examples/constraint_convergence.cpp, and the equivalence of a one-link structure with the rigid-body optimiser."""
import importlib

import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial.transform import Rotation

synth = importlib.import_module("3dobjecttracking_b200.synth")


def _T(p):
    m = np.eye(4)
    m[:3] = np.asarray(p, np.float64).reshape(3, 4)
    return m


def _p(m):
    return np.asarray(m[:3], np.float32)


def _rand_pose(rng, angle=np.pi, trans=0.3):
    rv = rng.normal(size=3)
    rv *= rng.uniform(0, angle) / np.linalg.norm(rv)
    p = np.zeros((3, 4), np.float32)
    p[:, :3] = Rotation.from_rotvec(rv).as_matrix()
    p[:, 3] = rng.normal(size=3) * trans
    return p


def _variation(theta):
    """[exp(skew(theta_r)) | theta_t] (link.cpp:222-224)"""
    m = np.eye(4)
    m[:3, :3] = Rotation.from_rotvec(theta[:3]).as_matrix()
    m[:3, 3] = theta[3:]
    return m


def test_angle_axis_matches_scipy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for k in range(200):
        ang = [0.0, 1e-7, 1e-4, np.pi - 1e-4, np.pi][k] if k < 5 else rng.uniform(0, np.pi)
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        R = Rotation.from_rotvec(ang * ax).as_matrix().astype(np.float32)
        angle = np.zeros(1, np.float32); axis = np.zeros(3, np.float32)
        L.orc_angle_axis(oracle.ptr(R.reshape(9)), oracle.ptr(angle), oracle.ptr(axis))
        rv_ref = Rotation.from_matrix(R.astype(np.float64)).as_rotvec()
        assert 0.0 <= angle[0] <= np.pi + 1e-6
        if ang > 1e-3 and ang < np.pi - 1e-3:
            assert abs(np.linalg.norm(axis) - 1.0) < 1e-5
            assert np.allclose(angle[0] * axis, rv_ref, atol=5e-6 + 2e-3 * (ang > 3.0))
        else:
            assert abs(angle[0] - np.linalg.norm(rv_ref)) < 2e-3


def test_xcotx(oracle):
    L = oracle.lib()
    assert L.orc_xcotx(0.0) == 1.0
    for x in (1e-3, 0.3, 1.0, 1.5):
        assert abs(L.orc_xcotx(x) - x / np.tan(x)) < 1e-6
    assert abs(L.orc_xcotx(np.float32(1.5707960))) < 1e-6
    # quirk kept from common.h:74: float(pi/2) lies just above pi/2, tanf is negative there and the first guard
    # (tan <= FLT_MIN) answers 1 instead of ~0
    assert L.orc_xcotx(np.float32(np.pi / 2)) == 1.0


def _two_link_structure(rng, free2=(1, 1, 1, 1, 1, 1), directions=(1, 1, 1, 1, 1, 1), soft=False, **soft_kw):
    b12j1, b22j2 = _rand_pose(rng, 1.0), _rand_pose(rng, 1.0)
    links = [synth.LinkSpec(body=-1, parent=-1, body2joint=synth.identity_pose(), joint2parent=synth.identity_pose()),
             synth.LinkSpec(body=-1, parent=0, body2joint=b22j2, joint2parent=_p(np.linalg.inv(_T(b12j1))),
                            free_directions=free2)]
    cons = [synth.ConstraintSpec(link1=0, link2=1, body12joint1=b12j1, body22joint2=b22j2, directions=directions,
                                 soft=soft, **soft_kw)]
    return synth.StructureSpec(links=links, constraints=cons)


def test_constraint_convergence(oracle):
    """examples/constraint_convergence.cpp: a random violation of a fully constrained joint vanishes within a few
    CalculateOptimization calls (Newton on the constraint with the exact log-map Jacobian)."""
    L = oracle.lib()
    rng = np.random.default_rng(5)
    worst = 0.0
    for run in range(30):
        spec = _two_link_structure(rng)
        spec.links[1].joint2parent = _p(_T(spec.links[1].joint2parent) @ _T(_rand_pose(rng, 0.8, 0.2)))
        so = oracle.OracleStructure(spec)
        S = so.as_struct()
        l2w = np.zeros((2, 12), np.float32)
        l2w[0] = _rand_pose(rng).reshape(12)
        L.orc_structure_consistent_poses(S, oracle.EXP_PADE, oracle.ptr(l2w))
        z = np.zeros(2 * 42, np.float32)
        errs = []
        for it in range(6):
            _, j2p = so.joint_poses()
            err = _T(spec.constraints[0].body12joint1) @ _T(j2p[1])
            errs.append((np.linalg.norm(Rotation.from_matrix(err[:3, :3]).as_rotvec()), np.linalg.norm(err[:3, 3])))
            assert L.orc_optimize_structure(S, oracle.ptr(z[:12]), oracle.ptr(z[12:]), oracle.ROTATION_POLAR,
                                            oracle.EXP_PADE, oracle.ptr(l2w), None) == 1
        assert errs[0][0] > 1e-3
        worst = max(worst, errs[-1][0], errs[-1][1])
    assert worst < 5e-6, worst


def test_constraint_jacobian_finite_differences(oracle):
    """d residual / d theta of both links, theta being the right-multiplied body-frame variation."""
    L = oracle.lib()
    rng = np.random.default_rng(7)
    for run in range(10):
        b12j1, b22j2 = _rand_pose(rng, 1.0), _rand_pose(rng, 1.0)
        links = [synth.LinkSpec(body=-1, parent=-1, body2joint=synth.identity_pose(), joint2parent=synth.identity_pose(),
                                free_directions=(0,) * 6),
                 synth.LinkSpec(body=-1, parent=0, body2joint=synth.identity_pose(), joint2parent=_rand_pose(rng)),
                 synth.LinkSpec(body=-1, parent=0, body2joint=synth.identity_pose(), joint2parent=_rand_pose(rng))]
        # joint violation of moderate size so that the log map is well inside (0, pi)
        spec = synth.StructureSpec(links=links, constraints=[synth.ConstraintSpec(
            link1=1, link2=2, body12joint1=b12j1, body22joint2=b22j2, directions=(1, 1, 1, 1, 1, 1))])
        so = oracle.OracleStructure(spec)
        S = so.as_struct()
        l2w = np.zeros((3, 12), np.float32)
        l2w[0] = synth.identity_pose().reshape(12)
        L.orc_structure_consistent_poses(S, oracle.EXP_PADE, oracle.ptr(l2w))
        dof = L.orc_structure_dof(S)
        assert dof == 12
        jac = np.zeros((3, 6, dof), np.float32)
        L.orc_structure_jacobians(S, oracle.ROTATION_POLAR, oracle.ptr(jac))
        assert np.allclose(jac[1][:, :6], np.eye(6), atol=1e-6) and np.allclose(jac[2][:, 6:], np.eye(6), atol=1e-6)
        res = np.zeros(6, np.float32); cj = np.zeros((6, dof), np.float32)
        assert L.orc_constraint_residual_jacobian(S, 0, oracle.ptr(l2w), oracle.ptr(jac), oracle.ROTATION_POLAR,
                                                  oracle.ptr(res), oracle.ptr(cj)) == 6

        def residual(T1, T2):
            j = _T(b12j1) @ np.linalg.inv(T1) @ T2 @ np.linalg.inv(_T(b22j2))
            return np.concatenate([Rotation.from_matrix(j[:3, :3]).as_rotvec(), j[:3, 3]])

        T1, T2 = _T(l2w[1]), _T(l2w[2])
        assert np.allclose(res, residual(T1, T2), atol=2e-5)
        eps = 1e-6
        num = np.zeros((6, 12))
        for k in range(6):
            th = np.zeros(6); th[k] = eps
            num[:, k] = (residual(T1 @ _variation(th), T2) - residual(T1 @ _variation(-th), T2)) / (2 * eps)
            num[:, 6 + k] = (residual(T1, T2 @ _variation(th)) - residual(T1, T2 @ _variation(-th))) / (2 * eps)
        assert np.allclose(cj, num, atol=3e-4), np.abs(cj - num).max()


def test_chain_jacobians_finite_differences(oracle):
    """Link::CalculateJacobian: the body-frame motion of every link caused by a small step in the joint unknowns."""
    L = oracle.lib()
    rng = np.random.default_rng(11)
    links = [synth.LinkSpec(body=-1, parent=-1, body2joint=_rand_pose(rng, 0.5, 0.05), joint2parent=synth.identity_pose()),
             synth.LinkSpec(body=-1, parent=0, body2joint=_rand_pose(rng, 0.5, 0.05), joint2parent=_rand_pose(rng, 1.0, 0.1),
                            free_directions=(1, 0, 0, 0, 0, 0)),
             synth.LinkSpec(body=-1, parent=1, body2joint=_rand_pose(rng, 0.5, 0.05), joint2parent=_rand_pose(rng, 1.0, 0.1),
                            free_directions=(0, 1, 0, 1, 0, 0), fixed_body2joint_pose=False),
             synth.LinkSpec(body=-1, parent=0, body2joint=synth.identity_pose(), joint2parent=_rand_pose(rng, 1.0, 0.1),
                            free_directions=(1, 1, 1, 0, 0, 0))]
    spec = synth.StructureSpec(links=links)

    so = oracle.OracleStructure(spec)
    S = so.as_struct()
    base = np.zeros((4, 12), np.float32)
    base[0] = synth.identity_pose().reshape(12)
    L.orc_structure_consistent_poses(S, oracle.EXP_PADE, oracle.ptr(base))
    dof = L.orc_structure_dof(S)
    assert dof == 6 + 1 + 2 + 3
    jac = np.zeros((4, 6, dof), np.float32)
    L.orc_structure_jacobians(S, oracle.ROTATION_POLAR, oracle.ptr(jac))

    # Link::UpdatePoses restated in numpy (double): the forward kinematics the Jacobians must be the derivative of
    def forward(theta):
        T = [None] * 4
        b2j = [_T(l.body2joint) for l in links]
        j2p = [_T(l.joint2parent) for l in links]
        idx = 0
        for i, l in enumerate(links):
            th = np.zeros(6)
            for d in range(6):
                if l.free_directions[d]:
                    th[d] = theta[idx]; idx += 1
            var = _variation(th)
            if l.parent < 0:
                T[i] = _T(base[i]) @ np.linalg.inv(b2j[i]) @ var @ b2j[i]
            else:
                if l.fixed_body2joint_pose:
                    j2p[i] = j2p[i] @ var
                else:
                    b2j[i] = var @ b2j[i]
                T[i] = T[l.parent] @ j2p[i] @ b2j[i]
        return T

    T0 = forward(np.zeros(dof))
    for i in range(4):
        assert np.allclose(T0[i][:3], base[i].reshape(3, 4), atol=2e-6)
    eps = 1e-6
    for k in range(dof):
        th = np.zeros(dof); th[k] = eps
        Tp, Tm = forward(th), forward(-th)
        for i in range(4):
            d = np.real(logm(np.linalg.inv(Tm[i]) @ Tp[i])) / (2 * eps)
            twist = np.array([d[2, 1], d[0, 2], d[1, 0], d[0, 3], d[1, 3], d[2, 3]])
            assert np.allclose(jac[i][:, k], twist, atol=2e-4), (i, k, jac[i][:, k], twist)


def test_one_link_structure_equals_rigid_optimizer(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(13)
    for run in range(5):
        A = rng.normal(size=(6, 6)).astype(np.float32)
        H = -(A @ A.T).astype(np.float32) * 100.0
        g = (rng.normal(size=6) * 10).astype(np.float32)
        pose = _rand_pose(rng)
        p_rigid = pose.reshape(12).copy()
        th_r = np.zeros(6, np.float32)
        L.orc_optimize_rigid(oracle.ptr(g), oracle.ptr(H.reshape(36)), 1000.0, 30000.0, oracle.EXP_PADE,
                             oracle.ptr(p_rigid), oracle.ptr(th_r))
        spec = synth.StructureSpec(links=[synth.LinkSpec(body=-1, parent=-1, body2joint=synth.identity_pose(),
                                                         joint2parent=synth.identity_pose())])
        so = oracle.OracleStructure(spec)
        S = so.as_struct()
        p_s = pose.reshape(1, 12).copy()
        th_s = np.zeros(6, np.float32)
        assert L.orc_optimize_structure(S, oracle.ptr(g), oracle.ptr(H.reshape(36)), oracle.ROTATION_LINEAR,
                                        oracle.EXP_PADE, oracle.ptr(p_s), oracle.ptr(th_s)) == 1
        assert np.array_equal(th_r, th_s)
        assert np.allclose(p_rigid, p_s[0], atol=1e-7)


def test_soft_constraint_gradient_is_energy_gradient(oracle):
    """SoftConstraint: g = -dE/dtheta, H = -(Gauss-Newton Hessian) of E = (|r| - d_max)^2 / (2 sigma^2) per part."""
    L = oracle.lib()
    rng = np.random.default_rng(17)
    for run in range(6):
        dmax_r, dmax_t = (0.0, 0.0) if run % 2 == 0 else (0.05, 0.01)
        sd_r, sd_t = 0.07, 0.013
        dirs = (1, 1, 1, 1, 1, 1) if run < 4 else (0, 1, 1, 1, 0, 1)
        b12j1, b22j2 = _rand_pose(rng, 1.0), _rand_pose(rng, 1.0)
        links = [synth.LinkSpec(body=-1, parent=-1, body2joint=synth.identity_pose(), joint2parent=synth.identity_pose(),
                                free_directions=(0,) * 6),
                 synth.LinkSpec(body=-1, parent=0, body2joint=synth.identity_pose(), joint2parent=_rand_pose(rng)),
                 synth.LinkSpec(body=-1, parent=0, body2joint=synth.identity_pose(), joint2parent=_rand_pose(rng))]
        spec = synth.StructureSpec(links=links, constraints=[synth.ConstraintSpec(
            link1=1, link2=2, body12joint1=b12j1, body22joint2=b22j2, directions=dirs, soft=True,
            max_distance_rotation=dmax_r, max_distance_translation=dmax_t, standard_deviation_rotation=sd_r,
            standard_deviation_translation=sd_t)])
        so = oracle.OracleStructure(spec)
        S = so.as_struct()
        l2w = np.zeros((3, 12), np.float32)
        l2w[0] = synth.identity_pose().reshape(12)
        L.orc_structure_consistent_poses(S, oracle.EXP_PADE, oracle.ptr(l2w))
        g = np.zeros((3, 6), np.float32); H = np.zeros((3, 36), np.float32)
        L.orc_soft_constraint_add(S, 0, oracle.ptr(l2w), oracle.ROTATION_POLAR, oracle.ptr(g), oracle.ptr(H))

        def energy(T1, T2):
            j = _T(b12j1) @ np.linalg.inv(T1) @ T2 @ np.linalg.inv(_T(b22j2))
            r = Rotation.from_matrix(j[:3, :3]).as_rotvec()[[d for d in range(3) if dirs[d]]]
            t = j[:3, 3][[d for d in range(3) if dirs[d + 3]]]
            e = 0.0
            for v, dm, sd in ((r, dmax_r, sd_r), (t, dmax_t, sd_t)):
                n = np.linalg.norm(v)
                if n > dm:
                    e += (n - dm) ** 2 / (2 * sd * sd)
            return e

        T1, T2 = _T(l2w[1]), _T(l2w[2])
        eps = 1e-6
        for li, (A, B) in ((1, (True, False)), (2, (False, True))):
            num = np.zeros(6)
            for k in range(6):
                th = np.zeros(6); th[k] = eps
                Vp, Vm = _variation(th), _variation(-th)
                ep = energy(T1 @ Vp if A else T1, T2 @ Vp if B else T2)
                em = energy(T1 @ Vm if A else T1, T2 @ Vm if B else T2)
                num[k] = -(ep - em) / (2 * eps)
            assert np.allclose(g[li], num, rtol=2e-3, atol=2e-3 * np.abs(num).max()), (run, li, g[li], num)
            Hm = H[li].reshape(6, 6)
            assert np.allclose(Hm, Hm.T, atol=1e-3 * np.abs(Hm).max())
            assert np.linalg.eigvalsh(-(Hm + Hm.T) / 2).min() > -1e-3 * np.abs(Hm).max()


def test_chain_tracking_converges(oracle):
    """config-5 shape on the oracle: both chain variants pull the links back to the ground truth."""
    from helpers import pose_error
    for variant in ("projected", "constrained"):
        wl = synth.make_chain_workload(n_chains=1, n_links=4, n_lines=120, n_points=120, n_divides=3, variant=variant,
                                       seed=2)
        tr = oracle.OracleTracker(wl)
        e0t, e0r = pose_error(wl.start_body2world, wl.gt_body2world)
        tr.start_modalities(0)
        for it in range(3):
            tr.tracking_step(it)
            tr.calculate_results(it)
        e1t, e1r = pose_error(tr.get_poses(), wl.gt_body2world)
        assert np.median(e1t) < 0.5 * np.median(e0t), (variant, e0t, e1t)
        assert np.median(e1r) < 0.5 * np.median(e0r), (variant, e0r, e1r)
        if variant == "constrained":  # the joints stay closed: consecutive links keep the Tx(0.01) offset
            p = tr.get_poses()
            for j in range(1, 4):
                rel = np.linalg.inv(_T(p[j - 1])) @ _T(p[j])
                assert np.linalg.norm(rel[:3, 3] - (0.01, 0, 0)) < 2e-4
                rv = Rotation.from_matrix(rel[:3, :3]).as_rotvec()
                assert abs(rv[1]) < 2e-3 and abs(rv[2]) < 2e-3


def test_link_with_two_modality_sets_sums_them(oracle, synth):
    """Link::CalculateGradientAndHessian adds up all modalities of a link (link.cpp:184-193): a one-link structure whose
    link carries a second body (the same physical body seen by a second camera pair) must take exactly the step of a
    rigid-body optimisation fed with the sum of the four modalities' gradients / Hessians, and both bodies end up with
    the same pose."""
    n = 2
    wl = synth.make_multi_camera_workload(n_objects=n, n_divides=3, seed=8)
    trk = oracle.OracleTracker(wl, rotation_mode=oracle.ROTATION_LINEAR, exp_mode=oracle.EXP_RODRIGUES)
    ref = oracle.OracleTracker(wl, rotation_mode=oracle.ROTATION_LINEAR, exp_mode=oracle.EXP_RODRIGUES)
    trk.start_modalities(0)
    ref.start_modalities(0)
    trk.tracking_step(0, n_corr=1, n_update=1)
    for i in range(n):
        g = np.zeros(6, np.float32)
        H = np.zeros((6, 6), np.float32)
        for b in (i, n + i):                      # modality list order: set A (region, depth), set B (region, depth)
            ref.region_correspondences(b, 0, 0)
            gr, Hr = ref.region_gradient_hessian(b, 0, 0)
            g, H = g + gr, H + Hr
            ref.depth_correspondences(b, 0, 0)
            gd, Hd = ref.depth_gradient_hessian(b, 0)
            g, H = g + gd, H + Hd
        ok, _ = ref.optimize(i, g, H)
        assert ok
    got, want = trk.get_poses(), ref.get_poses()
    assert np.array_equal(got[:n].view(np.uint32), got[n:].view(np.uint32))
    assert np.abs(got[:n] - want[:n]).max() < 2e-6
