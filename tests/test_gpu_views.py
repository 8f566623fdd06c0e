"""Pruned GetClosestView on the device (k_track2): many bodies with random orientations, one correspondence launch,
every selected view must equal the oracle's full scan; repeated with the k_track kernel forced (block-wide scan)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_poses(wl, rng):
    poses = wl.start_body2world.copy()
    for b in range(wl.n_bodies):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        poses[b, :, :3] = R.astype(np.float32)
    return poses


@pytest.mark.parametrize("force_old", [False, True])
def test_closest_views_equal_full_scan(capi, oracle, synth, monkeypatch, force_old):
    if force_old:
        monkeypatch.setenv("M3TB_KERNEL", "1")
    wl = synth.make_workload("c2", n_bodies=96, n_lines=32, n_points=32, n_divides=4, seed=77)
    rng = np.random.default_rng(3)
    ctx = capi.context_from_workload(wl)
    orc = oracle.OracleTracker(wl, rotation_mode=oracle.ROTATION_LINEAR, exp_mode=oracle.EXP_RODRIGUES)
    seen = set()
    for rep in range(4):
        poses = _random_poses(wl, rng)
        ctx.set_poses(poses)
        orc.set_poses(poses)
        ctx.region_correspondences(0, 0)
        ctx.depth_correspondences(0, 0)
        for b in range(wl.n_bodies):
            _, vr = orc.region_correspondences(b, 0, 0)
            _, vd = orc.depth_correspondences(b, 0, 0)
            got = ctx.get_closest_views(b)
            assert got == (vr, vd), (rep, b, got, (vr, vd))
            seen.add(vr)
    assert len(seen) > 200   # the poses really covered the view sphere
    ctx.close()
