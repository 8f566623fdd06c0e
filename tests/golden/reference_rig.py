"""Software re-creation of the reference's modality test rig (M3T/test/common_test.cpp + modality_test.cpp) so that the
oracle can be run on the inputs the reference's golden matrices were computed on.

The known answers data/modality_test/{region_modality_global,region_modality_local,depth_modality}_{gradient,hessian}.txt
and data/optimizer_test/triangle_pose.txt were produced with a sparse viewpoint model that the reference GENERATES at test
time with OpenGL (2000x2000 off-screen render, cv::findContours, std::mt19937{7}) and does not check in.
This module regenerates the one template view those tests use (the view closest to the test pose) without OpenGL:
the body is the convex prism data/_body/triangle.obj, so the render is an exact ray cast; contours come from the same
cv2.findContours; sampling replays std::mt19937{7} (numpy's RandomState(7) yields the identical 32-bit stream).
Pixel-exact agreement with an OpenGL rasteriser is not guaranteed (+-1 px on edges), hence the comparison against the
golden matrices is a SOFT check (a few per cent), not a bit-level gate.

Reference code followed: Model::GenerateGeodesicPoses / SetUpRenderer (src/model.cpp:120-152,386-454),
RegionModel::GeneratePointData / GenerateValidContours / CalculateContourSegment / ApproximateNormalVector /
CalculateLineDistances (src/region_model.cpp:482-784), DepthModel::GeneratePointData / SampleSurfacePointCoordinate
(src/depth_model.cpp:302-352), FullDepthRenderer::PointVector (src/renderer.cpp:445-452).
"""
import json
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "reference_known_answers.json")))

SPHERE_RADIUS = np.float32(0.8)
N_DIVIDES = 4
IMAGE_SIZE = 2000
N_POINTS = 200
K_CONTOUR_NORMAL_APPROX_RADIUS = 3
K_MIN_CONTOUR_LENGTH = 15


def body_vertices():
    v = np.array(KA["triangle_obj"]["vertices"], np.float64) + np.array(KA["triangle_obj"]["geometry2body_translation"])
    f = np.array(KA["triangle_obj"]["faces"], np.int64) - 1
    return v, f


def geodesic_points(n_divides=N_DIVIDES):
    """Model::GenerateGeodesicPoints in float32, std::set ordering (CompareSmallerVector3f: lexicographic)."""
    x, z = np.float32(0.525731112119133606), np.float32(0.850650808352039932)
    o = np.float32(0.0)
    ico = [(-x, o, z), (x, o, z), (-x, o, -z), (x, o, -z), (o, z, x), (o, z, -x), (o, -z, x), (o, -z, -x), (z, x, o), (-z, x, o),
           (z, -x, o), (-z, -x, o)]
    ids = [(0, 4, 1), (0, 9, 4), (9, 5, 4), (4, 5, 8), (4, 8, 1), (8, 10, 1), (8, 3, 10), (5, 3, 8), (5, 2, 3), (2, 7, 3),
           (7, 10, 3), (7, 6, 10), (7, 11, 6), (11, 0, 6), (0, 1, 6), (6, 1, 10), (9, 0, 11), (9, 11, 2), (9, 2, 5), (7, 2, 11)]
    pts = set()

    def norm(v):
        v = np.asarray(v, np.float32)
        return v / np.float32(np.sqrt(np.float32(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))

    def sub(v1, v2, v3, n):
        if n == 0:
            for v in (v1, v2, v3):
                pts.add(tuple(np.asarray(v, np.float32).tolist()))
            return
        v12, v13, v23 = norm(np.float32(v1) + np.float32(v2)), norm(np.float32(v1) + np.float32(v3)), norm(np.float32(v2) + np.float32(v3))
        sub(v1, v12, v13, n - 1); sub(v2, v12, v23, n - 1); sub(v3, v13, v23, n - 1); sub(v12, v13, v23, n - 1)

    for a, b, c in ids:
        sub(np.array(ico[a], np.float32), np.array(ico[b], np.float32), np.array(ico[c], np.float32), n_divides)
    return np.array(sorted(pts), np.float64)


def camera2body_from_point(p):
    """Model::GenerateGeodesicPoses (src/model.cpp:386-411)."""
    p = np.asarray(p, np.float64)
    T = np.eye(4)
    T[:3, 3] = p * float(SPHERE_RADIUS)
    c2 = -p
    if p[0] == 0.0 and p[2] == 0.0:
        c0 = np.array([1.0, 0.0, 0.0])
    else:
        c0 = np.cross([0.0, 1.0, 0.0], c2)
        c0 /= np.linalg.norm(c0)
    c1 = np.cross(c2, c0)
    T[:3, 0], T[:3, 1], T[:3, 2] = c0, c1, c2
    return T


def render(camera2body):
    """Silhouette mask, depth image and per-pixel face normal (camera frame) of the prism for one template view."""
    verts, faces = body_vertices()
    diameter = 2.0 * np.max(np.linalg.norm(verts, axis=1))
    f = 0.5 * (IMAGE_SIZE - 20) / np.tan(np.arcsin(0.5 * diameter / float(SPHERE_RADIUS)))  # SetUpRenderer
    pp = IMAGE_SIZE / 2.0
    b2c = np.linalg.inv(camera2body)
    vc = (b2c[:3, :3] @ verts.T).T + b2c[:3, 3]
    uv = np.stack([f * vc[:, 0] / vc[:, 2] + pp, f * vc[:, 1] / vc[:, 2] + pp], 1)
    depth = np.full((IMAGE_SIZE, IMAGE_SIZE), np.inf)
    normal = np.zeros((IMAGE_SIZE, IMAGE_SIZE, 3))
    for tri in faces:
        A, B, Cc = vc[tri]
        n = np.cross(B - A, Cc - A)
        n /= np.linalg.norm(n)
        centre_c = b2c[:3, 3]
        if n @ (A - centre_c) < 0:
            n = -n
        if n @ A >= 0:  # back face (culling on): normal points away from the camera at the origin
            continue
        a, b, c = uv[tri]
        x0, x1 = int(np.floor(min(a[0], b[0], c[0]))), int(np.ceil(max(a[0], b[0], c[0])))
        y0, y1 = int(np.floor(min(a[1], b[1], c[1]))), int(np.ceil(max(a[1], b[1], c[1])))
        x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, IMAGE_SIZE - 1), min(y1, IMAGE_SIZE - 1)
        xs, ys = np.meshgrid(np.arange(x0, x1 + 1, dtype=np.float64), np.arange(y0, y1 + 1, dtype=np.float64))

        def edge(p, q):
            return (q[0] - p[0]) * (ys - p[1]) - (q[1] - p[1]) * (xs - p[0])
        e0, e1, e2 = edge(a, b), edge(b, c), edge(c, a)
        inside = ((e0 >= 0) & (e1 >= 0) & (e2 >= 0)) | ((e0 <= 0) & (e1 <= 0) & (e2 <= 0))
        # ray through pixel (u, v): d = ((u-pp)/f, (v-pp)/f, 1); plane n.x = n.A
        dz = (n @ A) / (n[0] * (xs - pp) / f + n[1] * (ys - pp) / f + n[2])
        sub_d = depth[y0:y1 + 1, x0:x1 + 1]
        upd = inside & (dz < sub_d)
        sub_d[upd] = dz[upd]
        normal[y0:y1 + 1, x0:x1 + 1][upd] = n
    mask = np.where(np.isfinite(depth), 255, 0).astype(np.uint8)
    return mask, depth, normal, f, pp


def closest_view_pose(body2camera):
    """The geodesic view GetClosestView selects for this body2camera pose, and all view orientations."""
    pts = geodesic_points()
    orientations = -pts  # camera2body rotation column 2
    t = body2camera[:3, 3]
    o = body2camera[:3, :3].T @ (t / np.linalg.norm(t))
    dots = orientations @ o
    best = int(np.argmax(dots))
    return best, orientations, camera2body_from_point(pts[best]), float(np.sort(dots)[-1] - np.sort(dots)[-2])


def region_view_points(camera2body):
    """RegionModel::GeneratePointData for one view -> [N_POINTS, 38] float32 DataPoints + contour_length."""
    mask, depth, _, f, pp = render(camera2body)
    contours, _ = cv2.findContours(mask, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    contours = [c.reshape(-1, 2) for c in contours if len(c) >= K_MIN_CONTOUR_LENGTH]
    valid = np.concatenate(contours, 0)
    pixel_to_meter0 = float(SPHERE_RADIUS) / f
    contour_length = len(valid) * pixel_to_meter0
    rs = np.random.RandomState(7)  # std::mt19937 generator{7}

    def gen():
        return int(rs.randint(0, 2 ** 32, dtype=np.uint64))

    all_pts = np.concatenate(contours, 0).astype(np.float32)
    out = np.zeros((N_POINTS, 38), np.float32)
    k = 0
    tries = 0
    while k < N_POINTS:
        tries += 1
        assert tries < 100
        center = valid[gen() % len(valid)]
        seg = None
        for c in contours:  # CalculateContourSegment
            hit = np.nonzero((c[:, 0] == center[0]) & (c[:, 1] == center[1]))[0]
            if len(hit):
                idx = int(hit[0])
                n = len(c)
                r = K_CONTOUR_NORMAL_APPROX_RADIUS
                ids = [(idx + d) % n for d in range(-r, r + 1)]
                seg = c[ids]
                break
        if float(np.hypot(*(seg[-1] - seg[0]).astype(np.float32))) <= K_CONTOUR_NORMAL_APPROX_RADIUS:
            continue
        nrm = np.array([-(seg[-1][1] - seg[0][1]), seg[-1][0] - seg[0][0]], np.float64)  # ApproximateNormalVector
        nrm /= np.linalg.norm(nrm)
        d = depth[center[1], center[0]]
        center_c = np.array([d * (center[0] - pp) / f, d * (center[1] - pp) / f, d])
        pixel_to_meter = center_c[2] / f
        # CalculateLineDistances
        if abs(nrm[1]) < abs(nrm[0]):
            u_step, v_step = float(np.sign(nrm[0])), nrm[1] / abs(nrm[0])
        else:
            u_step, v_step = nrm[0] / abs(nrm[1]), float(np.sign(nrm[1]))
        u_in, v_in = center[0] + 0.5, center[1] + 0.5
        while True:
            u_in -= u_step; v_in -= v_step
            if mask[int(v_in), int(u_in)] != 255:
                dd = np.hypot(all_pts[:, 0] - (u_in + u_step - 0.5), all_pts[:, 1] - (v_in + v_step - 0.5))
                e = all_pts[int(np.argmin(dd))]
                fg = pixel_to_meter * float(np.hypot(e[0] - center[0], e[1] - center[1]))
                break
        u_out, v_out = center[0] + 0.5, center[1] + 0.5
        while True:
            u_out += u_step; v_out += v_step
            if int(u_out) < 0 or int(u_out) >= IMAGE_SIZE or int(v_out) < 0 or int(v_out) >= IMAGE_SIZE:
                bg = np.finfo(np.float32).max
                break
            if mask[int(v_out), int(u_out)] == 255:
                dd = np.hypot(all_pts[:, 0] - (u_out - 0.5), all_pts[:, 1] - (v_out - 0.5))
                e = all_pts[int(np.argmin(dd))]
                bg = pixel_to_meter * float(np.hypot(e[0] - center[0], e[1] - center[1]))
                break
        out[k, 0:3] = (camera2body[:3, :3] @ center_c + camera2body[:3, 3]).astype(np.float32)
        out[k, 3:6] = (camera2body[:3, :3] @ np.array([nrm[0], nrm[1], 0.0])).astype(np.float32)
        out[k, 6], out[k, 7] = fg, bg
        k += 1
        tries = 0
    return out, np.float32(contour_length)


def depth_view_points(camera2body):
    """DepthModel::GeneratePointData for one view -> [N_POINTS, 36] float32 DataPoints + surface_area."""
    mask, depth, normal, f, pp = render(camera2body)
    surface_area = np.count_nonzero(mask) * (float(SPHERE_RADIUS) / f) ** 2
    rs = np.random.RandomState(7)
    out = np.zeros((N_POINTS, 36), np.float32)
    n_pixels = IMAGE_SIZE * IMAGE_SIZE
    k = 0
    while k < N_POINTS:
        idx = int(rs.randint(0, 2 ** 32, dtype=np.uint64)) % n_pixels
        x, y = idx // IMAGE_SIZE, idx % IMAGE_SIZE  # sic: coordinate{idx / rows, idx % cols}
        if not mask[y, x]:
            continue
        d = depth[y, x]
        center_c = np.array([d * (x - pp) / f, d * (y - pp) / f, d])
        # NormalVector: decoded from an 8-bit normal image, not renormalised
        n = normal[y, x]
        q = np.round((1.0 - n) * 127.5)
        n8 = 1.0 - q / 127.5
        out[k, 0:3] = (camera2body[:3, :3] @ center_c + camera2body[:3, 3]).astype(np.float32)
        out[k, 3:6] = (camera2body[:3, :3] @ n8).astype(np.float32)
        k += 1
    return out, np.float32(surface_area)


def rig():
    """Everything the oracle needs for RegionModalityTest / DepthModalityTest / OptimizerTest."""
    w2b = np.array(KA["triangle_world2body"], np.float64)
    b2w = np.linalg.inv(w2b)
    cc, dc = KA["color_camera"], KA["depth_camera"]
    color_w2c = np.linalg.inv(np.array(cc["camera2world"], np.float64))
    depth_w2c = np.linalg.inv(np.array(dc["camera2world"], np.float64))
    color = cv2.imread(os.path.join(HERE, "color_camera_image_200.png"), cv2.IMREAD_COLOR)
    depth = cv2.imread(os.path.join(HERE, "depth_camera_image_200.png"), cv2.IMREAD_UNCHANGED)
    assert color.shape == (540, 960, 3) and depth.shape == (480, 848) and depth.dtype == np.uint16
    return dict(body2world=b2w, color_w2c=color_w2c, depth_w2c=depth_w2c, color=np.ascontiguousarray(color),
                depth=np.ascontiguousarray(depth), cc=cc, dc=dc)


def make_fixtures():
    """Regenerates tests/golden/triangle_test_view.npz (run in the build container; needs cv2)."""
    r = rig()
    out = {}
    for kind, w2c in (("region", r["color_w2c"]), ("depth", r["depth_w2c"])):
        b2c = w2c @ r["body2world"]
        best, orientations, c2b, margin = closest_view_pose(b2c)
        pts, scalar = (region_view_points if kind == "region" else depth_view_points)(c2b)
        out[f"{kind}_view"] = np.int32(best)
        out[f"{kind}_points"] = pts
        out[f"{kind}_scalar"] = scalar
        out[f"{kind}_argmax_margin"] = np.float32(margin)
        out["orientations"] = orientations.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "triangle_test_view.npz"), **out)
    return out


if __name__ == "__main__":
    o = make_fixtures()
    print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in o.items()})
