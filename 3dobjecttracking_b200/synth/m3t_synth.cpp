// m3t_synth.cpp — analytic synthetic inputs (see m3t_synth.h). Geometry in double, outputs float.
#include "m3t_synth.h"

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace {

struct V3 {
  double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double Dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 Cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double Norm(V3 a) { return std::sqrt(Dot(a, a)); }
inline V3 Normalized(V3 a) {
  double n = Norm(a);
  return n > 0 ? (1.0 / n) * a : a;
}

struct Pose {  // x_out = R x + t
  double R[3][3];
  V3 t;
};
inline V3 Apply(const Pose& p, V3 v) {
  return {p.R[0][0] * v.x + p.R[0][1] * v.y + p.R[0][2] * v.z + p.t.x,
          p.R[1][0] * v.x + p.R[1][1] * v.y + p.R[1][2] * v.z + p.t.y,
          p.R[2][0] * v.x + p.R[2][1] * v.y + p.R[2][2] * v.z + p.t.z};
}
inline V3 Rotate(const Pose& p, V3 v) {
  return {p.R[0][0] * v.x + p.R[0][1] * v.y + p.R[0][2] * v.z, p.R[1][0] * v.x + p.R[1][1] * v.y + p.R[1][2] * v.z,
          p.R[2][0] * v.x + p.R[2][1] * v.y + p.R[2][2] * v.z};
}
inline Pose InverseRigid(const Pose& p) {
  Pose r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.R[i][j] = p.R[j][i];
  V3 t = Rotate(r, p.t);
  r.t = {-t.x, -t.y, -t.z};
  return r;
}
inline Pose FromFloat12(const float* f) {
  Pose p;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) p.R[i][j] = f[4 * i + j];
  }
  p.t = {f[3], f[7], f[11]};
  return p;
}
inline void ToFloat12(const Pose& p, float* f) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) f[4 * i + j] = float(p.R[i][j]);
  f[3] = float(p.t.x); f[7] = float(p.t.y); f[11] = float(p.t.z);
}

// ---- body: the reference's triangle prism (data/_body/triangle.obj: 6 vertices, 8 triangles, metres;
// geometry2body_pose translates z by -0.006, data/_body/triangle.yaml) -------------------------------
constexpr int kNV = 6, kNF = 8;
const V3 kVerts[kNV] = {{-0.038305, 0.0, -0.006},     {-0.038305, 0.0, 0.006},     {0.019152, -0.033231, -0.006},
                        {0.019152, -0.033231, 0.006}, {0.019152, 0.033231, -0.006}, {0.019152, 0.033231, 0.006}};
const int kFaces[kNF][3] = {{0, 2, 3}, {2, 4, 3}, {3, 5, 1}, {4, 0, 1}, {0, 4, 2}, {1, 0, 3}, {4, 5, 3}, {5, 4, 1}};

V3 FaceNormal(int f) {
  V3 a = kVerts[kFaces[f][0]], b = kVerts[kFaces[f][1]], c = kVerts[kFaces[f][2]];
  V3 n = Normalized(Cross(b - a, c - a));
  V3 centroid = (1.0 / 3.0) * (a + b + c);
  if (Dot(n, centroid) < 0) n = -1.0 * n;  // body is convex and contains the origin -> outward
  return n;
}

// ---- counter-based PRNG -------------------------------------------------------------------------
inline uint64_t Mix(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t Hash(uint64_t seed, uint64_t a, uint64_t b = 0, uint64_t c = 0) {
  return Mix(Mix(Mix(Mix(seed) ^ a) ^ (b * 0xD1B54A32D192ED03ull)) ^ (c * 0x8CB92BA72F3D8DD7ull));
}
inline double U01(uint64_t h) { return double(h >> 11) * (1.0 / 9007199254740992.0); }
inline double Gauss4(uint64_t h) {  // Irwin-Hall(4) -> approx N(0,1)
  double s = double(h & 0xFFFF) + double((h >> 16) & 0xFFFF) + double((h >> 32) & 0xFFFF) + double((h >> 48) & 0xFFFF);
  return (s * (1.0 / 65536.0) - 2.0) * 1.7320508075688772;
}

// ---- geodesic view grid (model.cpp:386-454) -------------------------------------------------------
struct KeyLess {
  bool operator()(const std::array<float, 3>& a, const std::array<float, 3>& b) const {
    // ordering of CompareSmallerVector3f (model.h): lexicographic on (x, y, z) of the float coordinates
    if (a[0] != b[0]) return a[0] < b[0];
    if (a[1] != b[1]) return a[1] < b[1];
    return a[2] < b[2];
  }
};
void Subdivide(V3 v1, V3 v2, V3 v3, int n, std::map<std::array<float, 3>, V3, KeyLess>* pts) {
  auto ins = [&](V3 v) {
    // deduplicate on rounded float coordinates (points shared between triangles agree to ~1e-16 in double)
    std::array<float, 3> k = {float(std::round(v.x * 1e6) / 1e6), float(std::round(v.y * 1e6) / 1e6),
                              float(std::round(v.z * 1e6) / 1e6)};
    pts->emplace(k, v);
  };
  if (n == 0) {
    ins(v1); ins(v2); ins(v3);
  } else {
    V3 v12 = Normalized(v1 + v2), v13 = Normalized(v1 + v3), v23 = Normalized(v2 + v3);
    Subdivide(v1, v12, v13, n - 1, pts);
    Subdivide(v2, v12, v23, n - 1, pts);
    Subdivide(v3, v13, v23, n - 1, pts);
    Subdivide(v12, v13, v23, n - 1, pts);
  }
}
std::vector<V3> GeodesicPoints(int n_divides) {
  const double x = 0.525731112119133606, z = 0.850650808352039932;
  const V3 ico[12] = {{-x, 0, z}, {x, 0, z}, {-x, 0, -z}, {x, 0, -z}, {0, z, x},  {0, z, -x},
                      {0, -z, x}, {0, -z, -x}, {z, x, 0}, {-z, x, 0}, {z, -x, 0}, {-z, -x, 0}};
  const int ids[20][3] = {{0, 4, 1},  {0, 9, 4},  {9, 5, 4},  {4, 5, 8},  {4, 8, 1},  {8, 10, 1}, {8, 3, 10},
                          {5, 3, 8},  {5, 2, 3},  {2, 7, 3},  {7, 10, 3}, {7, 6, 10}, {7, 11, 6}, {11, 0, 6},
                          {0, 1, 6},  {6, 1, 10}, {9, 0, 11}, {9, 11, 2}, {9, 2, 5},  {7, 2, 11}};
  std::map<std::array<float, 3>, V3, KeyLess> pts;
  for (auto& t : ids) Subdivide(ico[t[0]], ico[t[1]], ico[t[2]], n_divides, &pts);
  std::vector<V3> out;
  out.reserve(pts.size());
  for (auto& kv : pts) out.push_back(kv.second);
  return out;
}
Pose GeodesicCamera2Body(V3 p, double sphere_radius) {  // model.cpp:386-411
  Pose pose;
  pose.t = sphere_radius * p;
  V3 c2 = -1.0 * p;
  V3 c0;
  if (std::fabs(p.x) < 1e-12 && std::fabs(p.z) < 1e-12)
    c0 = {1, 0, 0};
  else
    c0 = Normalized(Cross(V3{0, 1, 0}, c2));
  V3 c1 = Cross(c2, c0);
  const V3 cols[3] = {c0, c1, c2};
  for (int j = 0; j < 3; ++j) {
    pose.R[0][j] = cols[j].x; pose.R[1][j] = cols[j].y; pose.R[2][j] = cols[j].z;
  }
  return pose;
}

// ---- convex silhouette ---------------------------------------------------------------------------
struct P2 {
  double x, y;
  int vid;
};
inline double Cross2(const P2& o, const P2& a, const P2& b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); }
std::vector<P2> ConvexHull(std::vector<P2> pts) {  // Andrew monotone chain, counter-clockwise in (x, y)
  std::sort(pts.begin(), pts.end(), [](const P2& a, const P2& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  int n = int(pts.size()), k = 0;
  std::vector<P2> h(2 * n);
  for (int i = 0; i < n; ++i) {
    while (k >= 2 && Cross2(h[k - 2], h[k - 1], pts[i]) <= 0) k--;
    h[k++] = pts[i];
  }
  for (int i = n - 2, t = k + 1; i >= 0; --i) {
    while (k >= t && Cross2(h[k - 2], h[k - 1], pts[i]) <= 0) k--;
    h[k++] = pts[i];
  }
  h.resize(k - 1);
  return h;
}
// Silhouette polygon of the prism under body2camera, in coordinates (fx*X/Z + cx, fy*Y/Z + cy).
std::vector<P2> Silhouette(const Pose& body2camera, double fx, double fy, double cx, double cy, V3* cam_verts) {
  std::vector<P2> pts;
  for (int i = 0; i < kNV; ++i) {
    V3 c = Apply(body2camera, kVerts[i]);
    if (cam_verts) cam_verts[i] = c;
    pts.push_back({fx * c.x / c.z + cx, fy * c.y / c.z + cy, i});
  }
  return ConvexHull(pts);
}
inline bool InsideHull(const std::vector<P2>& h, double x, double y) {
  int n = int(h.size());
  for (int i = 0; i < n; ++i) {
    const P2& a = h[i];
    const P2& b = h[(i + 1) % n];
    if ((b.x - a.x) * (y - a.y) - (b.y - a.y) * (x - a.x) < 0) return false;
  }
  return true;
}

}  // namespace

extern "C" {

int m3ts_n_views(int n_divides) {
  int n = 10;
  for (int i = 0; i < n_divides; ++i) n *= 4;
  return n + 2;
}

int m3ts_generate_region_model(int n_divides, int n_points, float sphere_radius, uint64_t seed, float* orientations,
                               float* contour_lengths, void* points_out) {
  std::vector<V3> geo = GeodesicPoints(n_divides);
  int nv = int(geo.size());
  if (nv != m3ts_n_views(n_divides)) return -1;
  float* out = static_cast<float*>(points_out);
  const int kF = 38;
#pragma omp parallel for schedule(static)
  for (int v = 0; v < nv; ++v) {
    Pose c2b = GeodesicCamera2Body(geo[v], sphere_radius);
    Pose b2c = InverseRigid(c2b);
    // orientation = camera2body rotation column 2 = direction camera -> body centre (region_model.cpp:236)
    orientations[3 * v + 0] = float(c2b.R[0][2]);
    orientations[3 * v + 1] = float(c2b.R[1][2]);
    orientations[3 * v + 2] = float(c2b.R[2][2]);
    V3 cam[kNV];
    std::vector<P2> hull = Silhouette(b2c, 1.0, 1.0, 0.0, 0.0, cam);
    int nh = int(hull.size());
    std::vector<double> cum(nh + 1, 0.0);
    double cxm = 0, cym = 0;
    for (int i = 0; i < nh; ++i) {
      const P2& a = hull[i];
      const P2& b = hull[(i + 1) % nh];
      cum[i + 1] = cum[i] + std::hypot(b.x - a.x, b.y - a.y);
      cxm += a.x; cym += a.y;
    }
    cxm /= nh; cym /= nh;
    double perimeter = cum[nh];
    contour_lengths[v] = float(perimeter * sphere_radius);
    for (int k = 0; k < n_points; ++k) {
      float* dp = out + (size_t(v) * n_points + k) * kF;
      std::memset(dp, 0, kF * sizeof(float));
      double s = U01(Hash(seed, 0x5245, uint64_t(v), uint64_t(k))) * perimeter;
      int e = int(std::upper_bound(cum.begin(), cum.end(), s) - cum.begin()) - 1;
      e = std::min(std::max(e, 0), nh - 1);
      const P2& a = hull[e];
      const P2& b = hull[(e + 1) % nh];
      double len = cum[e + 1] - cum[e];
      double f = len > 0 ? (s - cum[e]) / len : 0.0;
      f = std::min(std::max(f, 1e-3), 1.0 - 1e-3);  // stay off the corners
      // perspective-correct position on the 3D edge
      V3 A = cam[a.vid], B = cam[b.vid];
      double t = f * A.z / (f * A.z + (1.0 - f) * B.z);
      V3 X = A + t * (B - A);
      double qx = X.x / X.z, qy = X.y / X.z;
      // outward 2D normal of the edge
      double nx = b.y - a.y, ny = -(b.x - a.x);
      double nn = std::hypot(nx, ny);
      nx /= nn; ny /= nn;
      if ((qx - cxm) * nx + (qy - cym) * ny < 0) { nx = -nx; ny = -ny; }
      // foreground distance: chord of the silhouette from q along -n
      double chord = 0.0;
      for (int i = 0; i < nh; ++i) {
        if (i == e) continue;
        const P2& c = hull[i];
        const P2& d = hull[(i + 1) % nh];
        double ex = d.x - c.x, ey = d.y - c.y;
        double den = (-nx) * ey - (-ny) * ex;
        if (std::fabs(den) < 1e-15) continue;
        double tt = ((c.x - qx) * ey - (c.y - qy) * ex) / den;      // along -n
        double uu = ((c.x - qx) * (-ny) - (c.y - qy) * (-nx)) / den; // along the edge
        if (tt > 1e-9 && uu >= -1e-9 && uu <= 1.0 + 1e-9) chord = std::max(chord, tt);
      }
      V3 cb = Apply(c2b, X);
      V3 nb = Rotate(c2b, V3{nx, ny, 0.0});
      dp[0] = float(cb.x); dp[1] = float(cb.y); dp[2] = float(cb.z);
      dp[3] = float(nb.x); dp[4] = float(nb.y); dp[5] = float(nb.z);
      dp[6] = float(chord * X.z);  // foreground_distance in metres at the point's depth
      dp[7] = FLT_MAX;             // background_distance: nothing behind (as in data/model_test/region_model.bin)
    }
  }
  return nv;
}

int m3ts_generate_depth_model(int n_divides, int n_points, float sphere_radius, uint64_t seed, float* orientations,
                              float* surface_areas, void* points_out) {
  std::vector<V3> geo = GeodesicPoints(n_divides);
  int nv = int(geo.size());
  if (nv != m3ts_n_views(n_divides)) return -1;
  float* out = static_cast<float*>(points_out);
  const int kF = 36;
  V3 fn[kNF];
  for (int f = 0; f < kNF; ++f) fn[f] = FaceNormal(f);
#pragma omp parallel for schedule(static)
  for (int v = 0; v < nv; ++v) {
    Pose c2b = GeodesicCamera2Body(geo[v], sphere_radius);
    Pose b2c = InverseRigid(c2b);
    orientations[3 * v + 0] = float(c2b.R[0][2]);
    orientations[3 * v + 1] = float(c2b.R[1][2]);
    orientations[3 * v + 2] = float(c2b.R[2][2]);
    // visible faces weighted by projected area
    double w[kNF], cum[kNF + 1];
    cum[0] = 0;
    for (int f = 0; f < kNF; ++f) {
      V3 a = kVerts[kFaces[f][0]], b = kVerts[kFaces[f][1]], c = kVerts[kFaces[f][2]];
      V3 centroid = (1.0 / 3.0) * (a + b + c);
      V3 to_cam = c2b.t - centroid;
      double facing = Dot(fn[f], Normalized(to_cam));
      double area = 0.5 * Norm(Cross(b - a, c - a));
      w[f] = facing > 0.05 ? area * facing : 0.0;
      cum[f + 1] = cum[f] + w[f];
    }
    surface_areas[v] = float(cum[kNF]);
    for (int k = 0; k < n_points; ++k) {
      float* dp = out + (size_t(v) * n_points + k) * kF;
      std::memset(dp, 0, kF * sizeof(float));
      double s = U01(Hash(seed, 0x4445, uint64_t(v), uint64_t(k))) * cum[kNF];
      int f = 0;
      while (f < kNF - 1 && s >= cum[f + 1]) ++f;
      double r1 = U01(Hash(seed, 0x4446, uint64_t(v), uint64_t(k)));
      double r2 = U01(Hash(seed, 0x4447, uint64_t(v), uint64_t(k)));
      double sq = std::sqrt(r1);
      double b0 = 1.0 - sq, b1 = sq * (1.0 - r2), b2 = sq * r2;
      V3 X = b0 * kVerts[kFaces[f][0]] + b1 * kVerts[kFaces[f][1]] + b2 * kVerts[kFaces[f][2]];
      dp[0] = float(X.x); dp[1] = float(X.y); dp[2] = float(X.z);
      dp[3] = float(fn[f].x); dp[4] = float(fn[f].y); dp[5] = float(fn[f].z);
    }
    (void)b2c;
  }
  return nv;
}

void m3ts_ground_truth_pose(uint64_t seed, int index, const m3ts_intrinsics* intr, float margin_px, float z_min,
                            float z_max, float body2camera[12]) {
  // uniform random rotation from a seeded unit quaternion
  double q[4];
  double n2 = 0;
  for (int i = 0; i < 4; ++i) {
    q[i] = Gauss4(Hash(seed, 0x5054, uint64_t(index), uint64_t(i)));
    n2 += q[i] * q[i];
  }
  if (n2 < 1e-12) { q[0] = 1; q[1] = q[2] = q[3] = 0; n2 = 1; }
  double inv = 1.0 / std::sqrt(n2);
  double w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
  Pose p;
  p.R[0][0] = 1 - 2 * (y * y + z * z); p.R[0][1] = 2 * (x * y - z * w);     p.R[0][2] = 2 * (x * z + y * w);
  p.R[1][0] = 2 * (x * y + z * w);     p.R[1][1] = 1 - 2 * (x * x + z * z); p.R[1][2] = 2 * (y * z - x * w);
  p.R[2][0] = 2 * (x * z - y * w);     p.R[2][1] = 2 * (y * z + x * w);     p.R[2][2] = 1 - 2 * (x * x + y * y);
  double zc = z_min + (z_max - z_min) * U01(Hash(seed, 0x5055, uint64_t(index)));
  double u = margin_px + (intr->width - 1 - 2.0 * margin_px) * U01(Hash(seed, 0x5056, uint64_t(index)));
  double v = margin_px + (intr->height - 1 - 2.0 * margin_px) * U01(Hash(seed, 0x5057, uint64_t(index)));
  p.t = {(u - intr->ppu) / intr->fu * zc, (v - intr->ppv) / intr->fv * zc, zc};
  ToFloat12(p, body2camera);
}

void m3ts_perturb_pose(uint64_t seed, int index, float rot_deg, float trans_m, const float in[12], float out[12]) {
  V3 axis = Normalized(V3{Gauss4(Hash(seed, 0x5058, uint64_t(index), 0)), Gauss4(Hash(seed, 0x5058, uint64_t(index), 1)),
                          Gauss4(Hash(seed, 0x5058, uint64_t(index), 2))});
  V3 dir = Normalized(V3{Gauss4(Hash(seed, 0x5059, uint64_t(index), 0)), Gauss4(Hash(seed, 0x5059, uint64_t(index), 1)),
                         Gauss4(Hash(seed, 0x5059, uint64_t(index), 2))});
  if (Norm(axis) == 0) axis = {0, 0, 1};
  if (Norm(dir) == 0) dir = {1, 0, 0};
  double a = rot_deg * 3.14159265358979323846 / 180.0;
  double c = std::cos(a), s = std::sin(a), C = 1 - c;
  Pose d;
  d.R[0][0] = c + axis.x * axis.x * C;          d.R[0][1] = axis.x * axis.y * C - axis.z * s; d.R[0][2] = axis.x * axis.z * C + axis.y * s;
  d.R[1][0] = axis.y * axis.x * C + axis.z * s; d.R[1][1] = c + axis.y * axis.y * C;          d.R[1][2] = axis.y * axis.z * C - axis.x * s;
  d.R[2][0] = axis.z * axis.x * C - axis.y * s; d.R[2][1] = axis.z * axis.y * C + axis.x * s; d.R[2][2] = c + axis.z * axis.z * C;
  d.t = double(trans_m) * dir;
  Pose p = FromFloat12(in);
  Pose r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.R[i][j] = p.R[i][0] * d.R[0][j] + p.R[i][1] * d.R[1][j] + p.R[i][2] * d.R[2][j];
  r.t = Apply(p, d.t);
  ToFloat12(r, out);
}

void m3ts_render_color(const m3ts_intrinsics* intr, const float body2camera[12], uint64_t seed, const uint8_t fg_mean[3],
                       const uint8_t bg_mean[3], float sigma, uint8_t* bgr, size_t pitch) {
  Pose b2c = FromFloat12(body2camera);
  std::vector<P2> hull = Silhouette(b2c, intr->fu, intr->fv, intr->ppu, intr->ppv, nullptr);
  double x0 = 1e30, x1 = -1e30, y0 = 1e30, y1 = -1e30;
  for (auto& h : hull) {
    x0 = std::min(x0, h.x); x1 = std::max(x1, h.x);
    y0 = std::min(y0, h.y); y1 = std::max(y1, h.y);
  }
  const int W = intr->width, H = intr->height;
#pragma omp parallel for schedule(static)
  for (int v = 0; v < H; ++v) {
    uint8_t* row = bgr + size_t(v) * pitch;
    for (int u = 0; u < W; ++u) {
      bool fg = u >= x0 && u <= x1 && v >= y0 && v <= y1 && InsideHull(hull, double(u), double(v));
      const uint8_t* mean = fg ? fg_mean : bg_mean;
      uint64_t pix = uint64_t(v) * uint64_t(W) + uint64_t(u);
      for (int c = 0; c < 3; ++c) {
        double val = double(mean[c]) + double(sigma) * Gauss4(Hash(seed, 0x434F, pix, uint64_t(c)));
        int iv = int(std::lround(val));
        row[3 * u + c] = uint8_t(std::min(255, std::max(0, iv)));
      }
    }
  }
}

void m3ts_render_depth(const m3ts_intrinsics* intr, const float body2camera[12], uint64_t seed, float background_z,
                       float noise_sigma, float invalid_fraction, float depth_scale, uint16_t* depth, size_t pitch) {
  Pose b2c = FromFloat12(body2camera);
  std::vector<P2> hull = Silhouette(b2c, intr->fu, intr->fv, intr->ppu, intr->ppv, nullptr);
  double x0 = 1e30, x1 = -1e30, y0 = 1e30, y1 = -1e30;
  for (auto& h : hull) {
    x0 = std::min(x0, h.x); x1 = std::max(x1, h.x);
    y0 = std::min(y0, h.y); y1 = std::max(y1, h.y);
  }
  // face planes in camera coordinates: n . x = c
  V3 pn[kNF];
  double pc[kNF];
  for (int f = 0; f < kNF; ++f) {
    pn[f] = Rotate(b2c, FaceNormal(f));
    pc[f] = Dot(pn[f], Apply(b2c, kVerts[kFaces[f][0]]));
  }
  const int W = intr->width, H = intr->height;
#pragma omp parallel for schedule(static)
  for (int v = 0; v < H; ++v) {
    uint16_t* row = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(depth) + size_t(v) * pitch);
    for (int u = 0; u < W; ++u) {
      double z = background_z;
      if (u >= x0 - 1 && u <= x1 + 1 && v >= y0 - 1 && v <= y1 + 1) {
        V3 d = {(u - intr->ppu) / intr->fu, (v - intr->ppv) / intr->fv, 1.0};
        double t_enter = 0.0, t_exit = 1e30;
        bool miss = false;
        for (int f = 0; f < kNF; ++f) {
          double nd = Dot(pn[f], d);
          if (std::fabs(nd) < 1e-12) {
            if (pc[f] < 0) miss = true;
            continue;
          }
          double t = pc[f] / nd;
          if (nd < 0) t_enter = std::max(t_enter, t);
          else t_exit = std::min(t_exit, t);
        }
        if (!miss && t_enter <= t_exit && t_enter > 0) z = t_enter;
      }
      uint64_t pix = uint64_t(v) * uint64_t(W) + uint64_t(u);
      z += double(noise_sigma) * Gauss4(Hash(seed, 0x4450, pix));
      double raw = z / depth_scale;
      int iv = int(std::lround(raw));
      if (U01(Hash(seed, 0x4451, pix)) < invalid_fraction) iv = 0;
      row[u] = uint16_t(std::min(65535, std::max(0, iv)));
    }
  }
}

}  // extern "C"
