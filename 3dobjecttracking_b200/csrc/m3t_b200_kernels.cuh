// m3t_b200_kernels.cuh — the kernels of the pose-optimisation path.
//
//   k_track      one CTA per body: K5 closest view -> K1 region lines (project, validate, DDA gather,
//                normalised-LUT lookup, segment products, distribution, moments) + K3 depth
//                correspondence search -> n_update x ( K2 gradient/Hessian accumulate with warp-shuffle
//                reduction -> K4 6x6 pivoted LDL^T + SE(3) exponential update ), for corr in
//                [corr_begin, corr_end). The same kernel serves the fine-grained C-ABI calls via `phases`.
//   k_histogram  RegionModality::StartModality / CalculateResults histogram side.
//   k_lut        per-bin normalisation of (hist_f, hist_b) -> float2 LUT.
#pragma once

#include "m3t_b200_device.cuh"

namespace m3tb {

struct Shared {
  float pose[12];            // body2world (Body::body2world_pose)
  float red[kWarps][54];     // per-warp partial sums: region g[6]+H[21], depth g[6]+H[21]
  float gh[54];              // block totals
  float a[36];               // normal matrix (lower)
  float b[6];
  float best_dot[kWarps];
  int best_idx[kWarps];
  int view[2];               // closest view: region, depth
  int n_items[2];            // n_lines, n_points
};

// index of (i, j), i >= j, in the packed lower triangle
__host__ __device__ __forceinline__ constexpr int Tri(int i, int j) { return i * (i + 1) / 2 + j; }

// ---------------------------------------------------------------------------------------------
// K5: RegionModel/DepthModel::GetClosestView (region_model.cpp:105-130, depth_model.cpp:81-106)
// orientation = R^T * normalize(t) (linear block of body2camera, see DESIGN.md "Numerics");
// argmax of the dot product, first maximum wins.
// ---------------------------------------------------------------------------------------------
__device__ int ClosestView(const ModelDev& m, const float* b2c, Shared& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float tx = b2c[3], ty = b2c[7], tz = b2c[11];
  float z = tx * tx + ty * ty + tz * tz;
  float norm = sqrtf(z);
  if (norm == 0.0f) return 0;  // uniform across the block
  if (z > 0.0f) { tx /= norm; ty /= norm; tz /= norm; }
  float o0 = b2c[0] * tx + b2c[4] * ty + b2c[8] * tz;
  float o1 = b2c[1] * tx + b2c[5] * ty + b2c[9] * tz;
  float o2 = b2c[2] * tx + b2c[6] * ty + b2c[10] * tz;
  float best = -1.0f;
  int idx = 0x7fffffff;
  for (int v = tid; v < m.n_views; v += kBlockThreads) {
    const float* vo = m.orientations + 3 * v;
    float dot = o0 * __ldg(vo) + o1 * __ldg(vo + 1) + o2 * __ldg(vo + 2);
    if (dot > best) { best = dot; idx = v; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    float ob = __shfl_down_sync(0xffffffffu, best, off);
    int oi = __shfl_down_sync(0xffffffffu, idx, off);
    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
  if (lane == 0) { sh.best_dot[warp] = best; sh.best_idx[warp] = idx; }
  __syncthreads();
  float rb = sh.best_dot[0];
  int ri = sh.best_idx[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) {
    float ob = sh.best_dot[w];
    int oi = sh.best_idx[w];
    if (ob > rb || (ob == rb && oi < ri)) { rb = ob; ri = oi; }
  }
  __syncthreads();  // best_* reused by the next call
  return ri == 0x7fffffff ? 0 : ri;
}

// ---------------------------------------------------------------------------------------------
// K1: one correspondence line (region_modality.cpp:1231-1291, 1433-1658)
// ---------------------------------------------------------------------------------------------
struct RegionIter {  // PrecalculateCameraVariables / PoseVariables / IterationDependentVariables
  float b2c[12];
  float fu, fv, ppu, ppv;
  int w_m1, h_m1, w_m2, h_m2;
  int scale;
  float fscale;
  int ll_m1;
  float ll_m1_half, ll_half_m1;
  float variance;
};

__device__ __forceinline__ void MakeRegionIter(const RegionParamsDev& rp, const CameraDev& cam, const float* pose,
                                               int corr_iteration, RegionIter& it) {
  PoseMul(cam.w2c, pose, it.b2c);  // region_modality.cpp:1001-1002
  it.fu = cam.fu; it.fv = cam.fv; it.ppu = cam.ppu; it.ppv = cam.ppv;
  it.w_m1 = cam.width - 1; it.h_m1 = cam.height - 1; it.w_m2 = cam.width - 2; it.h_m2 = cam.height - 2;
  it.scale = LastValid(rp.scales, rp.n_scales, corr_iteration);  // :1011-1023
  it.fscale = float(it.scale);
  int line_length = kLineSegments * it.scale;
  it.ll_m1 = line_length - 1;
  it.ll_m1_half = float(line_length - 1) * 0.5f;
  it.ll_half_m1 = float(line_length) * 0.5f - 1.0f;
  float sd = LastValid(rp.standard_deviations, rp.n_standard_deviations, corr_iteration);
  it.variance = sd * sd;
}

// Writes the line state into st[field * cap + i]. Returns validity.
__device__ bool RegionLine(const RegionIter& it, const RegionParamsDev& rp, const float4 p0, const float4 p1,
                           const uint8_t* __restrict__ img, unsigned pitch, const float2* __restrict__ lut,
                           float* st, int cap, int i) {
  // CalculateBasicLineData (:1231-1250)
  float x, y, z;
  PoseApply(it.b2c, p0.x, p0.y, p0.z, x, y, z);
  float nu = it.b2c[0] * p0.w + it.b2c[1] * p1.x + it.b2c[2] * p1.y;
  float nv = it.b2c[4] * p0.w + it.b2c[5] * p1.x + it.b2c[6] * p1.y;
  {
    float zz = nu * nu + nv * nv;
    if (zz > 0.0f) { float n = sqrtf(zz); nu /= n; nv /= n; }
  }
  float center_u = x * it.fu / z + it.ppu;
  float center_v = y * it.fv / z + it.ppv;
  st[RF_CBX * cap + i] = p0.x; st[RF_CBY * cap + i] = p0.y; st[RF_CBZ * cap + i] = p0.z;
  st[RF_CU * cap + i] = center_u; st[RF_CV * cap + i] = center_v;
  st[RF_NU * cap + i] = nu; st[RF_NV * cap + i] = nv;
  float continuous_distance = fminf(p1.w, p1.z) * it.fu / (z * it.fscale);
  // IsLineValid (:1252-1291)
  if (continuous_distance < rp.min_continuous_distance) return false;
  if (z <= 0.0f) return false;
  int icu = int(center_u + 0.5f), icv = int(center_v + 0.5f);
  if (icu < 0 || icu > it.w_m1 || icv < 0 || icv > it.h_m1) return false;

  // CalculateSegmentProbabilities (:1433-1573); horizontal / vertical cases folded into major / minor axes
  const bool horizontal = fabsf(nv) < fabsf(nu);
  const float c_major = horizontal ? center_u : center_v;
  const float c_minor = horizontal ? center_v : center_u;
  const float n_major = horizontal ? nu : nv;
  const float n_minor = horizontal ? nv : nu;
  const int major_m1 = horizontal ? it.w_m1 : it.h_m1;
  const int minor_m1 = horizontal ? it.h_m1 : it.w_m1;
  const int minor_m2 = horizontal ? it.h_m2 : it.w_m2;
  const float step = n_minor / n_major;
  int major = int(c_major - it.ll_half_m1);
  const int major_end = major + it.ll_m1;
  float minor_f = c_minor + step * (float(major) - c_major) + 0.5f;
  const float minor_f_end = minor_f + step * float(it.ll_m1);
  if (major < 0 || major_end > major_m1 || int(minor_f) < 0 || int(minor_f) > minor_m1 || int(minor_f_end) < 1 ||
      int(minor_f_end) > minor_m2)
    return false;
  const size_t stride_major = horizontal ? 3u : size_t(pitch);
  const size_t stride_minor = horizontal ? size_t(pitch) : 3u;
  const int bs = rp.bitshift, nb = rp.n_bins;
  float sf[kLineSegments], sb[kLineSegments];
  const uint8_t* pmaj = img + size_t(major) * stride_major;
#pragma unroll
  for (int s = 0; s < kLineSegments; ++s) {
    float pf = 1.0f, pb = 1.0f;
    for (int k = 0; k < it.scale; ++k) {
      const uint8_t* px = pmaj + size_t(int(minor_f)) * stride_minor;
      // ColorHistograms::GetProbabilities index (color_histograms.cpp:97-99), BGR memory order
      int idx = (int(__ldg(px)) >> bs) * nb * nb + (int(__ldg(px + 1)) >> bs) * nb + (int(__ldg(px + 2)) >> bs);
      float2 l = __ldg(lut + idx);  // already normalised per bin (MultiplyPixelColorProbability :1575-1598)
      pf *= l.x;
      pb *= l.y;
      pmaj += stride_major;
      minor_f += step;
    }
    sf[s] = pf;
    sb[s] = pb;
  }
  if (!(n_major > 0.0f)) {  // segments are filled back to front (:1470-1484)
#pragma unroll
    for (int s = 0; s < kLineSegments / 2; ++s) {
      float t = sf[s]; sf[s] = sf[kLineSegments - 1 - s]; sf[kLineSegments - 1 - s] = t;
      t = sb[s]; sb[s] = sb[kLineSegments - 1 - s]; sb[kLineSegments - 1 - s] = t;
    }
  }
  if (it.scale > 1) {  // :1555-1571
#pragma unroll
    for (int s = 0; s < kLineSegments; ++s) {
      if (sf[s] != 0.0f || sb[s] != 0.0f) {
        float sum = sf[s];
        sum += sb[s];
        sf[s] /= sum;
        sb[s] /= sum;
      } else {
        sf[s] = 0.5f;
        sb[s] = 0.5f;
      }
    }
  }
  float ncts = fabsf(n_major) / it.fscale;
  float delta_r = (roundf(c_major - it.ll_m1_half) + it.ll_m1_half - c_major) / n_major;
  st[RF_NCTS * cap + i] = ncts;
  st[RF_DR * cap + i] = delta_r;

  // CalculateDistribution (:1600-1637)
  float dist[kDistributionLength];
  float area = 0.0f;
#pragma unroll
  for (int d = 0; d < kDistributionLength; ++d) {
    float val = 1.0f;
#pragma unroll
    for (int k = 0; k < kFunctionLength; ++k) val *= sf[d + k] * rp.lookup_f[k] + sb[d + k] * rp.lookup_b[k];
    dist[d] = val;
    area += val;
  }
#pragma unroll
  for (int d = 0; d < kDistributionLength; ++d) dist[d] /= area;
  // CalculateDistributionMoments (:1639-1658)
  float mean_from_begin = 0.0f;
#pragma unroll
  for (int d = 0; d < kDistributionLength; ++d) mean_from_begin += float(d) * dist[d];
  float var = 0.0f;
#pragma unroll
  for (int d = 0; d < kDistributionLength; ++d) {
    float dd = float(d) - mean_from_begin;
    var += (dd * dd) * dist[d];
  }
#pragma unroll
  for (int d = 0; d < kDistributionLength; ++d) st[(RF_DIST0 + d) * cap + i] = dist[d];
  st[RF_MEAN * cap + i] = mean_from_begin - (float(kDistributionLength) - 1.0f) / 2.0f;
  st[RF_VAR * cap + i] = fmaxf(var, rp.min_expected_variance);
  return true;
}

// ---------------------------------------------------------------------------------------------
// K3: depth correspondence search (depth_modality.cpp:656-726, 826-884)
// ---------------------------------------------------------------------------------------------
struct DepthIter {
  float b2c[12], c2b[12];
  float fu, fv, ppu, ppv, depth_scale;
  int w_m1, h_m1;
  float considered_distance, standard_deviation;
  int max_n_strides;
};

__device__ __forceinline__ void MakeDepthIter(const DepthParamsDev& dp, const CameraDev& cam, const float* pose,
                                              int corr_iteration, DepthIter& it) {
  PoseMul(cam.w2c, pose, it.b2c);  // depth_modality.cpp:641-646
  PoseInverse(it.b2c, it.c2b);
  it.fu = cam.fu; it.fv = cam.fv; it.ppu = cam.ppu; it.ppv = cam.ppv;
  it.depth_scale = cam.depth_scale;
  it.w_m1 = cam.width - 1; it.h_m1 = cam.height - 1;
  it.considered_distance = LastValid(dp.considered_distances, dp.n_considered_distances, corr_iteration);
  it.max_n_strides = int(it.considered_distance / dp.stride_length + 0.5f);  // :651
  it.standard_deviation = LastValid(dp.standard_deviations, dp.n_standard_deviations, corr_iteration);
}

__device__ bool DepthPoint(const DepthIter& it, const DepthParamsDev& dp, const float4 p0, const float4 p1,
                           const uint8_t* __restrict__ img, unsigned pitch, float* st, int cap, int i) {
  float x, y, z;
  PoseApply(it.b2c, p0.x, p0.y, p0.z, x, y, z);
  st[DF_CBX * cap + i] = p0.x; st[DF_CBY * cap + i] = p0.y; st[DF_CBZ * cap + i] = p0.z;
  st[DF_NX * cap + i] = p0.w; st[DF_NY * cap + i] = p1.x; st[DF_NZ * cap + i] = p1.y;
  float center_u = x * it.fu / z + it.ppu;
  float center_v = y * it.fv / z + it.ppv;
  // IsPointValid (:697-726)
  if (z <= 0.0f) return false;
  int icu = int(center_u + 0.5f), icv = int(center_v + 0.5f);
  if (icu < 0 || icu > it.w_m1 || icv < 0 || icv > it.h_m1) return false;
  // FindCorrespondence (:826-884)
  float considered_distance = it.considered_distance;
  if (dp.use_depth_scaling) considered_distance *= z;
  float meter_to_pixel = it.fu / z;
  float diameter = 2.0f * considered_distance * meter_to_pixel;
  int stride = int(diameter / float(it.max_n_strides) + 1.0f);
  int n_strides = int(diameter / float(stride) + 0.5f);
  int rounded_diameter = n_strides * stride;
  float rounded_radius = 0.5f * float(rounded_diameter);
  int u_min = int(center_u - rounded_radius + 0.5f);
  int v_min = int(center_v - rounded_radius + 0.5f);
  int u_max = u_min + rounded_diameter;
  int v_max = v_min + rounded_diameter;
  u_min = max(u_min, 0);
  v_min = max(v_min, 0);
  u_max = min(u_max, it.w_m1);
  v_max = min(v_max, it.h_m1);
  float min_depth_value = fminf(0.0f, (z - considered_distance) / it.depth_scale);  // sic, :851-852
  float max_depth_value = (z + considered_distance) / it.depth_scale;
  float min_considered_distance_square = considered_distance * considered_distance;
  float best = min_considered_distance_square;
  float bx = 0.0f, by = 0.0f, bz = 0.0f;
  for (int v = v_min; v <= v_max; v += stride) {
    const uint16_t* row = reinterpret_cast<const uint16_t*>(img + size_t(v) * pitch);
    for (int u = u_min; u <= u_max; u += stride) {
      float depth = float(__ldg(row + u));
      if (depth > min_depth_value && depth < max_depth_value) {
        depth *= it.depth_scale;
        float tx = (float(u) - it.ppu) * depth / it.fu;
        float ty = (float(v) - it.ppv) * depth / it.fv;
        float dx = tx - x, dy = ty - y, dz = depth - z;
        float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best) { bx = tx; by = ty; bz = depth; best = d2; }
      }
    }
  }
  if (best == min_considered_distance_square) return false;
  st[DF_YX * cap + i] = bx; st[DF_YY * cap + i] = by; st[DF_YZ * cap + i] = bz;
  return true;
}

// ---------------------------------------------------------------------------------------------
// K4: Optimizer::CalculateOptimization for a rigid body (optimizer.cpp:144-167): Eigen LDLT<Lower>
// with diagonal pivoting restated for n = 6, then Link::UpdatePoses (link.cpp:205-241).
// Runs on one thread; a: 6x6 row-major (lower used), b: rhs, result in theta.
// ---------------------------------------------------------------------------------------------
__device__ void LdltSolve6(float* a, const float* b, float* theta) {
  constexpr int n = 6;
  int trans[n];
  float temp[n];
#define A_(i, j) a[(i) * n + (j)]
  for (int k = 0; k < n; ++k) {
    int biggest = k;
    float big = fabsf(A_(k, k));
    for (int i = k + 1; i < n; ++i) {
      float v = fabsf(A_(i, i));
      if (v > big) { big = v; biggest = i; }
    }
    trans[k] = biggest;
    if (k != biggest) {
      for (int j = 0; j < k; ++j) { float t = A_(k, j); A_(k, j) = A_(biggest, j); A_(biggest, j) = t; }
      for (int i = biggest + 1; i < n; ++i) { float t = A_(i, k); A_(i, k) = A_(i, biggest); A_(i, biggest) = t; }
      { float t = A_(k, k); A_(k, k) = A_(biggest, biggest); A_(biggest, biggest) = t; }
      for (int i = k + 1; i < biggest; ++i) { float t = A_(i, k); A_(i, k) = A_(biggest, i); A_(biggest, i) = t; }
    }
    if (k > 0) {
      for (int j = 0; j < k; ++j) temp[j] = A_(j, j) * A_(k, j);
      float dot = 0.0f;
      for (int j = 0; j < k; ++j) dot += A_(k, j) * temp[j];
      A_(k, k) -= dot;
      for (int i = k + 1; i < n; ++i) {
        float acc = 0.0f;
        for (int j = 0; j < k; ++j) acc += A_(i, j) * temp[j];
        A_(i, k) -= acc;
      }
    }
    float akk = A_(k, k);
    bool pivot_is_valid = fabsf(akk) > 0.0f;
    if (k == 0 && !pivot_is_valid) {
      for (int j = 0; j < n; ++j) trans[j] = j;
      break;
    }
    if (k < n - 1 && pivot_is_valid)
      for (int i = k + 1; i < n; ++i) A_(i, k) /= akk;
  }
  float dst[n];
  for (int i = 0; i < n; ++i) dst[i] = b[i];
  for (int k = 0; k < n; ++k) { float t = dst[k]; dst[k] = dst[trans[k]]; dst[trans[k]] = t; }
  for (int j = 0; j < n; ++j)
    for (int i = j + 1; i < n; ++i) dst[i] -= A_(i, j) * dst[j];
  const float tolerance = 1.0f / 3.402823466e+38f;
  for (int i = 0; i < n; ++i) {
    if (fabsf(A_(i, i)) > tolerance) dst[i] /= A_(i, i);
    else dst[i] = 0.0f;
  }
  for (int j = n - 1; j >= 0; --j)
    for (int i = 0; i < j; ++i) dst[i] -= A_(j, i) * dst[j];
  for (int k = n - 1; k >= 0; --k) { float t = dst[k]; dst[k] = dst[trans[k]]; dst[trans[k]] = t; }
  for (int i = 0; i < n; ++i) theta[i] = dst[i];
#undef A_
}

// Vector2Skewsymmetric(w).exp() in closed form (Rodrigues); the reference uses Eigen's Pade
// approximant (link.cpp:224), the two agree to < 1e-7 for |w| <= 1 (tests/test_oracle_math.py).
__device__ void ExpSkew(const float* w, float* r) {
  float t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  float a, b;
  if (t2 < 1e-8f) {
    a = 1.0f - t2 / 6.0f;
    b = 0.5f - t2 / 24.0f;
  } else {
    float t = sqrtf(t2);
    float sh = sinf(0.5f * t);
    a = sinf(t) / t;
    b = 2.0f * sh * sh / t2;
  }
  float A[9] = {0.0f, -w[2], w[1], w[2], 0.0f, -w[0], -w[1], w[0], 0.0f};
  float A2[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A2[3 * i + j] = A[3 * i + 0] * A[0 + j] + A[3 * i + 1] * A[3 + j] + A[3 * i + 2] * A[6 + j];
#pragma unroll
  for (int k = 0; k < 9; ++k) r[k] = ((k % 4 == 0) ? 1.0f : 0.0f) + a * A[k] + b * A2[k];
}

__device__ void SolveAndUpdate(const BodyDev& body, Shared& sh) {
  // Link::CalculateGradientAndHessian (link.cpp:184-193): region first, then depth.
  // Optimizer: b = J^T g, a(lower) = -J^T H J with J = I6, a.diagonal() += tikhonov (:144-159)
  float* a = sh.a;
  for (int i = 0; i < 6; ++i) {
    sh.b[i] = 0.0f + (0.0f + sh.gh[i] + sh.gh[27 + i]);
    for (int j = 0; j < 6; ++j) {
      float v = 0.0f;
      if (j <= i) {
        float h = 0.0f + sh.gh[6 + Tri(i, j)] + sh.gh[27 + 6 + Tri(i, j)];
        v = 0.0f - h;
      }
      a[6 * i + j] = v;
    }
  }
  for (int i = 0; i < 6; ++i) a[6 * i + i] += (i < 3) ? body.tikhonov_rotation : body.tikhonov_translation;
  float theta[6];
  LdltSolve6(a, sh.b, theta);
  bool nan = false;
  for (int i = 0; i < 6; ++i) nan = nan || isnan(theta[i]);
  if (nan) return;  // optimizer.cpp:165
  float e[9];
  ExpSkew(theta, e);
  float var[12] = {e[0], e[1], e[2], theta[3], e[3], e[4], e[5], theta[4], e[6], e[7], e[8], theta[5]};
  float np[12];
  PoseMul(sh.pose, var, np);  // link2world * [exp | t] (link.cpp:222-238, body2joint = I)
  for (int i = 0; i < 12; ++i) sh.pose[i] = np[i];
}

// ---------------------------------------------------------------------------------------------
// The fused kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads) k_track(TrackArgs args) {
  extern __shared__ float dyn[];
  __shared__ Shared sh;
  const int body_id = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const BodyDev& body = args.bodies[body_id];
  if (!body.set) return;
  const bool has_region = body.has_region, has_depth = body.has_depth;
  const int lcap = args.line_cap, pcap = args.point_cap;
  float* rst = dyn;                     // [RF_COUNT][lcap]
  float* dst = dyn + RF_COUNT * lcap;   // [DF_COUNT][pcap]
  float* g_rst = args.region_state + size_t(body_id) * RF_COUNT * lcap;
  float* g_dst = args.depth_state + size_t(body_id) * DF_COUNT * pcap;
  int* counts = args.counts + 4 * body_id;

  if (tid < 12) sh.pose[tid] = args.poses[12 * body_id + tid];
  if (tid == 0) {
    sh.n_items[0] = (args.phases & PH_LOAD_REGION) ? counts[0] : 0;
    sh.n_items[1] = (args.phases & PH_LOAD_DEPTH) ? counts[1] : 0;
  }
  __syncthreads();
  if (args.phases & PH_LOAD_REGION)
    for (int k = tid; k < RF_COUNT * lcap; k += kBlockThreads) rst[k] = g_rst[k];
  if (args.phases & PH_LOAD_DEPTH)
    for (int k = tid; k < DF_COUNT * pcap; k += kBlockThreads) dst[k] = g_dst[k];
  if (args.phases & PH_LOAD_GH) {
    if (tid < 27) { sh.gh[tid] = args.gh_region[27 * body_id + tid]; sh.gh[27 + tid] = args.gh_depth[27 * body_id + tid]; }
  }
  __syncthreads();

  const CameraDev* ccam = has_region ? &args.color_cams[body.color_camera] : nullptr;
  const CameraDev* dcam = has_depth ? &args.depth_cams[body.depth_camera] : nullptr;
  const ModelDev* rmodel = has_region ? &args.region_models[body.region_model] : nullptr;
  const ModelDev* dmodel = has_depth ? &args.depth_models[body.depth_model] : nullptr;
  const float2* lut = args.lut + size_t(body_id) * args.lut_stride;

  for (int corr = args.corr_begin; corr < args.corr_end; ++corr) {
    // ---------------- CalculateCorrespondences -------------------------------------------------
    if (has_region && (args.phases & PH_REGION_CORR)) {
      RegionIter it;
      MakeRegionIter(body.rp, *ccam, sh.pose, corr, it);
      int view = ClosestView(*rmodel, it.b2c, sh);
      int n_lines = AdaptiveCount(body.rp.n_lines_max, body.rp.use_adaptive_coverage, body.rp.reference_contour_length,
                                  __ldg(rmodel->view_scalars + view), rmodel->max_view_scalar, rmodel->n_points);
      n_lines = min(n_lines, lcap);
      if (tid == 0) { sh.view[0] = view; sh.n_items[0] = n_lines; }
      const float4* pts = rmodel->points + size_t(view) * rmodel->n_points * 2;
      for (int i = tid; i < n_lines; i += kBlockThreads) {
        float4 p0 = __ldg(pts + 2 * i), p1 = __ldg(pts + 2 * i + 1);
        bool ok = RegionLine(it, body.rp, p0, p1, ccam->image, ccam->pitch, lut, rst, lcap, i);
        rst[RF_VALID * lcap + i] = ok ? 1.0f : 0.0f;
      }
    }
    if (has_depth && (args.phases & PH_DEPTH_CORR)) {
      DepthIter it;
      MakeDepthIter(body.dp, *dcam, sh.pose, corr, it);
      int view = ClosestView(*dmodel, it.b2c, sh);
      int n_points = AdaptiveCount(body.dp.n_points_max, body.dp.use_adaptive_coverage, body.dp.reference_surface_area,
                                   __ldg(dmodel->view_scalars + view), dmodel->max_view_scalar, dmodel->n_points);
      n_points = min(n_points, pcap);
      if (tid == 0) { sh.view[1] = view; sh.n_items[1] = n_points; }
      const float4* pts = dmodel->points + size_t(view) * dmodel->n_points * 2;
      for (int i = tid; i < n_points; i += kBlockThreads) {
        float4 p0 = __ldg(pts + 2 * i), p1 = __ldg(pts + 2 * i + 1);
        bool ok = DepthPoint(it, body.dp, p0, p1, dcam->image, dcam->pitch, dst, pcap, i);
        dst[DF_VALID * pcap + i] = ok ? 1.0f : 0.0f;
      }
    }
    __syncthreads();

    // ---------------- n_update x (CalculateGradientAndHessian + CalculateOptimization) ---------
    for (int upd = 0; upd < args.n_update; ++upd) {
      const int opt_iteration = args.opt_base + upd;
      float acc[54];
#pragma unroll
      for (int k = 0; k < 54; ++k) acc[k] = 0.0f;
      if (has_region && (args.phases & PH_REGION_GH)) {
        // K2 region: RegionModality::CalculateGradientAndHessian (region_modality.cpp:485-558)
        RegionIter it;
        MakeRegionIter(body.rp, *ccam, sh.pose, corr, it);
        const int n_lines = sh.n_items[0];
        for (int i = tid; i < n_lines; i += kBlockThreads) {
          if (rst[RF_VALID * lcap + i] == 0.0f) continue;
          float cbx = rst[RF_CBX * lcap + i], cby = rst[RF_CBY * lcap + i], cbz = rst[RF_CBZ * lcap + i];
          float x, y, z;
          PoseApply(it.b2c, cbx, cby, cbz, x, y, z);
          float fu_z = it.fu / z, fv_z = it.fv / z;
          float xfu_z = x * fu_z, yfv_z = y * fv_z;
          float nu = rst[RF_NU * lcap + i], nv = rst[RF_NV * lcap + i];
          float ncts = rst[RF_NCTS * lcap + i];
          float measured_variance = rst[RF_VAR * lcap + i];
          float delta_cs = (nu * (xfu_z + it.ppu - rst[RF_CU * lcap + i]) + nv * (yfv_z + it.ppv - rst[RF_CV * lcap + i]) -
                            rst[RF_DR * lcap + i]) * ncts;
          float dll;
          if (opt_iteration < body.rp.n_global_iterations) {
            dll = (rst[RF_MEAN * lcap + i] - delta_cs) / measured_variance;
          } else {
            int upper = int(delta_cs + (float(kDistributionLength) + 1.0f) / 2.0f);
            int lower = upper - 1;
            if (upper <= 0 || upper >= kDistributionLength) continue;
            dll = (logf(rst[(RF_DIST0 + upper) * lcap + i]) - logf(rst[(RF_DIST0 + lower) * lcap + i])) *
                  body.rp.learning_rate / measured_variance;
          }
          float dc0 = ncts * nu * fu_z;
          float dc1 = ncts * nv * fv_z;
          float dc2 = ncts * (-nu * xfu_z - nv * yfv_z) / z;
          float J[6];
          J[3] = dc0 * it.b2c[0] + dc1 * it.b2c[4] + dc2 * it.b2c[8];
          J[4] = dc0 * it.b2c[1] + dc1 * it.b2c[5] + dc2 * it.b2c[9];
          J[5] = dc0 * it.b2c[2] + dc1 * it.b2c[6] + dc2 * it.b2c[10];
          J[0] = cby * J[5] - cbz * J[4];
          J[1] = cbz * J[3] - cbx * J[5];
          J[2] = cbx * J[4] - cby * J[3];
          float weight = body.rp.min_expected_variance / (ncts * ncts * it.variance);
          float wg = weight * dll;
          float wh = weight / measured_variance;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            acc[r] += wg * J[r];
#pragma unroll
            for (int c = 0; c <= r; ++c) acc[6 + Tri(r, c)] -= (wh * J[r]) * J[c];
          }
        }
      }
      if (has_depth && (args.phases & PH_DEPTH_GH)) {
        // K2 depth: DepthModality::CalculateGradientAndHessian (depth_modality.cpp:333-381)
        DepthIter it;
        MakeDepthIter(body.dp, *dcam, sh.pose, corr, it);
        const int n_points = sh.n_items[1];
        for (int i = tid; i < n_points; i += kBlockThreads) {
          if (dst[DF_VALID * pcap + i] == 0.0f) continue;
          float yc0 = dst[DF_YX * pcap + i], yc1 = dst[DF_YY * pcap + i], yc2 = dst[DF_YZ * pcap + i];
          float yb0, yb1, yb2;
          PoseApply(it.c2b, yc0, yc1, yc2, yb0, yb1, yb2);
          float n0 = dst[DF_NX * pcap + i], n1 = dst[DF_NY * pcap + i], n2 = dst[DF_NZ * pcap + i];
          float epsilon = n0 * (dst[DF_CBX * pcap + i] - yb0) + n1 * (dst[DF_CBY * pcap + i] - yb1) +
                          n2 * (dst[DF_CBZ * pcap + i] - yb2);
          float cx[3] = {yb1 * n2 - yb2 * n1, yb2 * n0 - yb0 * n2, yb0 * n1 - yb1 * n0};
          float weight = 1.0f / (it.standard_deviation * yc2);
          float squared_weight = weight * weight;
          float v[6] = {weight * cx[0], weight * cx[1], weight * cx[2], weight * n0, weight * n1, weight * n2};
          float se = squared_weight * epsilon;
          float nn[3] = {n0, n1, n2};
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            acc[27 + r] -= se * cx[r];
            acc[27 + 3 + r] -= se * nn[r];
          }
          // upper-triangle products v[r] * v[c], r <= c, stored at the mirrored lower index
#pragma unroll
          for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int r = 0; r <= c; ++r) acc[27 + 6 + Tri(c, r)] -= v[r] * v[c];
        }
      }
      // warp-shuffle reduction of the 54 partial sums, then across warps through shared memory
#pragma unroll
      for (int k = 0; k < 54; ++k) {
        float v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sh.red[warp][k] = v;
      }
      __syncthreads();
      if (tid < 54) {
        const bool mine = tid < 27 ? (args.phases & PH_REGION_GH) : (args.phases & PH_DEPTH_GH);
        if (mine || !(args.phases & PH_LOAD_GH)) {
          float v = sh.red[0][tid];
#pragma unroll
          for (int w = 1; w < kWarps; ++w) v += sh.red[w][tid];
          sh.gh[tid] = v;
        }
      }
      __syncthreads();
      if (args.phases & PH_STORE_GH) {
        if (tid < 27 && (args.phases & PH_REGION_GH)) args.gh_region[27 * body_id + tid] = sh.gh[tid];
        if (tid >= 27 && tid < 54 && (args.phases & PH_DEPTH_GH)) args.gh_depth[27 * body_id + tid - 27] = sh.gh[tid];
      }
      if (args.phases & PH_SOLVE) {
        if (tid == 0) SolveAndUpdate(body, sh);
        __syncthreads();
      }
    }
  }

  // ---------------- epilogue ----------------------------------------------------------------------
  if (args.phases & PH_SOLVE)
    if (tid < 12) args.poses[12 * body_id + tid] = sh.pose[tid];
  if (args.phases & PH_STORE_REGION) {
    for (int k = tid; k < RF_COUNT * lcap; k += kBlockThreads) g_rst[k] = rst[k];
    if (tid == 0) { counts[0] = sh.n_items[0]; counts[2] = sh.view[0]; }
  }
  if (args.phases & PH_STORE_DEPTH) {
    for (int k = tid; k < DF_COUNT * pcap; k += kBlockThreads) g_dst[k] = dst[k];
    if (tid == 0) { counts[1] = sh.n_items[1]; counts[3] = sh.view[1]; }
  }
}

// ---------------------------------------------------------------------------------------------
// k_lut: per-bin normalisation (pf, pb) -> (pf/(pf+pb), pb/(pf+pb)) or (0.5, 0.5) if both are zero.
// This is MultiplyPixelColorProbability's per-pixel normalisation (region_modality.cpp:1585-1593)
// hoisted to once per bin; same IEEE divisions, hence bit-identical per-pixel values.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 NormaliseBin(float pf, float pb) {
  if (pf != 0.0f || pb != 0.0f) {
    float sum = pf;
    sum += pb;
    return make_float2(pf / sum, pb / sum);
  }
  return make_float2(0.5f, 0.5f);
}

__global__ void k_lut(const float* hist_f, const float* hist_b, float2* lut, int n, size_t stride, int first_body) {
  const int body = first_body + blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lut[size_t(body) * stride + i] = NormaliseBin(hist_f[size_t(body) * stride + i], hist_b[size_t(body) * stride + i]);
}

// ---------------------------------------------------------------------------------------------
// k_histogram: RegionModality::AddLinePixelColorsToTempHistograms (region_modality.cpp:1025-1155)
// + ColorHistograms::InitializeHistograms / UpdateHistograms (color_histograms.cpp:72-92,174-214).
// One CTA per body. Counts are integer-valued floats (< 2^24), so atomic accumulation order and the
// tree-shaped sum are exact; the blend h = h*(1-lr) + mem*(lr/sum) rounds as the reference does.
// mode 0: StartModality (learning rate 1, also records first_iteration on the host side), 1: CalculateResults.
// ---------------------------------------------------------------------------------------------
struct HistArgs {
  const BodyDev* bodies;
  const float* poses;
  const CameraDev* color_cams;
  const ModelDev* region_models;
  float* hist_f;
  float* hist_b;
  float* mem_f;
  float* mem_b;
  float2* lut;
  size_t stride;
  int mode;
};

__device__ __forceinline__ float sgnf_dev(float v) { return v < 0.0f ? -1.0f : (v > 0.0f ? 1.0f : 0.0f); }

__global__ void __launch_bounds__(kBlockThreads) k_histogram(HistArgs args) {
  __shared__ Shared sh;
  __shared__ float s_sum[2][kWarps];
  const int body_id = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const BodyDev& body = args.bodies[body_id];
  if (!body.set || !body.has_region) return;
  const RegionParamsDev& rp = body.rp;
  const int nbins3 = rp.n_bins * rp.n_bins * rp.n_bins;
  float* mem_f = args.mem_f + size_t(body_id) * args.stride;
  float* mem_b = args.mem_b + size_t(body_id) * args.stride;
  float* hist_f = args.hist_f + size_t(body_id) * args.stride;
  float* hist_b = args.hist_b + size_t(body_id) * args.stride;
  float2* lut = args.lut + size_t(body_id) * args.stride;
  for (int k = tid; k < nbins3; k += kBlockThreads) { mem_f[k] = 0.0f; mem_b[k] = 0.0f; }  // ClearMemory
  if (tid < 12) sh.pose[tid] = args.poses[12 * body_id + tid];
  __syncthreads();
  const CameraDev& cam = args.color_cams[body.color_camera];
  const ModelDev& model = args.region_models[body.region_model];
  RegionIter it;
  MakeRegionIter(rp, cam, sh.pose, 0, it);
  int view = ClosestView(model, it.b2c, sh);
  int n_lines = AdaptiveCount(rp.n_lines_max, rp.use_adaptive_coverage, rp.reference_contour_length,
                              __ldg(model.view_scalars + view), model.max_view_scalar, model.n_points);
  const float4* pts = model.points + size_t(view) * model.n_points * 2;
  const uint8_t* img = cam.image;
  const size_t pitch = cam.pitch;
  const int bs = rp.bitshift, nb = rp.n_bins;
  for (int i = tid; i < n_lines; i += kBlockThreads) {
    float4 p0 = __ldg(pts + 2 * i), p1 = __ldg(pts + 2 * i + 1);
    float x, y, z;
    PoseApply(it.b2c, p0.x, p0.y, p0.z, x, y, z);
    if (z <= 0.0f) continue;
    float center_u = x * it.fu / z + it.ppu;
    float center_v = y * it.fv / z + it.ppv;
    int icu = int(center_u + 0.5f), icv = int(center_v + 0.5f);
    if (icu < 0 || icu > it.w_m1 || icv < 0 || icv > it.h_m1) continue;
    float length_f = rp.max_considered_line_length, length_b = rp.max_considered_line_length;
    float l_f = p1.z * it.fu / z;
    float l_b = p1.w * it.fu / z;
    length_f = fminf(length_f, l_f - 2.0f * rp.unconsidered_line_length);
    length_b = fminf(length_b, l_b - 2.0f * rp.unconsidered_line_length);
    float nu = it.b2c[0] * p0.w + it.b2c[1] * p1.x + it.b2c[2] * p1.y;
    float nv = it.b2c[4] * p0.w + it.b2c[5] * p1.x + it.b2c[6] * p1.y;
    {
      float zz = nu * nu + nv * nv;
      if (zz > 0.0f) { float n = sqrtf(zz); nu /= n; nv /= n; }
    }
    float u_step, v_step;
    int plf, plb;
    float anu = fabsf(nu), anv = fabsf(nv);
    if (anu > anv) {
      u_step = sgnf_dev(nu);
      v_step = nv / anu;
      plf = int(length_f * anu + 0.5f);
      plb = int(length_b * anu + 0.5f);
    } else {
      u_step = nu / anv;
      v_step = sgnf_dev(nv);
      plf = int(length_f * anv + 0.5f);
      plb = int(length_b * anv + 0.5f);
    }
    float u = center_u - nu * rp.unconsidered_line_length + 0.5f;
    float v = center_v - nv * rp.unconsidered_line_length + 0.5f;
    for (int k = 0; k < plf; ++k) {
      int iu = int(u), iv = int(v);
      if (iu < 0 || iu > it.w_m1 || iv < 0 || iv > it.h_m1) break;
      const uint8_t* px = img + size_t(iv) * pitch + 3 * size_t(iu);
      int idx = (int(__ldg(px)) >> bs) * nb * nb + (int(__ldg(px + 1)) >> bs) * nb + (int(__ldg(px + 2)) >> bs);
      atomicAdd(mem_f + idx, 1.0f);
      u -= u_step;
      v -= v_step;
    }
    u = center_u + nu * rp.unconsidered_line_length + 0.5f;
    v = center_v + nv * rp.unconsidered_line_length + 0.5f;
    for (int k = 0; k < plb; ++k) {
      int iu = int(u), iv = int(v);
      if (iu < 0 || iu > it.w_m1 || iv < 0 || iv > it.h_m1) break;
      const uint8_t* px = img + size_t(iv) * pitch + 3 * size_t(iu);
      int idx = (int(__ldg(px)) >> bs) * nb * nb + (int(__ldg(px + 1)) >> bs) * nb + (int(__ldg(px + 2)) >> bs);
      atomicAdd(mem_b + idx, 1.0f);
      u += u_step;
      v += v_step;
    }
  }
  __threadfence();
  __syncthreads();
  // CalculateHistogram (color_histograms.cpp:174-214)
  float sf = 0.0f, sb = 0.0f;
  for (int k = tid; k < nbins3; k += kBlockThreads) { sf += mem_f[k]; sb += mem_b[k]; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sf += __shfl_down_sync(0xffffffffu, sf, off);
    sb += __shfl_down_sync(0xffffffffu, sb, off);
  }
  if (lane == 0) { s_sum[0][warp] = sf; s_sum[1][warp] = sb; }
  __syncthreads();
  float sum_f = 0.0f, sum_b = 0.0f;
  for (int w = 0; w < kWarps; ++w) { sum_f += s_sum[0][w]; sum_b += s_sum[1][w]; }
  const float lr_f = args.mode == 0 ? 1.0f : rp.learning_rate_f;
  const float lr_b = args.mode == 0 ? 1.0f : rp.learning_rate_b;
  const float uniform = 1.0f / float(nbins3);
  const float cf = 1.0f - lr_f, cb = 1.0f - lr_b;
  const float rf = lr_f / sum_f, rb = lr_b / sum_b;
  for (int k = tid; k < nbins3; k += kBlockThreads) {
    float hf = hist_f[k], hb = hist_b[k];
    if (sum_f == 0.0f) {
      if (lr_f == 1.0f) hf = uniform;
    } else if (cf == 0.0f) {
      hf = mem_f[k] * rf;
    } else {
      hf *= cf;
      hf += mem_f[k] * rf;
    }
    if (sum_b == 0.0f) {
      if (lr_b == 1.0f) hb = uniform;
    } else if (cb == 0.0f) {
      hb = mem_b[k] * rb;
    } else {
      hb *= cb;
      hb += mem_b[k] * rb;
    }
    hist_f[k] = hf;
    hist_b[k] = hb;
    lut[k] = NormaliseBin(hf, hb);
  }
}

}  // namespace m3tb
