// Per-pixel geometry of the FlowMap hot path, written once as host+device inline code.
//
// The CUDA kernels in fm_kernels.cu call these from device code; tests/host_emulation
// compiles the very same header with g++ to check the analytic gradients against the
// oracle in a container without a GPU (test infrastructure -- the shipped library has
// no CPU path).  All semantics follow SURVEY.md Appendix A; citations are to
// /root/reference/flowmap/...
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD inline
#endif

namespace fm {

struct K4 {  // normalised intrinsics [[fx 0 cx][0 fy cy][0 0 1]] (intrinsics/common.py:6-20)
  float fx, fy, cx, cy;
};

struct Rt {  // rigid 3x4 [R | t], row-major R
  float r[9];
  float t[3];
};

enum Mapping : int { MAP_HUBER = 0, MAP_L1 = 1, MAP_L2 = 2 };

constexpr float kProjEps = 1e-5f;   // projection.py:52
constexpr float kProjInf = 1e8f;    // projection.py:53

// ---------------------------------------------------------------------------------
// Pixel grid (projection.py:93-113): x = (col + .5) / W, y = (row + .5) / H.
// ---------------------------------------------------------------------------------
FM_HD float pix_x(int c, int W) { return ((float)c + 0.5f) / (float)W; }
FM_HD float pix_y(int r, int H) { return ((float)r + 0.5f) / (float)H; }

// Ray of K^-1 [x y 1]^T with z = 1 (projection.py:84-87).
FM_HD void ray_of(float x, float y, const K4& k, float& rx, float& ry) {
  rx = (x - k.cx) / k.fx;
  ry = (y - k.cy) / k.fy;
}

// ---------------------------------------------------------------------------------
// Bilinear tap set of grid_sample(bilinear, border, align_corners=False) at a
// normalised location (ex, ey) (projection.py:235-241).  Follows ATen's
// unnormalise -> clip -> floor sequence.  Out-of-range taps get weight 0 and a clamped
// (in-bounds) index, so callers may load unconditionally.
// ---------------------------------------------------------------------------------
struct Taps {
  int x0, y0, x1, y1;      // clamped tap coordinates
  float w00, w01, w10, w11;  // weights: w{row}{col}: (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};

FM_HD Taps bilinear_taps(float ex, float ey, int H, int W) {
  float gx = ex * 2.0f - 1.0f, gy = ey * 2.0f - 1.0f;
  float px = ((gx + 1.0f) * (float)W - 1.0f) * 0.5f;
  float py = ((gy + 1.0f) * (float)H - 1.0f) * 0.5f;
  px = fminf((float)(W - 1), fmaxf(px, 0.0f));
  py = fminf((float)(H - 1), fmaxf(py, 0.0f));
  // NaN locations: fmaxf(NaN, 0) = 0, matching ATen's clip of NaN to 0 is not needed for
  // finite flows; finite input is a documented precondition.
  float fx0 = floorf(px), fy0 = floorf(py);
  float tx = px - fx0, ty = py - fy0;
  Taps t;
  t.x0 = (int)fx0;
  t.y0 = (int)fy0;
  t.x1 = t.x0 + 1;
  t.y1 = t.y0 + 1;
  float wx1 = tx, wx0 = 1.0f - tx, wy1 = ty, wy0 = 1.0f - ty;
  if (t.x1 > W - 1) { t.x1 = W - 1; wx1 = 0.0f; }
  if (t.y1 > H - 1) { t.y1 = H - 1; wy1 = 0.0f; }
  t.w00 = wy0 * wx0;
  t.w01 = wy0 * wx1;
  t.w10 = wy1 * wx0;
  t.w11 = wy1 * wx1;
  return t;
}

// Bilinear sample of the xyz image D*ray(K) of a frame (NOT interp(D)*ray(e)): returns
// q = (qx, qy, qz).  `D` points at the frame's (H, W) depth.
template <typename Load>
FM_HD void sample_surface(const Taps& t, int W, int H, const K4& k, Load load, float& qx,
                          float& qy, float& qz) {
  float d00 = load(t.y0 * W + t.x0), d01 = load(t.y0 * W + t.x1);
  float d10 = load(t.y1 * W + t.x0), d11 = load(t.y1 * W + t.x1);
  float a00 = t.w00 * d00, a01 = t.w01 * d01, a10 = t.w10 * d10, a11 = t.w11 * d11;
  float rx0, rx1, ry0, ry1, dummy;
  ray_of(pix_x(t.x0, W), pix_y(t.y0, H), k, rx0, ry0);
  ray_of(pix_x(t.x1, W), pix_y(t.y1, H), k, rx1, ry1);
  (void)dummy;
  qx = (a00 + a10) * rx0 + (a01 + a11) * rx1;
  qy = (a00 + a01) * ry0 + (a10 + a11) * ry1;
  qz = (a00 + a01) + (a10 + a11);
}

// ---------------------------------------------------------------------------------
// Projection u = X / (X_z + eps) on all three components, nan_to_num, then K
// (projection.py:49-58).  Returns uv and keeps what the adjoint needs.
// ---------------------------------------------------------------------------------
struct Proj {
  float u[3];      // after nan_to_num
  float inv;       // 1 / (z + eps)
  bool finite[3];  // gradient passes only through finite components
  float uvx, uvy;
};

FM_HD float nan_to_num1(float v, bool& fin) {
  if (v != v) { fin = false; return 0.0f; }
  if (v > 3.0e38f) { fin = false; return kProjInf; }
  if (v < -3.0e38f) { fin = false; return -kProjInf; }
  fin = true;
  return v;
}

FM_HD Proj project_point(float X, float Y, float Z, const K4& k) {
  Proj p;
  float den = Z + kProjEps;
  p.inv = 1.0f / den;
  p.u[0] = nan_to_num1(X / den, p.finite[0]);
  p.u[1] = nan_to_num1(Y / den, p.finite[1]);
  p.u[2] = nan_to_num1(Z / den, p.finite[2]);
  p.uvx = k.fx * p.u[0] + k.cx * p.u[2];
  p.uvy = k.fy * p.u[1] + k.cy * p.u[2];
  return p;
}

// Adjoint of project_point: given d(uv), returns d(X, Y, Z) and accumulates dK.
FM_HD void project_point_adj(const Proj& p, float X, float Y, float Z, const K4& k, float duvx,
                             float duvy, float& dX, float& dY, float& dZ, float& dfx, float& dfy,
                             float& dcx, float& dcy) {
  dfx += duvx * p.u[0];
  dfy += duvy * p.u[1];
  dcx += duvx * p.u[2];
  dcy += duvy * p.u[2];
  float du0 = p.finite[0] ? k.fx * duvx : 0.0f;
  float du1 = p.finite[1] ? k.fy * duvy : 0.0f;
  float du2 = p.finite[2] ? (k.cx * duvx + k.cy * duvy) : 0.0f;
  dX = du0 * p.inv;
  dY = du1 * p.inv;
  // d/dZ of (X, Y, Z) / (Z + eps)
  dZ = du2 * p.inv - (du0 * X + du1 * Y + du2 * Z) * (p.inv * p.inv);
}

// ---------------------------------------------------------------------------------
// Robust mapping of the aspect-corrected residual (mapping.py:9-43, mapping_huber.py:19-34,
// mapping_l1.py:16-20, mapping_l2.py:16-21).  Returns the loss value and writes
// d(loss)/d(r) for the *uncorrected* residual components (aspect folded in).
// ax = W / sqrt(HW), ay = H / sqrt(HW).
// ---------------------------------------------------------------------------------
FM_HD float robust_map(float rx, float ry, float ax, float ay, int mapping, float delta, float& gx,
                       float& gy) {
  float sx = rx * ax, sy = ry * ay;
  float n2 = sx * sx + sy * sy;
  if (mapping == MAP_L2) {
    gx = sx * ax;
    gy = sy * ay;
    return 0.5f * n2;
  }
  float n = sqrtf(n2);
  float inv_n = n > 0.0f ? 1.0f / n : 0.0f;  // norm has subgradient 0 at the origin
  if (mapping == MAP_L1) {
    gx = sx * inv_n * ax;
    gy = sy * inv_n * ay;
    return n;
  }
  // huber_loss(n, 0, delta) / delta
  if (n <= delta) {
    float id = 1.0f / delta;
    gx = sx * id * ax;
    gy = sy * id * ay;
    return 0.5f * n2 * id;
  }
  gx = sx * inv_n * ax;
  gy = sy * inv_n * ay;
  return n - 0.5f * delta;
}

}  // namespace fm
