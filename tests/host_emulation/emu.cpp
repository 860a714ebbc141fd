// TEST INFRASTRUCTURE -- never loaded by flowmap_b200.
//
// Serial host driver for the per-pixel bodies in flowmap_b200/csrc/fm_pixel.cuh and the
// float64 solver in fm_procrustes.cuh.  It exists so that the analytic gradients can be
// checked against the oracle in the build container, which has no GPU; the CUDA kernels
// instantiate the very same inline functions.  Compiled by tests/test_host_emulation.py
// with g++ into tests/host_emulation/_build/ (git-ignored).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../flowmap_b200/csrc/fm_pixel.cuh"

using namespace fm;

static K4 k4_of(const float* k4, int frame) {
  K4 k; k.fx = k4[frame * 4 + 0]; k.fy = k4[frame * 4 + 1]; k.cx = k4[frame * 4 + 2]; k.cy = k4[frame * 4 + 3];
  return k;
}

static PairGeom geom(const float* depth, const float* k4, int pair, int F, int H, int W, int& a) {
  int bi = pair / (F - 1), i = pair - bi * (F - 1);
  a = bi * F + i;
  PairGeom g;
  g.ka = make_cam(k4_of(k4, a)); g.kb = make_cam(k4_of(k4, a + 1)); g.grid = make_grid(H, W);
  g.z0 = depth[(size_t)(a + 1) * H * W + (size_t)(H / 2) * W + W / 2];
  return g;
}

// Serial twin of k_distribute_tiled / k_distribute_tiled64 (fm_kernels.cu): same tiles (32 x TH),
// window origin and per-tile fixed-point scale (from the upper 32 rows of a tile), (high, low)
// integer cells, flush and float fallback, built from the same inline functions.  stats[0] = tap
// rows added to a window, [1] = tap rows that took the float fallback because they fell outside the
// window, [2] = ... because of the fixed-point range, [3] = high-word adds, [4] = flushed groups
// of four cells, [5] = touched cells outside the image (must stay 0), [6] = largest scaled
// contribution in millionths of 2^29.
template <int TH>
static void bwd_tiled_impl(const float* depth, const float* k4, const float* bflow, const float* weights,
                           const PairState* state, const double* g_rt, float* g_depth, float* g_weights,
                           double* k4acc, long long* stats, int B, int F, int H, int W) {
  constexpr int WINH = TH + 2 * kHalo;
  const int N = H * W, BP = B * (F - 1);
  std::vector<unsigned> lo(kWin * WINH, kFixBias);
  std::vector<int> hi(kWin * WINH, 0);
  long long hi_adds = 0;
  auto add_u = [&lo](int cell, unsigned v) { const unsigned old = lo[cell]; lo[cell] = old + v; return old; };
  auto add_i = [&hi, &hi_adds](int cell, int v) { hi[cell] += v; ++hi_adds; };
  for (int pair = 0; pair < BP; ++pair) {
    int a;
    PairGeom g = geom(depth, k4, pair, F, H, W, a);
    PairAdjoint ad;
    procrustes_adjoint(state[pair], g_rt + (size_t)pair * 12, ad);
    const float* da = depth + (size_t)a * N; const float* db = da + N;
    float* gda = g_depth + (size_t)a * N; float* gdb = gda + N;
    const float* fl = bflow + (size_t)pair * N * 2;
    const float* wt = weights ? weights + (size_t)pair * N : nullptr;
    float bnd_z, bnd_c;
    scatter_bound_consts(g, ad, bnd_z, bnd_c);
    const int tiles_x = W / kTile, tiles_y = (H + TH - 1) / TH;
    for (int tile = 0; tile < tiles_x * tiles_y; ++tile) {
      const int X0 = (tile % tiles_x) * kTile, Y0 = (tile / tiles_x) * TH;
      const int rows = std::min(TH, H - Y0), stat_rows = std::min(kTile, H - Y0);
      float sx = 0.f, sy = 0.f, zmax = 0.f, wmax = 1.f;
      for (int r = Y0; r < Y0 + stat_rows; ++r)
        for (int c = X0; c < X0 + kTile; ++c) {
          const int j = r * W + c;
          sx += fl[2 * j]; sy += fl[2 * j + 1];
          zmax = fmaxf(zmax, fabsf(db[j]));
          wmax = fmaxf(wmax, fabsf(wt ? wt[j] : 1.f));
        }
      int wx0, wy0;
      tile_window_origin_t<WINH>(sx, sy, stat_rows * kTile, X0, Y0, g.grid, wx0, wy0);
      const FixScale fs = fix_scale_for(wmax * fmaf(bnd_z, zmax, bnd_c));
      auto scatter = [&](int y0, int x0, float v0, float v1) {
        // stats[6]: largest |scaled contribution| seen, in millionths of 2^29 (the bound maps to
        // [2^28, 2^29): values above 1e6 would mean the bound does not hold)
        const float big = fmaxf(fabsf(v0 * fs.scale), fabsf(v1 * fs.scale));
        if (big == big) stats[6] = std::max(stats[6], (long long)(big * (1.0e6f / 536870912.0f)));
        if (window_add_t<WINH>(wx0, wy0, fs.scale, y0, x0, v0, v1, add_u, add_i)) {
          ++stats[0];
        } else {
          const int ux = x0 - wx0, uy = y0 - wy0;
          ++stats[((unsigned)ux < (unsigned)(kWin - 1) && (unsigned)uy < (unsigned)WINH) ? 2 : 1];
          gda[y0 * W + x0] += v0;
          if (x0 + 1 < W) gda[y0 * W + x0 + 1] += v1;
        }
      };
      for (int r = Y0; r < Y0 + rows; ++r)
        for (int c = X0; c < X0 + kTile; ++c) {
          const int j = r * W + c;
          float kacc[8] = {0}; float gdj, gwj;
          distribute_point(g, ad, pix_coord(c, g.grid.Wf, g.grid.invW), pix_coord(r, g.grid.Hf, g.grid.invH), db[j],
                           wt ? wt[j] : 1.f, fl[2 * j], fl[2 * j + 1],
                           [da](int o) { return da[o]; }, scatter, gdj, gwj, kacc);
          gdb[j] += gdj;
          if (g_weights) g_weights[(size_t)pair * N + j] += gwj;
          for (int k = 0; k < 8; ++k) k4acc[(size_t)a * 4 + k] += kacc[k];
        }
      for (int i = 0; i < kWin * WINH / 4; ++i) {
        bool clean = true;
        for (int k = 0; k < 4; ++k) clean = clean && lo[i * 4 + k] == kFixBias && hi[i * 4 + k] == 0;
        if (clean) continue;
        const int uy = (i * 4) / kWin, ux = (i * 4) - uy * kWin;
        const int gy = wy0 + uy, gx = wx0 + ux;
        if (gy >= 0 && gy < H && gx >= 0 && gx + 3 < W) {
          for (int k = 0; k < 4; ++k) gda[gy * W + gx + k] += fix_value(lo[i * 4 + k], hi[i * 4 + k]) * fs.inv_scale;
          ++stats[4];
        } else {
          stats[5] += 1;
        }
        for (int k = 0; k < 4; ++k) { lo[i * 4 + k] = kFixBias; hi[i * 4 + k] = 0; }
      }
    }
  }
  stats[3] = hi_adds;
}

extern "C" {

size_t emu_state_bytes() { return sizeof(PairState); }

void emu_procrustes_fwd(const float* depth, const float* k4, const float* bflow, const float* weights,
                        const int64_t* indices, int n_idx, float* rt, PairState* state, int B, int F,
                        int H, int W) {
  const int N = H * W, BP = B * (F - 1);
  for (int pair = 0; pair < BP; ++pair) {
    int a;
    PairGeom g = geom(depth, k4, pair, F, H, W, a);
    const float* da = depth + (size_t)a * N; const float* db = da + N;
    double m[kNumMoments] = {0};
    const int cnt = indices ? n_idx : N;
    for (int t = 0; t < cnt; ++t) {
      const int j = indices ? (int)indices[t] : t;
      float acc[kNumMoments] = {0}; float p[3], q[3]; Taps taps;
      point_pq(g, pix_coord(j % W, g.grid.Wf, g.grid.invW), pix_coord(j / W, g.grid.Hf, g.grid.invH), db[j], bflow[((size_t)pair * N + j) * 2], bflow[((size_t)pair * N + j) * 2 + 1],
               [da](int o) { return da[o]; }, p, q, taps);
      moments_add(acc, weights ? weights[(size_t)pair * N + j] : 1.f, p, q);
      for (int k = 0; k < kNumMoments; ++k) m[k] += acc[k];
    }
    double shift[3] = {0, 0, (double)g.z0};
    procrustes_solve(m, shift, rt + (size_t)pair * 12, state[pair]);
  }
}

// flowacc: (B*F, 40) doubles, zero-filled by the caller.
void emu_flow(const float* depth, const float* k4, const float* rt, const float* ff, const float* fb,
              const float* mf, const float* mb, double mask_sum, int mapping, float delta, float weight,
              float* g_depth, double* flowacc, int B, int F, int H, int W) {
  const int N = H * W;
  if (mask_sum == 0.0) mask_sum = 1.0;
  const float g = (float)((double)weight / mask_sum);
  const RobustCfg rc = make_robust(mapping, delta, H, W);
  const GridDims grid = make_grid(H, W);
  for (int frame = 0; frame < B * F; ++frame) {
    int bi = frame / F, i = frame - bi * F;
    FlowFrame f;
    f.hasF = i < F - 1; f.hasB = i > 0;
    f.kk = make_cam(k4_of(k4, frame)); f.kn = make_cam(k4_of(k4, f.hasF ? frame + 1 : frame)); f.kp = make_cam(k4_of(k4, f.hasB ? frame - 1 : frame));
    int pairF = bi * (F - 1) + i, pairB = pairF - 1;
    auto ld = [&](int pair) { Rt t; for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) t.r[r * 3 + c] = rt[(size_t)pair * 12 + r * 4 + c]; t.t[r] = rt[(size_t)pair * 12 + r * 4 + 3]; } return t; };
    if (f.hasF) f.tf = ld(pairF);
    if (f.hasB) f.tb = ld(pairB);
    for (int j = 0; j < N; ++j) {
      float acc[kFlowVals] = {0};
      const size_t jf = (size_t)(f.hasF ? pairF : 0) * N + j, jb = (size_t)(f.hasB ? pairB : 0) * N + j;
      auto fp = f.hasF && f.hasB ? flow_pixel<true, true> : (f.hasF ? flow_pixel<true, false> : flow_pixel<false, true>);
      g_depth[(size_t)frame * N + j] = fp(
          f, pix_coord(j % W, grid.Wf, grid.invW), pix_coord(j / W, grid.Hf, grid.invH), depth[(size_t)frame * N + j], f.hasF ? ff[jf * 2] : 0.f,
          f.hasF ? ff[jf * 2 + 1] : 0.f, f.hasF ? mf[jf] : 0.f, f.hasB ? fb[jb * 2] : 0.f,
          f.hasB ? fb[jb * 2 + 1] : 0.f, f.hasB ? mb[jb] : 0.f, g, rc, acc);
      for (int k = 0; k < kFlowVals; ++k) flowacc[(size_t)frame * 40 + k] += acc[k];
    }
  }
}

// Lean phase C: writes the STANDARD accumulator layout (after lean_to_standard) into flowacc.
void emu_flow_lean(const float* depth, const float* k4, const float* rt, const float* ff, const float* fb,
                   const float* mf, const float* mb, double mask_sum, int mapping, float delta, float weight,
                   int focal_mode, float* g_depth, double* flowacc, int B, int F, int H, int W) {
  const int N = H * W;
  if (mask_sum == 0.0) mask_sum = 1.0;
  const float g = (float)((double)weight / mask_sum);
  const RobustCfg rc = make_robust(mapping, delta, H, W);
  const GridDims grid = make_grid(H, W);
  for (int frame = 0; frame < B * F; ++frame) {
    int bi = frame / F, i = frame - bi * F;
    const bool hasF = i < F - 1, hasB = i > 0;
    FlowFrameLean f;
    f.kk = make_cam(k4_of(k4, frame)); f.kn = make_cam(k4_of(k4, hasF ? frame + 1 : frame)); f.kp = make_cam(k4_of(k4, hasB ? frame - 1 : frame));
    int pairF = bi * (F - 1) + i, pairB = pairF - 1;
    auto ld = [&](int pair) { Rt t; for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) t.r[r * 3 + c] = rt[(size_t)pair * 12 + r * 4 + c]; t.t[r] = rt[(size_t)pair * 12 + r * 4 + 3]; } return t; };
    Rt tf, tb;
    if (hasF) tf = ld(pairF);
    if (hasB) tb = ld(pairB);
    fill_lean(f, hasF ? &tf : nullptr, hasB ? &tb : nullptr);
    double lean[kFlowLeanVals] = {0};
    // even widths: the packed two-pixel code path (what the vector kernel instantiation runs)
    if (W % 2 == 0) {
      for (int j = 0; j < N; j += 2) {
        F2 acc2[kFlowLeanVals];
        for (int k = 0; k < kFlowLeanVals; ++k) acc2[k] = f2s(0.f);
        const size_t jf = (size_t)(hasF ? pairF : 0) * N + j, jb = (size_t)(hasB ? pairB : 0) * N + j;
        const F2 x = f2(pix_coord(j % W, grid.Wf, grid.invW), pix_coord(j % W + 1, grid.Wf, grid.invW));
        const float y = pix_coord(j / W, grid.Hf, grid.invH);
        const F2 D = f2(depth[(size_t)frame * N + j], depth[(size_t)frame * N + j + 1]);
        const F2 z = f2s(0.f);
        F2 o;
#define FM_CALL2(HF, HB) o = focal_mode ? flow_pixel_lean2<HF, HB, true>(f, x, y, D, HF ? f2(ff[jf * 2], ff[jf * 2 + 2]) : z, HF ? f2(ff[jf * 2 + 1], ff[jf * 2 + 3]) : z, HF ? f2(mf[jf], mf[jf + 1]) : z, HB ? f2(fb[jb * 2], fb[jb * 2 + 2]) : z, HB ? f2(fb[jb * 2 + 1], fb[jb * 2 + 3]) : z, HB ? f2(mb[jb], mb[jb + 1]) : z, g, rc, acc2) \
                                   : flow_pixel_lean2<HF, HB, false>(f, x, y, D, HF ? f2(ff[jf * 2], ff[jf * 2 + 2]) : z, HF ? f2(ff[jf * 2 + 1], ff[jf * 2 + 3]) : z, HF ? f2(mf[jf], mf[jf + 1]) : z, HB ? f2(fb[jb * 2], fb[jb * 2 + 2]) : z, HB ? f2(fb[jb * 2 + 1], fb[jb * 2 + 3]) : z, HB ? f2(mb[jb], mb[jb + 1]) : z, g, rc, acc2)
        if (hasF && hasB) { FM_CALL2(true, true); } else if (hasF) { FM_CALL2(true, false); } else { FM_CALL2(false, true); }
#undef FM_CALL2
        g_depth[(size_t)frame * N + j] = o.x;
        g_depth[(size_t)frame * N + j + 1] = o.y;
        for (int k = 0; k < kFlowLeanVals; ++k) lean[k] += (double)acc2[k].x + (double)acc2[k].y;
      }
    } else
    for (int j = 0; j < N; ++j) {
      float acc[kFlowLeanVals] = {0};
      const size_t jf = (size_t)(hasF ? pairF : 0) * N + j, jb = (size_t)(hasB ? pairB : 0) * N + j;
      const float x = pix_coord(j % W, grid.Wf, grid.invW), y = pix_coord(j / W, grid.Hf, grid.invH);
      const float D = depth[(size_t)frame * N + j];
      float out;
#define FM_CALL(HF, HB) out = focal_mode ? flow_pixel_lean<HF, HB, true>(f, x, y, D, HF ? ff[jf * 2] : 0.f, HF ? ff[jf * 2 + 1] : 0.f, HF ? mf[jf] : 0.f, HB ? fb[jb * 2] : 0.f, HB ? fb[jb * 2 + 1] : 0.f, HB ? mb[jb] : 0.f, g, rc, acc) \
                                 : flow_pixel_lean<HF, HB, false>(f, x, y, D, HF ? ff[jf * 2] : 0.f, HF ? ff[jf * 2 + 1] : 0.f, HF ? mf[jf] : 0.f, HB ? fb[jb * 2] : 0.f, HB ? fb[jb * 2 + 1] : 0.f, HB ? mb[jb] : 0.f, g, rc, acc)
      if (hasF && hasB) { FM_CALL(true, true); } else if (hasF) { FM_CALL(true, false); } else { FM_CALL(false, true); }
#undef FM_CALL
      g_depth[(size_t)frame * N + j] = out;
      for (int k = 0; k < kFlowLeanVals; ++k) lean[k] += acc[k];
    }
    const double s = sqrt((double)H * (double)W);
    const double focal = (double)k4[(size_t)frame * 4] * (double)W / s;
    double out_std[kFlowVals];
    lean_to_standard<double>(lean, hasF ? rt + (size_t)pairF * 12 : nullptr, hasB ? rt + (size_t)pairB * 12 : nullptr, focal, (double)W / s, focal_mode != 0, out_std);
    for (int k = 0; k < kFlowVals; ++k) flowacc[(size_t)frame * 40 + k] = out_std[k];
  }
}

// g_rt: (BP, 12) doubles = dL/d[R|t]; g_depth accumulated into; g_weights written/accumulated;
// k4acc: (B*F, 4) doubles accumulated into.
void emu_procrustes_bwd(const float* depth, const float* k4, const float* bflow, const float* weights,
                        const int64_t* indices, int n_idx, const PairState* state, const double* g_rt,
                        float* g_depth, float* g_weights, double* k4acc, int B, int F, int H, int W) {
  const int N = H * W, BP = B * (F - 1);
  for (int pair = 0; pair < BP; ++pair) {
    int a;
    PairGeom g = geom(depth, k4, pair, F, H, W, a);
    PairAdjoint ad;
    procrustes_adjoint(state[pair], g_rt + (size_t)pair * 12, ad);
    const float* da = depth + (size_t)a * N; const float* db = da + N;
    float* gda = g_depth + (size_t)a * N; float* gdb = gda + N;
    const int cnt = indices ? n_idx : N;
    for (int t = 0; t < cnt; ++t) {
      const int j = indices ? (int)indices[t] : t;
      float kacc[8] = {0}; float gdj, gwj;
      distribute_point(g, ad, pix_coord(j % W, g.grid.Wf, g.grid.invW), pix_coord(j / W, g.grid.Hf, g.grid.invH), db[j], weights ? weights[(size_t)pair * N + j] : 1.f,
                       bflow[((size_t)pair * N + j) * 2], bflow[((size_t)pair * N + j) * 2 + 1],
                       [da](int o) { return da[o]; }, [gda, W](int y0, int x0, float v0, float v1) { gda[y0 * W + x0] += v0; if (x0 + 1 < W) gda[y0 * W + x0 + 1] += v1; }, gdj, gwj, kacc);
      gdb[j] += gdj;
      if (g_weights) g_weights[(size_t)pair * N + j] += gwj;
      for (int k = 0; k < 8; ++k) k4acc[(size_t)a * 4 + k] += kacc[k];
    }
  }
}
void emu_procrustes_bwd_tiled(const float* depth, const float* k4, const float* bflow, const float* weights,
                              const PairState* state, const double* g_rt, float* g_depth,
                              float* g_weights, double* k4acc, long long* stats, int B, int F, int H,
                              int W) {
  bwd_tiled_impl<32>(depth, k4, bflow, weights, state, g_rt, g_depth, g_weights, k4acc, stats, B, F, H, W);
}

void emu_procrustes_bwd_tiled64(const float* depth, const float* k4, const float* bflow, const float* weights,
                                const PairState* state, const double* g_rt, float* g_depth,
                                float* g_weights, double* k4acc, long long* stats, int B, int F, int H,
                                int W) {
  bwd_tiled_impl<64>(depth, k4, bflow, weights, state, g_rt, g_depth, g_weights, k4acc, stats, B, F, H, W);
}

// fix_add / fix_value on ONE cell: adds the n scaled values, returns the (high, low) words and the
// float the flush would produce.
void emu_fix_accumulate(const float* scaled, int n, unsigned* lo_out, int* hi_out, float* value_out) {
  unsigned lo = kFixBias; int hi = 0;
  auto add_u = [&lo](int, unsigned v) { const unsigned old = lo; lo = old + v; return old; };
  auto add_i = [&hi](int, int v) { hi += v; };
  for (int i = 0; i < n; ++i) fix_add(0, scaled[i], add_u, add_i);
  *lo_out = lo; *hi_out = hi; *value_out = fix_value(lo, hi);
}

void emu_fix_scale(float bound, float* scale, float* inv_scale) {
  const FixScale f = fix_scale_for(bound);
  *scale = f.scale; *inv_scale = f.inv_scale;
}
}
