// epipolar.cuh -- per-feature epipolar photometric inverse-depth update (FB_EPI_LANES lanes per feature).
//
// Replaces flame::stereo::inverse_depth_filter::{search,update}, line_stereo and
// InverseDepthMeasModel of the external `flame` core (the `update_idepths` stage,
// /root/reference/src/utils.cc:150; parameters /root/reference/src/flame_nodelet.cc:227-245).
// Algorithm (thresholds documented in DESIGN.md):
//   1. A = K R Kinv, b = K t for T_cmp<-ref; the feature's ray maps to p(xi) = A u + xi b.
//   2. Search interval mu +- k sigma projected to a segment on the epipolar line through u(mu).
//   3. 1-D reference patch (win samples at 1 px along the epipolar direction in ref), gradient gate.
//   4. The comparison image is sampled once along the line (lanes = positions), then a sliding SSD
//      gives one cost per candidate; warp-shuffle arg-min, second-best outside +-radius, parabola.
//   5. Triangulate idepth along the dominant axis; variance = alpha^2 (geo + photo); Gaussian fuse.
// Images are read through the read-only (texture/L1) path with explicit fp32 bilinear blends:
// hardware texture filtering has 9-bit weights and cannot meet the 1e-4 parity bar.
#pragma once

#include "common.cuh"

// ---- relative geometry: one thread per (stream, slot) -----------------------------------------
__device__ __forceinline__ void fb_quat_to_R(const float* q, float* R) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float n = x * x + y * y + z * z + w * w;
  const float s = 2.0f / n;
  const float xx = x * x * s, yy = y * y * s, zz = z * z * s;
  const float xy = x * y * s, xz = x * z * s, yz = y * z * s;
  const float wx = w * x * s, wy = w * y * s, wz = w * z * s;
  R[0] = 1.0f - (yy + zz); R[1] = xy - wz;          R[2] = xz + wy;
  R[3] = xy + wz;          R[4] = 1.0f - (xx + zz); R[5] = yz - wx;
  R[6] = xz - wy;          R[7] = yz + wx;          R[8] = 1.0f - (xx + yy);
}

// poses / cmp_slot may point into PINNED HOST memory (the staging ring): the kernel then reads the few
// hundred bytes over the bus itself and leaves device copies in pose_out / cmp_out for the kernels that
// follow.  A cudaMemcpyAsync for them would queue on the host-to-device copy engine behind the next
// frame's 2.4 MB image upload and stall the compute stream for ~40 us per step (measured).
__device__ __forceinline__ void epi_geometry_body(const float* __restrict__ poses, const float* __restrict__ Ks,
                                                  const int32_t* __restrict__ cmp_slot, int n_slots,
                                                  float* __restrict__ geo, int s0, float* pose_out,
                                                  int32_t* cmp_out, int32_t* counters) {
  const int s = s0 + blockIdx.x, slot = threadIdx.x;  // s0: first stream of the launch
  // the status histogram of every stream this update touches starts from zero (one launch less than a memset)
  if (counters && slot < FB_NUM_COUNTERS && cmp_slot[blockIdx.x + (cmp_out ? 0 : s0)] >= 0) counters[s * FB_NUM_COUNTERS + slot] = 0;
  if (slot >= n_slots) return;
  const int cs = cmp_slot[blockIdx.x + (cmp_out ? 0 : s0)];
  const float* pr = poses + ((size_t)(cmp_out ? blockIdx.x : s) * n_slots + slot) * 7;
  const float* pc = poses + ((size_t)(cmp_out ? blockIdx.x : s) * n_slots + (cs < 0 ? 0 : cs)) * 7;
  if (cmp_out) {  // sources are indexed from the launch's first stream; publish device copies
    if (slot == 0) cmp_out[s] = cs;
    for (int k = 0; k < 7; ++k) pose_out[((size_t)s * n_slots + slot) * 7 + k] = pr[k];
  }
  if (cs < 0) return;
  const float* K = Ks + (size_t)s * 9;
  float* G = geo + ((size_t)s * n_slots + slot) * FB_GEO_STRIDE;
  float Rr[9], Rc[9], R[9], t[3], d[3];
  fb_quat_to_R(pr, Rr);
  fb_quat_to_R(pc, Rc);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      R[3 * r + c] = Rc[0 + r] * Rr[0 + c] + Rc[3 + r] * Rr[3 + c] + Rc[6 + r] * Rr[6 + c];
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = pr[4 + k] - pc[4 + k];
#pragma unroll
  for (int r = 0; r < 3; ++r) t[r] = Rc[0 + r] * d[0] + Rc[3 + r] * d[1] + Rc[6 + r] * d[2];
  const float fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  float M[9];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    M[c] = fx * R[c] + cx * R[6 + c];
    M[3 + c] = fy * R[3 + c] + cy * R[6 + c];
    M[6 + c] = R[6 + c];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float a0 = M[3 * r] / fx, a1 = M[3 * r + 1] / fy;
    G[3 * r] = a0;
    G[3 * r + 1] = a1;
    G[3 * r + 2] = M[3 * r + 2] - (a0 * cx + a1 * cy);
  }
  G[9] = fx * t[0] + cx * t[2];
  G[10] = fy * t[1] + cy * t[2];
  G[11] = t[2];
  float c3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) c3[k] = -(R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2]);
  G[12] = fx * c3[0] + cx * c3[2];
  G[13] = fy * c3[1] + cy * c3[2];
  G[14] = c3[2];
  G[15] = 0.0f;
}
__global__ void k_epi_geometry(const float* __restrict__ poses, const float* __restrict__ Ks,
                               const int32_t* __restrict__ cmp_slot, int n_slots,
                               float* __restrict__ geo, int s0 = 0, float* pose_out = nullptr,
                               int32_t* cmp_out = nullptr, int32_t* counters = nullptr) {
  epi_geometry_body(poses, Ks, cmp_slot, n_slots, geo, s0, pose_out, cmp_out, counters);
}
// The same with the poses and slots travelling IN the launch (kernel parameters, < 4 KB): reading
// them from pinned host memory costs a PCIe round trip per access, and under a saturated host-to-device
// link (the e2e leg uploads 2.4 MB per step) those reads stretched the step's critical path
// (epipolar update k -> assembly k -> epipolar update k+1) by ~20 us.
#define FB_GEO_REC_FLOATS 896
#define FB_GEO_REC_STREAMS 32
struct GeoRecord {
  float poses[FB_GEO_REC_FLOATS];
  int32_t cmp[FB_GEO_REC_STREAMS];
};
__global__ void k_epi_geometry_rec(const __grid_constant__ GeoRecord rec, const float* __restrict__ Ks, int n_slots,
                                   float* __restrict__ geo, int s0, float* pose_out, int32_t* cmp_out, int32_t* counters) {
  epi_geometry_body(rec.poses, Ks, rec.cmp, n_slots, geo, s0, pose_out, cmp_out, counters);
}

// ---- image sampling ---------------------------------------------------------------------------
__device__ __forceinline__ bool fb_inside(float x, float y, int W, int H) {
  return x >= 0.0f && y >= 0.0f && x < (float)(W - 1) && y < (float)(H - 1);
}

__device__ __forceinline__ float fb_bilin(const uint8_t* __restrict__ img, int W, float x, float y) {
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;
  const float fx = x - xf, fy = y - yf;
  const uint8_t* p = img + (size_t)y0 * (size_t)W + (size_t)x0;
  const float i00 = (float)__ldg(p), i10 = (float)__ldg(p + 1);
  const float i01 = (float)__ldg(p + W), i11 = (float)__ldg(p + W + 1);
  const float a = fmaf(fx, i10 - i00, i00);
  const float b = fmaf(fx, i11 - i01, i01);
  return fmaf(fy, b - a, a);
}

__device__ __forceinline__ float fb_idepth_at(float x, float y, bool use_x, float P0x, float P0y,
                                              float P0z, float bx, float by, float bz) {
  if (use_x) return fmaf(x, P0z, -P0x) / fmaf(-x, bz, bx);
  return fmaf(y, P0z, -P0y) / fmaf(-y, bz, by);
}

// Lanes cooperating on one feature.  The scalar set-up (geometry, interval, direction) is uniform
// over the group, so a full warp per feature spends ~3/4 of its issue slots on redundant work;
// 8 lanes keep the sampling / SSD loops parallel while 4 features share a warp.
#define FB_EPI_LANES 8
#define FB_EPI_GROUPS (32 / FB_EPI_LANES)

struct EpiArgs {
  const uint8_t* imgs;
  const float* geo;
  const int32_t* cmp_slot;
  const float2* u_ref;
  const int32_t* ref_slot;
  float* mu;
  float* var;
  int32_t* dropouts;
  int32_t* alive;
  int32_t* status;
  float2* u_cmp;
  const int32_t* nF;
  int32_t* counters;
  int W, H, n_slots, maxF;
  int s0;  // first stream of the launch (blockIdx.y counts from it)
  const uint8_t* cmp_frames;  // non-NULL: the comparison frames live here, stream-major and contiguous
                              // (the landing buffer of a single host-to-device transfer), not in their slot
  fb_epi_params p;
};

// Warp-uniform search; returns the status, updates mu/var (all lanes hold identical scalars).
__device__ int fb_epi_update_one(const EpiArgs& a, const uint8_t* __restrict__ iref,
                                 const uint8_t* __restrict__ icmp, const float* __restrict__ G,
                                 float ux, float uy, float& mu_io, float& var_io, float2& ucmp,
                                 float* s_line, float* s_cost, float* s_ref, int lane,
                                 unsigned gmask) {
  const fb_epi_params& p = a.p;
  const int W = a.W, H = a.H;
  const int win = p.win_size, h = win / 2;
  const float bx = G[9], by = G[10], bz = G[11];
  const float P0x = fmaf(G[0], ux, fmaf(G[1], uy, G[2]));
  const float P0y = fmaf(G[3], ux, fmaf(G[4], uy, G[5]));
  const float P0z = fmaf(G[6], ux, fmaf(G[7], uy, G[8]));
  const float m = mu_io, v = var_io;
  const float sigma = sqrtf(v);
  float xi_lo = fmaxf(fmaf(-p.search_sigma, sigma, m), p.idepth_min);
  float xi_hi = fminf(fmaf(p.search_sigma, sigma, m), p.idepth_max);
  xi_hi = fmaxf(xi_hi, xi_lo);
  const float pz_mu = fmaf(m, bz, P0z);
  const float pz_lo = fmaf(xi_lo, bz, P0z);
  const float pz_hi = fmaf(xi_hi, bz, P0z);
  if (!(pz_mu > 1e-6f) || !(pz_lo > 1e-6f) || !(pz_hi > 1e-6f)) return FB_FAIL_OUT_OF_IMAGE;
  const float umx = fmaf(m, bx, P0x) / pz_mu, umy = fmaf(m, by, P0y) / pz_mu;
  const float dxv = fmaf(-umx, bz, bx), dyv = fmaf(-umy, bz, by);
  const float dn = sqrtf(fmaf(dxv, dxv, dyv * dyv));
  const float gpar = dn / pz_mu;
  if (!(gpar >= p.min_parallax)) return FB_NO_PARALLAX;
  const float lx = dxv / dn, ly = dyv / dn;
  const float ulx = fmaf(xi_lo, bx, P0x) / pz_lo, uly = fmaf(xi_lo, by, P0y) / pz_lo;
  const float uhx = fmaf(xi_hi, bx, P0x) / pz_hi, uhy = fmaf(xi_hi, by, P0y) / pz_hi;
  float s_lo = fmaf(ulx - umx, lx, (uly - umy) * ly);
  float s_hi = fmaf(uhx - umx, lx, (uhy - umy) * ly);
  const float half = 0.5f * (float)(p.max_search_px - 4);
  s_lo = fminf(fmaxf(s_lo, -half), 0.0f);
  s_hi = fmaxf(fminf(s_hi, half), 0.0f);
  const float s0 = floorf(s_lo) - 1.0f;
  int n_steps = (int)(ceilf(s_hi) - s0) + 2;
  if (n_steps > p.max_search_px) n_steps = p.max_search_px;

  // reference patch along the epipolar direction in ref
  float lrx = fmaf(ux, G[14], -G[12]), lry = fmaf(uy, G[14], -G[13]);
  const float lrn = sqrtf(fmaf(lrx, lrx, lry * lry));
  if (!(lrn > 1e-12f)) return FB_NO_PARALLAX;
  lrx = lrx / lrn;
  lry = lry / lrn;
  bool ref_in = true;
  for (int k = lane; k < win; k += FB_EPI_LANES) {
    const float kk = (float)(k - h);
    const float x = fmaf(kk, lrx, ux), y = fmaf(kk, lry, uy);
    const bool in = fb_inside(x, y, W, H);
    ref_in = ref_in && in;
    s_ref[k] = in ? fb_bilin(iref, W, x, y) : 0.0f;
  }
  if (!__all_sync(gmask, ref_in)) return FB_FAIL_OUT_OF_IMAGE;
  __syncwarp(gmask);
  float grad2 = 0.0f;
  for (int k = 0; k + 1 < win; ++k) {
    const float d = s_ref[k + 1] - s_ref[k];
    grad2 = fmaf(d, d, grad2);
  }
  grad2 = grad2 / (float)(win - 1);
  if (grad2 < p.min_grad_mag * p.min_grad_mag) return FB_FAIL_REF_PATCH_GRADIENT;
  if (!fb_inside(ux - 1.0f, uy - 1.0f, W, H) || !fb_inside(ux + 1.0f, uy + 1.0f, W, H))
    return FB_FAIL_OUT_OF_IMAGE;
  const float gx = 0.5f * (fb_bilin(iref, W, ux + 1.0f, uy) - fb_bilin(iref, W, ux - 1.0f, uy));
  const float gy = 0.5f * (fb_bilin(iref, W, ux, uy + 1.0f) - fb_bilin(iref, W, ux, uy - 1.0f));

  // sample the comparison image once along the line: lanes = positions (NaN marks "outside")
  const int n_samp = n_steps + 2 * h;
  for (int mI = lane; mI < n_samp; mI += FB_EPI_LANES) {
    const float s = s0 + (float)(mI - h);
    const float x = fmaf(s, lx, umx), y = fmaf(s, ly, umy);
    s_line[mI] = fb_inside(x, y, W, H) ? fb_bilin(icmp, W, x, y) : __int_as_float(0x7fc00000);
  }
  __syncwarp(gmask);
  // sliding SSD: lanes = candidates; per-lane running best keeps the smallest n on ties
  const float INF = __int_as_float(0x7f800000);
  float best = INF;
  int nbest = 0x7fffffff;
  for (int n = lane; n < n_steps; n += FB_EPI_LANES) {
    float c = 0.0f;
    for (int k = 0; k < win; ++k) {
      const float d = s_line[n + k] - s_ref[k];
      c = fmaf(d, d, c);
    }
    s_cost[n] = c;  // NaN when any sample was outside the image
    if (c == c && c < best) {
      best = c;
      nbest = n;
    }
  }
#pragma unroll
  for (int o = FB_EPI_LANES / 2; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(gmask, best, o);
    const int on = __shfl_xor_sync(gmask, nbest, o);
    if (ob < best || (ob == best && on < nbest)) {
      best = ob;
      nbest = on;
    }
  }
  if (nbest == 0x7fffffff) return FB_FAIL_OUT_OF_IMAGE;
  __syncwarp(gmask);
  float second = INF;
  for (int n = lane; n < n_steps; n += FB_EPI_LANES) {
    const float c = s_cost[n];
    int dn2 = n - nbest;
    if (dn2 < 0) dn2 = -dn2;
    if (c == c && dn2 > p.ambiguity_radius && c < second) second = c;
  }
#pragma unroll
  for (int o = FB_EPI_LANES / 2; o > 0; o >>= 1) second = fminf(second, __shfl_xor_sync(gmask, second, o));
  if (best > p.max_cost * (float)win) return FB_FAIL_MAX_COST;
  const float floor_c = p.pixel_noise_var * (float)win;
  if (second < INF && second < p.ambiguity_ratio * fmaxf(best, floor_c))
    return FB_FAIL_AMBIGUOUS_MATCH;
  float delta = 0.0f;
  if (nbest > 0 && nbest + 1 < n_steps) {
    const float cm = s_cost[nbest - 1], c0 = s_cost[nbest], cp = s_cost[nbest + 1];
    if (cm == cm && cp == cp) {
      const float den = (cm - 2.0f * c0) + cp;
      if (den > 1e-12f) {
        delta = (0.5f * (cm - cp)) / den;
        delta = fminf(fmaxf(delta, -0.5f), 0.5f);
      }
    }
  }
  const float sstar = (s0 + (float)nbest) + delta;
  const float ucx = fmaf(sstar, lx, umx), ucy = fmaf(sstar, ly, umy);
  ucmp = make_float2(ucx, ucy);
  const bool use_x = fabsf(lx) >= fabsf(ly);
  const float xi_m = fb_idepth_at(ucx, ucy, use_x, P0x, P0y, P0z, bx, by, bz);
  const float sp = sstar + 0.5f, sm = sstar - 0.5f;
  const float xi_p =
      fb_idepth_at(fmaf(sp, lx, umx), fmaf(sp, ly, umy), use_x, P0x, P0y, P0z, bx, by, bz);
  const float xi_n =
      fb_idepth_at(fmaf(sm, lx, umx), fmaf(sm, ly, umy), use_x, P0x, P0y, P0z, bx, by, bz);
  const float alpha = xi_p - xi_n;
  const float g2 = fmaf(gx, gx, gy * gy);
  const float gl = fmaf(gx, lrx, gy * lry);
  const float var_geo = p.epipolar_line_var * ((g2 + 1e-6f) / fmaf(gl, gl, 1e-6f));
  const float var_photo = (2.0f * p.pixel_noise_var) / (grad2 + 1e-6f);
  const float var_m = (alpha * alpha) * (var_geo + var_photo);
  if (!(var_m <= p.meas_var_max) || !(xi_m == xi_m)) return FB_FAIL_MAX_VAR;
  const float den = v + var_m;
  mu_io = fmaf(var_m, m, v * xi_m) / den;
  var_io = (v * var_m) / den;
  return FB_SUCCESS;
}

// grid = (ceil(maxF / (warps_per_block * FB_EPI_GROUPS)), S);
// dynamic smem = warps * FB_EPI_GROUPS * (2*max_search + 2*FB_MAX_WIN + 2) floats
__global__ void __launch_bounds__(256) k_epipolar_search(EpiArgs a) {
  extern __shared__ float smem[];
  const int s = a.s0 + blockIdx.y;
  const int wlane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int group = wlane / FB_EPI_LANES, lane = wlane % FB_EPI_LANES;
  const unsigned gmask = (FB_EPI_LANES == 32 ? 0xffffffffu : ((1u << FB_EPI_LANES) - 1u)) << (group * FB_EPI_LANES);
  const int f = (blockIdx.x * wpb + wib) * FB_EPI_GROUPS + group;
  const int cs = a.cmp_slot[s];
  if (cs < 0 || f >= a.nF[s]) return;
  const int per_group = 2 * a.p.max_search_px + 2 * FB_MAX_WIN + 2;
  float* s_line = smem + (size_t)(wib * FB_EPI_GROUPS + group) * per_group;
  float* s_cost = s_line + a.p.max_search_px + FB_MAX_WIN + 1;
  float* s_ref = s_cost + a.p.max_search_px;
  const size_t fb = (size_t)s * a.maxF + f;
  const float qnan = __int_as_float(0x7fc00000);
  float2 ucmp = make_float2(qnan, qnan);
  if (!a.alive[fb]) {
    if (lane == 0) {
      a.status[fb] = FB_SKIPPED;
      a.u_cmp[fb] = ucmp;
    }
    return;
  }
  const int r = a.ref_slot[fb];
  int st;
  float mu = a.mu[fb], var = a.var[fb];
  if (r == cs) {
    st = FB_NO_PARALLAX;
  } else {
    const size_t fsz = (size_t)a.W * a.H;
    const uint8_t* iref = a.imgs + ((size_t)s * a.n_slots + r) * fsz;
    const uint8_t* icmp = a.cmp_frames ? a.cmp_frames + (size_t)s * fsz : a.imgs + ((size_t)s * a.n_slots + cs) * fsz;
    const float* G = a.geo + ((size_t)s * a.n_slots + r) * FB_GEO_STRIDE;
    const float2 u = a.u_ref[fb];
    st = fb_epi_update_one(a, iref, icmp, G, u.x, u.y, mu, var, ucmp, s_line, s_cost, s_ref, lane, gmask);
  }
  if (lane == 0) {
    if (st == FB_SUCCESS) {
      a.mu[fb] = mu;
      a.var[fb] = var;
      a.dropouts[fb] = 0;
    } else if (st != FB_NO_PARALLAX) {
      ucmp = make_float2(qnan, qnan);
      const int d = a.dropouts[fb] + 1;
      a.dropouts[fb] = d;
      if (d > a.p.max_dropouts) {
        a.alive[fb] = 0;
        st = FB_FAIL_MAX_DROPOUTS;
      }
    }
    a.status[fb] = st;
    a.u_cmp[fb] = ucmp;
    atomicAdd(&a.counters[s * FB_NUM_COUNTERS + st], 1);
  }
}

// project_features stage: one thread per feature.
__global__ void __launch_bounds__(256)
k_project_features(const float* __restrict__ geo, int n_slots, int s, int N, int W, int H,
                   const float2* __restrict__ u_ref, const int32_t* __restrict__ ref_slot,
                   const float* __restrict__ mu, const float* __restrict__ var,
                   const int32_t* __restrict__ alive, float2* u_cur, float* mu_cur, float* var_cur,
                   int32_t* valid) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N) return;
  const float qnan = __int_as_float(0x7fc00000);
  valid[f] = 0;
  u_cur[f] = make_float2(qnan, qnan);
  mu_cur[f] = qnan;
  var_cur[f] = qnan;
  if (!alive[f]) return;
  const float* g = geo + ((size_t)s * n_slots + ref_slot[f]) * FB_GEO_STRIDE;
  const float ux = u_ref[f].x, uy = u_ref[f].y, m = mu[f];
  const float px = fmaf(m, g[9], fmaf(g[0], ux, fmaf(g[1], uy, g[2])));
  const float py = fmaf(m, g[10], fmaf(g[3], ux, fmaf(g[4], uy, g[5])));
  const float pz = fmaf(m, g[11], fmaf(g[6], ux, fmaf(g[7], uy, g[8])));
  if (!(pz > 1e-6f)) return;
  const float x = px / pz, y = py / pz;
  if (!(x >= 0.0f && y >= 0.0f && x <= (float)(W - 1) && y <= (float)(H - 1))) return;
  const float r = 1.0f / pz;
  const float r2 = r * r;
  u_cur[f] = make_float2(x, y);
  mu_cur[f] = m * r;
  var_cur[f] = var[f] * (r2 * r2);
  valid[f] = 1;
}

// Fresh filter state for every feature of the selected streams (ref_slot[s] < 0: leave untouched).
__global__ void __launch_bounds__(256)
k_features_reinit(const int32_t* __restrict__ new_ref, const int32_t* __restrict__ nF, int maxF,
                  float mu0, float var0, float* mu, float* var, int32_t* dropouts, int32_t* alive,
                  int32_t* ref_slot) {
  const int s = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = new_ref[s];
  if (r < 0 || f >= nF[s]) return;
  const size_t fb = (size_t)s * maxF + f;
  mu[fb] = mu0;
  var[fb] = var0;
  dropouts[fb] = 0;
  alive[fb] = 1;
  ref_slot[fb] = r;
}
// The same with the slots in the launch (see k_epi_geometry_rec).
struct SlotRecord { int32_t v[FB_GEO_REC_STREAMS]; };
__global__ void __launch_bounds__(256)
k_features_reinit_rec(const __grid_constant__ SlotRecord new_ref, const int32_t* __restrict__ nF, int maxF,
                      float mu0, float var0, float* mu, float* var, int32_t* dropouts, int32_t* alive,
                      int32_t* ref_slot) {
  const int s = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = new_ref.v[s];
  if (r < 0 || f >= nF[s]) return;
  const size_t fb = (size_t)s * maxF + f;
  mu[fb] = mu0;
  var[fb] = var0;
  dropouts[fb] = 0;
  alive[fb] = 1;
  ref_slot[fb] = r;
}
