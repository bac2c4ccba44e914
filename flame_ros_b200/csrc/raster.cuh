// raster.cuh -- mesh -> dense inverse-depth map (the `interpolate` stage, /root/reference/src/utils.cc:146)
// and the display filters behind flame::Flame::getFilteredInverseDepthMap
// (/root/reference/src/flame_nodelet.cc:182-206, 682-683).
//
// Pixel ownership is "the smallest-index valid triangle whose inclusive edge test passes", made
// deterministic with an integer atomicMin claim pass followed by a per-pixel shading pass that
// recomputes the barycentric weights with the same expressions.
#pragma once

#include "common.cuh"

#define FB_OWNER_NONE 0x7f7f7f7f

__device__ __forceinline__ float fb_edge_fn(float ax, float ay, float bx, float by, float px,
                                            float py) {
  return fmaf(bx - ax, py - ay, -((by - ay) * (px - ax)));
}

__global__ void __launch_bounds__(256)
k_tri_validity(int W, const float* __restrict__ K, const float2* __restrict__ vtx,
               const float* __restrict__ idepth, int T, const int32_t* __restrict__ tri,
               fb_tri_filter_params fp, float cos_thresh, int use_filter,
               uint8_t* __restrict__ valid, const int32_t* __restrict__ Tdev = nullptr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (Tdev) T = *Tdev;  // device-built meshes: the count never visits the host
  if (t >= T) return;
  const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
  const float d[3] = {idepth[a], idepth[b], idepth[c]};
  if (!use_filter) {
    valid[t] = 1;
    return;
  }
  const float px[3] = {vtx[a].x, vtx[b].x, vtx[c].x};
  const float py[3] = {vtx[a].y, vtx[b].y, vtx[c].y};
  const float fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  const float len_max = fp.edge_length_thresh * (float)W;
  const float len2_max = len_max * len_max;
  int ok = 1;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (!(d[k] > 0.0f)) ok = 0;
  if (ok && fp.do_idepth) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (d[k] < fp.min_triangle_idepth) ok = 0;
  }
  if (ok && fp.do_edge_length) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int k2 = (k + 1) % 3;
      const float ex = px[k] - px[k2], ey = py[k] - py[k2];
      if (fmaf(ex, ex, ey * ey) > len2_max) ok = 0;
    }
  }
  if (ok && fp.do_oblique) {
    const float dmax = fmaxf(d[0], fmaxf(d[1], d[2]));
    const float dmin = fminf(d[0], fminf(d[1], d[2]));
    const float diff = dmax - dmin;
    if (diff > fmaxf(fp.oblique_idepth_diff_factor * dmax, fp.oblique_idepth_diff_abs)) ok = 0;
    float P[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float zk = 1.0f / d[k];
      P[k][0] = ((px[k] - cx) / fx) * zk;
      P[k][1] = ((py[k] - cy) / fy) * zk;
      P[k][2] = zk;
    }
    float e1[3], e2[3], n[3], ctr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      e1[k] = P[1][k] - P[0][k];
      e2[k] = P[2][k] - P[0][k];
      ctr[k] = (P[0][k] + P[1][k]) + P[2][k];
    }
    n[0] = fmaf(e1[1], e2[2], -(e1[2] * e2[1]));
    n[1] = fmaf(e1[2], e2[0], -(e1[0] * e2[2]));
    n[2] = fmaf(e1[0], e2[1], -(e1[1] * e2[0]));
    const float nn = fmaf(n[0], n[0], fmaf(n[1], n[1], n[2] * n[2]));
    const float cc = fmaf(ctr[0], ctr[0], fmaf(ctr[1], ctr[1], ctr[2] * ctr[2]));
    const float dot = fmaf(n[0], ctr[0], fmaf(n[1], ctr[1], n[2] * ctr[2]));
    if (cos_thresh > 0.0f && dot * dot < (cos_thresh * cos_thresh) * (nn * cc)) ok = 0;
  }
  valid[t] = (uint8_t)ok;
}

// One warp per triangle: lanes stride over the bounding box and claim covered pixels.
__global__ void __launch_bounds__(256)
k_raster_claim(int W, int H, const float2* __restrict__ vtx, int T,
               const int32_t* __restrict__ tri, const uint8_t* __restrict__ valid,
               int32_t* __restrict__ owner, const int32_t* __restrict__ Tdev = nullptr,
               const uint8_t* __restrict__ valid2 = nullptr, int32_t* __restrict__ owner2 = nullptr,
               int32_t* __restrict__ covered_zero = nullptr) {
  // the shading pass that follows counts covered pixels into this word: zeroed here instead of by a memset node
  if (covered_zero && blockIdx.x == 0 && threadIdx.x == 0) *covered_zero = 0;
  // owner2 != NULL: a second ownership map over the triangles that pass `valid2` is claimed in the same
  // pass (fb_update renders the filtered map of the next getter call beside the unfiltered one)
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (Tdev) T = *Tdev;
  if (t >= T || (valid && !valid[t])) return;  // valid == NULL: every triangle (unfiltered map)
  const bool second = owner2 && valid2[t];
  const float2 A = vtx[tri[3 * t]], B = vtx[tri[3 * t + 1]], C = vtx[tri[3 * t + 2]];
  const float area = fb_edge_fn(A.x, A.y, B.x, B.y, C.x, C.y);
  if (area == 0.0f || !(area == area)) return;
  const float xmin = fminf(A.x, fminf(B.x, C.x)), xmax = fmaxf(A.x, fmaxf(B.x, C.x));
  const float ymin = fminf(A.y, fminf(B.y, C.y)), ymax = fmaxf(A.y, fmaxf(B.y, C.y));
  const int x0 = (int)ceilf(fmaxf(xmin, 0.0f)), x1 = (int)floorf(fminf(xmax, (float)(W - 1)));
  const int y0 = (int)ceilf(fmaxf(ymin, 0.0f)), y1 = (int)floorf(fminf(ymax, (float)(H - 1)));
  if (x1 < x0 || y1 < y0) return;
  const float inv = 1.0f / area;
  const int bw = x1 - x0 + 1, n = bw * (y1 - y0 + 1);
  for (int k = lane; k < n; k += 32) {
    const int x = x0 + k % bw, y = y0 + k / bw;
    const float px = (float)x, py = (float)y;
    const float w0 = fb_edge_fn(B.x, B.y, C.x, C.y, px, py) * inv;
    const float w1 = fb_edge_fn(C.x, C.y, A.x, A.y, px, py) * inv;
    const float w2 = fb_edge_fn(A.x, A.y, B.x, B.y, px, py) * inv;
    if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) {
      atomicMin(&owner[y * W + x], t);
      if (second) atomicMin(&owner2[y * W + x], t);
    }
  }
}

__device__ __forceinline__ float fb_shade_one(int W, const float2* __restrict__ vtx, const float* __restrict__ idepth,
                                              const int32_t* __restrict__ tri, int t, int i) {
  const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
  const float2 A = vtx[a], B = vtx[b], C = vtx[c];
  const float inv = 1.0f / fb_edge_fn(A.x, A.y, B.x, B.y, C.x, C.y);
  const float px = (float)(i % W), py = (float)(i / W);
  const float w0 = fb_edge_fn(B.x, B.y, C.x, C.y, px, py) * inv;
  const float w1 = fb_edge_fn(C.x, C.y, A.x, A.y, px, py) * inv;
  const float w2 = fb_edge_fn(A.x, A.y, B.x, B.y, px, py) * inv;
  return fmaf(w0, idepth[a], fmaf(w1, idepth[b], w2 * idepth[c]));
}

__global__ void __launch_bounds__(256)
k_raster_shade(int W, int H, const float2* __restrict__ vtx, const float* __restrict__ idepth,
               const int32_t* __restrict__ tri, int32_t* __restrict__ owner,
               float* __restrict__ map, const int32_t* __restrict__ Tdev = nullptr,
               int32_t* __restrict__ covered = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (Tdev && *Tdev == 0) return;  // no mesh this frame: the previous map stays
  if (i >= W * H) return;
  const int t = owner[i];
  owner[i] = FB_OWNER_NONE;  // the ownership map is left clean for the next claim pass (no memset per frame)
  float out = __int_as_float(0x7fc00000);
  if (t != FB_OWNER_NONE) out = fb_shade_one(W, vtx, idepth, tri, t, i);
  map[i] = out;
  if (covered) {  // `coverage` stat (/root/reference/src/utils.cc:122): pixels with a depth
    const unsigned m = __ballot_sync(__activemask(), out == out);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(covered, __popc(m));
  }
}
