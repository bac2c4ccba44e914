// frontend.cuh -- frame creation + feature detection.
//
// Stands in for the `frame_creation` / `detection` stages of flame::Flame::update (timing keys
// /root/reference/src/utils.cc:145,148-149; parameters features/detection/{min_grad_mag,win_size},
// /root/reference/src/flame_nodelet.cc:225-233, defaults cfg/flame_nodelet.yaml:68-71).
//   * pyramid level: 2x2 rounded box filter;
//   * gradient magnitude: central differences (0 on the 1 px border);
//   * grid detector: the image is tiled into win x win cells; an unoccupied cell yields its
//     max-gradient pixel (ties: first in row-major order) when the magnitude >= min_grad_mag.
// One warp per cell; the gradient is recomputed from the uint8 image on the fly (the image is
// L2/L1 resident), so no gradient image has to be materialised for detection.
#pragma once

#include "common.cuh"

__global__ void __launch_bounds__(256)
k_pyr_down(int W, int H, const uint8_t* __restrict__ img, uint8_t* __restrict__ out) {
  const int w2 = W / 2, h2 = H / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w2 * h2) return;
  const int x = i % w2, y = i / w2;
  const uint8_t* p = img + (size_t)(2 * y) * W + 2 * x;
  const int s = (int)__ldg(p) + (int)__ldg(p + 1) + (int)__ldg(p + W) + (int)__ldg(p + W + 1);
  out[i] = (uint8_t)((s + 2) >> 2);
}

__device__ __forceinline__ float fb_grad_mag_at(const uint8_t* __restrict__ img, int W, int H, int x,
                                                int y) {
  if (!(x > 0 && y > 0 && x < W - 1 && y < H - 1)) return 0.0f;
  const uint8_t* p = img + (size_t)y * W + x;
  const float gx = 0.5f * ((float)__ldg(p + 1) - (float)__ldg(p - 1));
  const float gy = 0.5f * ((float)__ldg(p + W) - (float)__ldg(p - W));
  return sqrtf(fmaf(gx, gx, gy * gy));
}

__global__ void __launch_bounds__(256)
k_gradient_mag(int W, int H, const uint8_t* __restrict__ img, float* __restrict__ mag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W * H) return;
  mag[i] = fb_grad_mag_at(img, W, H, i % W, i / W);
}

// occupied[cell] = 1 for every cell that holds a valid projected live feature.
__global__ void __launch_bounds__(256)
k_mark_occupied(int W, int H, int win, int N, const float2* __restrict__ u_cur,
                const int32_t* __restrict__ valid, uint8_t* __restrict__ occupied) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N || !valid[f]) return;
  const int cx = W / win, cy = H / win;
  const int i = (int)floorf(u_cur[f].x) / win, j = (int)floorf(u_cur[f].y) / win;
  if (i >= 0 && j >= 0 && i < cx && j < cy) occupied[j * cx + i] = 1;
}

// One warp per cell. det_xy / det_ok are indexed by cell.
__global__ void __launch_bounds__(256)
k_detect_features(int W, int H, int win, int border, float min_grad_mag,
                  const uint8_t* __restrict__ img, const uint8_t* __restrict__ occupied,
                  float2* __restrict__ det_xy, int32_t* __restrict__ det_ok, int y_lo = 0, int y_hi = 1 << 30) {
  const int lane = threadIdx.x & 31;
  const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int cx = W / win, cy = H / win;
  if (cell >= cx * cy) return;
  const int ci = cell % cx, cj = cell / cx;
  float best = -1.0f;
  int bidx = 0x7fffffff;
  if (!(occupied && occupied[cell])) {
    for (int k = lane; k < win * win; k += 32) {  // ascending k per lane: first maximum is kept
      const int x = ci * win + k % win, y = cj * win + k / win;
      if (x < border || y < border || x >= W - border || y >= H - border || y < y_lo || y >= y_hi) continue;  // rows: do_letterbox
      const float g = fb_grad_mag_at(img, W, H, x, y);
      if (g > best) {
        best = g;
        bidx = k;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) {
      best = ob;
      bidx = oi;
    }
  }
  if (lane == 0) {
    const bool ok = bidx != 0x7fffffff && best >= min_grad_mag;
    det_ok[cell] = ok ? 1 : 0;
    det_xy[cell] = ok ? make_float2((float)(ci * win + bidx % win), (float)(cj * win + bidx / win))
                      : make_float2(0.f, 0.f);
  }
}

// Exclusive scan of flags[n] (n <= a few 10k) in one block: out[i] = number of set flags before i,
// total written to *count.  Deterministic compaction order = cell order.
__global__ void __launch_bounds__(1024)
k_scan_flags(int n, const int32_t* __restrict__ flags, int32_t* __restrict__ out, int32_t* count) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const int v = (i < n && flags[i]) ? 1 : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int before = s_carry + (wid ? s_warp[wid - 1] : 0) + incl - v;
    if (i < n) out[i] = before;
    __syncthreads();
    if (tid == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (tid == 0) *count = s_carry;
}
