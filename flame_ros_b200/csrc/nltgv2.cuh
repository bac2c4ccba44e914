// nltgv2.cuh -- NLTGV2-L1 Chambolle-Pock iteration on a Delaunay graph, streaming variant.
//
// Replaces flame::optimizers::nltgv2_l1_graph_regularizer::step (external `flame` core; parameters
// forwarded at /root/reference/src/flame_nodelet.cc:256-259).  Iteration as in SURVEY.md Appendix A:
//
//   dual   (edge e=(i,j)):  k1 = a*(xb_i - xb_j - dx*w1b_i - dy*w2b_i), k2 = b*(w1b_i - w1b_j),
//                           k3 = b*(w2b_i - w2b_j);  q_c <- clamp(q_c + sigma*k_c, -1, 1)
//   primal (vertex v):      g = sum over incident edges in ascending edge id of K^T q
//                           x' = x - tau*gx ; L1 shrink toward z with threshold tau*lambda*wt ; box
//                           xb = x + theta*(x - x_old)      (same for w1, w2)
//
// Layout: gathered quantities are 16-byte records (vbar, q4, ec) so every gather is one LDG.128;
// per-vertex private state is planar fp32 (coalesced 128 B per warp).  Batched streams are laid
// out block-diagonally: blockIdx.y = stream.
#pragma once

#include "common.cuh"

struct GraphView {
  float4* vbar;
  float* x;
  float* w1;
  float* w2;
  const float* z;
  const float* wt;
  const float4* ec;
  const int2* eij;
  float4* q4;
  const int32_t* row;
  const int32_t* inc;
  const int32_t* epos;  // position of edge e in its target's CSR row
  const int32_t* vnin;  // in-degree per vertex
  const int32_t* nV;
  const int32_t* nE;
  int maxV, maxE;
  int only;  // >= 0: process this stream only (fb_update solves one stream of a batch)
};

static GraphView graph_view(fb_ctx* c) {
  GraphView g;
  g.vbar = c->vbar; g.x = c->x; g.w1 = c->w1; g.w2 = c->w2; g.z = c->z; g.wt = c->wt;
  g.ec = c->ec; g.eij = c->eij; g.q4 = c->q4; g.row = c->row; g.inc = c->inc;
  g.epos = c->epos; g.vnin = c->vnin;
  g.nV = c->nV; g.nE = c->nE; g.maxV = c->maxV; g.maxE = c->maxE;
  g.only = -1;
  return g;
}

__device__ __forceinline__ float fb_clamp1(float t) { return fminf(fmaxf(t, -1.0f), 1.0f); }

// One thread per edge.  Algorithmic traffic per edge: eij 8 + ec 16 + q 16 r + 16 w + 2 gathers.
__global__ void __launch_bounds__(256) k_dual_edges(GraphView g, float sigma) {
  const int s = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.nE[s] || (g.only >= 0 && s != g.only)) return;
  const size_t eb = (size_t)s * g.maxE + e;
  const size_t vb = (size_t)s * g.maxV;
  const int2 ij = g.eij[eb];
  const float4 c = g.ec[eb];
  float4 q = g.q4[eb];
  const float4 bi = g.vbar[vb + ij.x];
  const float4 bj = g.vbar[vb + ij.y];
  float t = bi.x - bj.x;
  t = fmaf(-c.z, bi.y, t);
  t = fmaf(-c.w, bi.z, t);
  const float k1 = c.x * t;
  const float k2 = c.y * (bi.y - bj.y);
  const float k3 = c.y * (bi.z - bj.z);
  q.x = fb_clamp1(fmaf(sigma, k1, q.x));
  q.y = fb_clamp1(fmaf(sigma, k2, q.y));
  q.z = fb_clamp1(fmaf(sigma, k3, q.z));
  g.q4[eb] = q;
}

// One thread per vertex; deterministic CSR gather (no atomics), fused prox + box + extragradient.
__global__ void __launch_bounds__(256)
k_primal_vertices(GraphView g, float tau, float tl, float theta, float xmin, float xmax) {
  const int s = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.nV[s] || (g.only >= 0 && s != g.only)) return;
  const size_t vb = (size_t)s * g.maxV + v;
  const size_t eb = (size_t)s * g.maxE;
  const int32_t* row = g.row + (size_t)s * (g.maxV + 1);
  const int32_t* inc = g.inc + (size_t)s * 2 * g.maxE;
  const int r0 = row[v], r1 = row[v + 1];
  float gx = 0.0f, g1 = 0.0f, g2 = 0.0f;
  for (int r = r0; r < r1; ++r) {
    const int code = inc[r];
    const size_t e = eb + (code >> 1);
    const float4 q = g.q4[e];
    const float4 c = g.ec[e];
    const float a1 = c.x * q.x;
    if ((code & 1) == 0) {
      gx += a1;
      g1 += fmaf(c.y, q.y, -(c.z * a1));
      g2 += fmaf(c.y, q.z, -(c.w * a1));
    } else {
      gx -= a1;
      g1 -= c.y * q.y;
      g2 -= c.y * q.z;
    }
  }
  const float xo = g.x[vb], w1o = g.w1[vb], w2o = g.w2[vb];
  const float xp = fmaf(-tau, gx, xo);
  const float w1n = fmaf(-tau, g1, w1o);
  const float w2n = fmaf(-tau, g2, w2o);
  const float th = tl * g.wt[vb];
  const float zz = g.z[vb];
  const float d = xp - zz;
  float xn = (d > th) ? (xp - th) : ((d < -th) ? (xp + th) : zz);
  xn = fminf(fmaxf(xn, xmin), xmax);
  g.x[vb] = xn;
  g.w1[vb] = w1n;
  g.w2[vb] = w2n;
  g.vbar[vb] = make_float4(fmaf(theta, xn - xo, xn), fmaf(theta, w1n - w1o, w1n),
                           fmaf(theta, w2n - w2o, w2n), 0.0f);
}

// nltgv2_total_{smoothness,data}_cost: per-term fp32, accumulated in fp64.
__global__ void __launch_bounds__(256) k_costs(GraphView g, float data_factor, double* costs) {
  const int s = blockIdx.y;
  const int nE = g.nE[s], nV = g.nV[s];
  const size_t vb = (size_t)s * g.maxV, eb = (size_t)s * g.maxE;
  double sm = 0.0, da = 0.0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += gridDim.x * blockDim.x) {
    const int2 ij = g.eij[eb + e];
    const float4 c = g.ec[eb + e];
    float t = g.x[vb + ij.x] - g.x[vb + ij.y];
    t = fmaf(-c.z, g.w1[vb + ij.x], t);
    t = fmaf(-c.w, g.w2[vb + ij.x], t);
    const float k1 = c.x * t;
    const float k2 = c.y * (g.w1[vb + ij.x] - g.w1[vb + ij.y]);
    const float k3 = c.y * (g.w2[vb + ij.x] - g.w2[vb + ij.y]);
    sm += (double)(fabsf(k1) + fabsf(k2) + fabsf(k3));
  }
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += gridDim.x * blockDim.x)
    da += (double)((data_factor * g.wt[vb + v]) * fabsf(g.x[vb + v] - g.z[vb + v]));
  for (int o = 16; o > 0; o >>= 1) {
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
    da += __shfl_xor_sync(0xffffffffu, da, o);
  }
  __shared__ double ssm[8], sda[8];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    ssm[wid] = sm;
    sda[wid] = da;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      a += ssm[k];
      b += sda[k];
    }
    atomicAdd(&costs[2 * s], a);
    atomicAdd(&costs[2 * s + 1], b);
  }
}

// Data-term assembly (sync_graph, row a10): z[v] = mu[f(v)], wt = adaptive ? 1/var : 1.
// Dead / unbound features keep the old z with weight 0.
__global__ void __launch_bounds__(256)
k_data_from_features(float* z, float* wt, const int32_t* vfeat, const int32_t* nV, int maxV,
                     const float* mu, const float* var, const int32_t* alive, const int32_t* nF,
                     int maxF, int adaptive) {
  const int s = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV[s]) return;
  const size_t vb = (size_t)s * maxV + v;
  const int f = vfeat[vb];
  if (f < 0 || f >= nF[s] || !alive[(size_t)s * maxF + f]) {
    wt[vb] = 0.0f;
    return;
  }
  const size_t fb = (size_t)s * maxF + f;
  z[vb] = mu[fb];
  wt[vb] = adaptive ? (1.0f / var[fb]) : 1.0f;
}

// x = z (cold start), w = 0, bar = (x, w).
__global__ void __launch_bounds__(256)
k_state_init(GraphView g, int s, int V, int E, int x_from_z, int zero_w, int zero_q) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t vb = (size_t)s * g.maxV, eb = (size_t)s * g.maxE;
  if (t < V) {
    if (x_from_z) g.x[vb + t] = g.z[vb + t];
    if (zero_w) {
      g.w1[vb + t] = 0.0f;
      g.w2[vb + t] = 0.0f;
    }
    g.vbar[vb + t] = make_float4(g.x[vb + t], g.w1[vb + t], g.w2[vb + t], 0.0f);
  }
  if (zero_q && t < E) g.q4[eb + t] = make_float4(0.f, 0.f, 0.f, 0.f);
}
