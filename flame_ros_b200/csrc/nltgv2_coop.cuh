// nltgv2_coop.cuh -- plan-free resident NLTGV2-L1 solver (variant 4): all iterations of a solve in
// ONE launch, for graphs whose topology was just rebuilt (flame::Flame::update re-triangulates every
// frame: ~2 % of the edges change from one frame to the next, so the per-topology tables of variants
// 2 / 3 can never be reused and cost more host time than they save).
//
// Same arithmetic as nltgv2.cuh (bit-identical), different schedule:
//   * one thread-block cluster (8 or 16 CTAs x 512 threads) per stream; a thread owns up to EPT edges
//     and VPT vertices by index (edge e -> thread e mod NT): NO partition, NO per-topology tables;
//   * edge state (q, alpha, beta, delta, endpoints) and vertex state (x, w, z, threshold) stay in
//     registers for the whole solve; the CSR incidence of a vertex sits in shared memory;
//   * what crosses threads goes through L2: the extragradient points (vbar, one 16 B record per
//     vertex) and the K^T q contributions (one 16 B record per edge end, indexed by the CSR code
//     (edge << 1 | role), so a vertex sums them in CSR order exactly like k_primal_vertices);
//   * the two half-steps are separated by the cluster's hardware barrier (release / acquire at
//     cluster scope); exchanged data is read with ld.global.cg, so no stale L1 line can be seen.
// HBM is touched once per solve; the per-iteration traffic (64 B per edge + ~130 B per vertex) stays
// in L2.  Bound: two cluster barriers + two L2 round trips per iteration.
#pragma once

#include "common.cuh"
#include "nltgv2.cuh"
#include "nltgv2_cluster.cuh"

#define FBK_THREADS 512
#define FBK_INC 12   // CSR entries per vertex held in shared memory (the rest is re-read through L2)

template <int EPT, int VPT>
__global__ void __launch_bounds__(FBK_THREADS, 1)
k_nltgv2_coop(GraphView g, float4* __restrict__ contrib, int* __restrict__ err, int iters, float sigma, float tau,
              float tl, float theta, float xmin, float xmax) {
  __shared__ int s_inc[VPT][FBK_INC][FBK_THREADS];
  const int C = (int)fbc_cluster_nctarank(), rank = (int)fbc_cluster_ctarank();
  const int s = (g.only >= 0) ? g.only : (int)blockIdx.x / C;
  const int nV = g.nV[s], nE = g.nE[s];
  if (nV == 0) return;  // uniform over the cluster
  const int NT = C * FBK_THREADS, tid = threadIdx.x, gt = rank * FBK_THREADS + tid;
  if (nE > EPT * NT || nV > VPT * NT) {  // host checks the capacities; never expected
    if (gt == 0) *err = 1;
    return;
  }
  const size_t vb = (size_t)s * g.maxV, eb = (size_t)s * g.maxE;
  float4* vbar = g.vbar + vb;
  float4* cb = contrib + 2 * eb;
  const int32_t* inc = g.inc + 2 * eb;
  const int32_t* row = g.row + (size_t)s * (g.maxV + 1);

  float q1[EPT], q2[EPT], q3[EPT], ea[EPT], ebt[EPT], dx[EPT], dy[EPT];
  int ei[EPT], ej[EPT];
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int e = gt + k * NT;
    ei[k] = -1; ej[k] = 0;
    q1[k] = q2[k] = q3[k] = ea[k] = ebt[k] = dx[k] = dy[k] = 0.f;
    if (e < nE) {
      const int2 ij = g.eij[eb + e];
      const float4 c = g.ec[eb + e];
      const float4 q = g.q4[eb + e];
      ei[k] = ij.x; ej[k] = ij.y;
      ea[k] = c.x; ebt[k] = c.y; dx[k] = c.z; dy[k] = c.w;
      q1[k] = q.x; q2[k] = q.y; q3[k] = q.z;
    }
  }
  float vx[VPT], vw1[VPT], vw2[VPT], vz[VPT], vth[VPT];
  int r0[VPT], r1[VPT];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = gt + k * NT;
    r0[k] = r1[k] = 0;
    vx[k] = vw1[k] = vw2[k] = vz[k] = vth[k] = 0.f;
    if (v < nV) {
      vx[k] = g.x[vb + v]; vw1[k] = g.w1[vb + v]; vw2[k] = g.w2[vb + v];
      vz[k] = g.z[vb + v];
      vth[k] = tl * g.wt[vb + v];
      r0[k] = row[v]; r1[k] = row[v + 1];
      for (int r = r0[k], j = 0; r < r1[k] && j < FBK_INC; ++r, ++j) s_inc[k][j][tid] = inc[r];
    }
  }
  // the extragradient points this solve starts from were written by earlier kernels of the stream

  for (int it = 0; it < iters; ++it) {
    // ---- dual half-step: loads first, then the arithmetic of nltgv2.cuh:k_dual_edges
    float4 bi[EPT], bj[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k)
      if (ei[k] >= 0) {
        bi[k] = __ldcg(vbar + ei[k]);
        bj[k] = __ldcg(vbar + ej[k]);
      }
#pragma unroll
    for (int k = 0; k < EPT; ++k)
      if (ei[k] >= 0) {
        float t = bi[k].x - bj[k].x;
        t = fmaf(-dx[k], bi[k].y, t);
        t = fmaf(-dy[k], bi[k].z, t);
        const float k1 = ea[k] * t;
        const float k2 = ebt[k] * (bi[k].y - bj[k].y);
        const float k3 = ebt[k] * (bi[k].z - bj[k].z);
        q1[k] = fb_clamp1(fmaf(sigma, k1, q1[k]));
        q2[k] = fb_clamp1(fmaf(sigma, k2, q2[k]));
        q3[k] = fb_clamp1(fmaf(sigma, k3, q3[k]));
        const float a1 = ea[k] * q1[k];
        const int e = gt + k * NT;
        // K^T q of this edge for its source (role 0) and its target (role 1)
        cb[2 * e] = make_float4(a1, fmaf(ebt[k], q2[k], -(dx[k] * a1)), fmaf(ebt[k], q3[k], -(dy[k] * a1)), 0.f);
        cb[2 * e + 1] = make_float4(-a1, -(ebt[k] * q2[k]), -(ebt[k] * q3[k]), 0.f);
      }
    fbc_cluster_sync();
    // ---- primal half-step (nltgv2.cuh:k_primal_vertices): CSR-order sum, prox, box, extragradient
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = gt + k * NT;
      if (v < nV) {
        float gx = 0.f, g1 = 0.f, g2 = 0.f;
        const int n = r1[k] - r0[k];
        for (int j = 0; j < n; ++j) {
          const int code = j < FBK_INC ? s_inc[k][j][tid] : __ldg(inc + r0[k] + j);
          const float4 c = __ldcg(cb + code);
          gx += c.x;
          g1 += c.y;
          g2 += c.z;
        }
        const float xo = vx[k], w1o = vw1[k], w2o = vw2[k];
        const float xp = fmaf(-tau, gx, xo);
        const float w1n = fmaf(-tau, g1, w1o);
        const float w2n = fmaf(-tau, g2, w2o);
        const float d = xp - vz[k];
        float xn = (d > vth[k]) ? (xp - vth[k]) : ((d < -vth[k]) ? (xp + vth[k]) : vz[k]);
        xn = fminf(fmaxf(xn, xmin), xmax);
        vx[k] = xn; vw1[k] = w1n; vw2[k] = w2n;
        vbar[v] = make_float4(fmaf(theta, xn - xo, xn), fmaf(theta, w1n - w1o, w1n), fmaf(theta, w2n - w2o, w2n), 0.f);
      }
    }
    if (it + 1 < iters) fbc_cluster_sync();
  }
#pragma unroll
  for (int k = 0; k < EPT; ++k)
    if (ei[k] >= 0) g.q4[eb + gt + k * NT] = make_float4(q1[k], q2[k], q3[k], 0.f);
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = gt + k * NT;
    if (v < nV) {
      g.x[vb + v] = vx[k]; g.w1[vb + v] = vw1[k]; g.w2[vb + v] = vw2[k];
    }
  }
}
