// nltgv2_cluster.cuh -- persistent thread-block-cluster NLTGV2-L1 solver (variant 2).
//
// One cluster of C CTAs (C in {1,2,4,8,16}) owns one stream's graph for ALL iterations of a solve:
//   * the graph is cut into C contiguous vertex ranges; an edge belongs to the CTA of its source
//     vertex (edges are sorted by (i,j), so each CTA's edges are one contiguous range too);
//   * per-edge state (q, alpha, beta, dx, dy, addresses) and per-vertex state (x, w, z, threshold)
//     live in REGISTERS of the owning thread for the whole solve;
//   * the only data exchanged are the extragradient points (16 B / vertex, in shared memory, read
//     by edge threads -- remotely over DSMEM for cut edges) and the K^T q contributions (16 B per
//     vertex-edge incidence, pushed by the edge thread into the slot of the owning vertex, local
//     or remote, in the vertex's CSR order so the accumulation order equals the streaming kernel's);
//   * two hardware cluster barriers per iteration replace two kernel launches.
// HBM is touched once per solve (state in, state out) instead of once per iteration; the
// extragradient tile is staged in/out of shared memory with TMA bulk copies (cp.async.bulk).
// Arithmetic and summation order are identical to nltgv2.cuh, so both variants agree bit for bit.
#pragma once

#include <algorithm>

#include "common.cuh"
#include "nltgv2.cuh"

#define FBC_THREADS 1024
#define FBC_EPT 2   // edges per thread (register resident)
#define FBC_VPT 1   // vertices per thread
#define FBC_MAXC 16
#define FBC_SMEM_LIMIT (227 * 1024)

struct ClusterPlan {
  int C = 0;                 // cluster size the device plan was built for (0 = none)
  int capV = 0, capI = 0;    // per-CTA capacities (max over streams and ranks) used for the layout
  int4* eplan = nullptr;     // [S*maxE] {i_local, j_rank<<20|j_local, slot_i, j_rank<<20|slot_j}
  int32_t* vpart = nullptr;  // [S*(FBC_MAXC+1)] vertex range boundaries per rank
  int32_t* epart = nullptr;  // [S*(FBC_MAXC+1)] edge range boundaries per rank
  // host copies of every stream's topology so plans can be rebuilt when C changes
  struct Topo {
    int V = 0, E = 0;
    std::vector<int2> eij;
    std::vector<int32_t> row, inc;
    int needC = 1;      // smallest feasible cluster size for this graph (0 = does not fit)
    bool dirty = true;  // device plan out of date
    int capV = 0, capI = 0;
  };
  std::vector<Topo> topo;
};

// ---------------------------------------------------------------------------------- device side
__device__ __forceinline__ uint32_t fbc_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t fbc_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 fbc_ld_cluster(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void fbc_st_cluster(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void fbc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}
__device__ __forceinline__ uint32_t fbc_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t fbc_cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

struct ClusterArgs {
  GraphView g;
  const int4* eplan;
  const int32_t* vpart;
  const int32_t* epart;
  int capV, capI;
};

__global__ void __launch_bounds__(FBC_THREADS, 1)
k_nltgv2_cluster(ClusterArgs a, int iters, float sigma, float tau, float tl, float theta,
                 float xmin, float xmax) {
  extern __shared__ __align__(128) uint8_t fbc_smem[];
  float4* s_bar = reinterpret_cast<float4*>(fbc_smem);              // [capV]
  float4* s_slot = s_bar + a.capV;                                  // [capI]
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_slot + a.capI);  // TMA completion barrier
  const GraphView& g = a.g;
  const int tid = threadIdx.x;
  const uint32_t C = fbc_cluster_nctarank(), rank = fbc_cluster_ctarank();
  const int s = blockIdx.x / C;
  if (g.nV[s] == 0) return;  // uniform over the cluster
  const int32_t* vp = a.vpart + (size_t)s * (FBC_MAXC + 1);
  const int32_t* ep = a.epart + (size_t)s * (FBC_MAXC + 1);
  const int v0 = vp[rank], v1 = vp[rank + 1], e0 = ep[rank], e1 = ep[rank + 1];
  const int Vc = v1 - v0;
  const size_t vb = (size_t)s * g.maxV, eb = (size_t)s * g.maxE;
  const int32_t* row = g.row + (size_t)s * (g.maxV + 1);
  const int slot0 = row[v0];

  // ---- stage this CTA's extragradient tile with one TMA bulk copy ---------------------------
  const uint32_t mbar = fbc_smem_u32(s_mbar);
  const uint32_t bar_bytes = (uint32_t)Vc * 16u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (bar_bytes) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar),
                   "r"(bar_bytes)
                   : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
              "r"(fbc_smem_u32(s_bar)),
          "l"(g.vbar + vb + v0), "r"(bar_bytes), "r"(mbar)
          : "memory");
    }
  }

  // ---- register-resident per-edge and per-vertex state (coalesced global loads) -------------
  float q1[FBC_EPT], q2[FBC_EPT], q3[FBC_EPT], ea[FBC_EPT], ebt[FBC_EPT], edx[FBC_EPT], edy[FBC_EPT];
  uint32_t a_bi[FBC_EPT], a_bj[FBC_EPT], a_si[FBC_EPT], a_sj[FBC_EPT];
  bool ev[FBC_EPT];
#pragma unroll
  for (int k = 0; k < FBC_EPT; ++k) {
    const int e = e0 + tid + k * FBC_THREADS;
    ev[k] = e < e1;
    q1[k] = q2[k] = q3[k] = ea[k] = ebt[k] = edx[k] = edy[k] = 0.f;
    a_bi[k] = a_bj[k] = a_si[k] = a_sj[k] = 0u;
    if (ev[k]) {
      const int4 pl = a.eplan[eb + e];
      const float4 c = g.ec[eb + e];
      const float4 q = g.q4[eb + e];
      ea[k] = c.x; ebt[k] = c.y; edx[k] = c.z; edy[k] = c.w;
      q1[k] = q.x; q2[k] = q.y; q3[k] = q.z;
      a_bi[k] = fbc_smem_u32(s_bar + pl.x);
      a_bj[k] = fbc_mapa(fbc_smem_u32(s_bar + (pl.y & 0xfffff)), (uint32_t)pl.y >> 20);
      a_si[k] = fbc_smem_u32(s_slot + pl.z);
      a_sj[k] = fbc_mapa(fbc_smem_u32(s_slot + (pl.w & 0xfffff)), (uint32_t)pl.w >> 20);
    }
  }
  float vx[FBC_VPT], vw1[FBC_VPT], vw2[FBC_VPT], vz[FBC_VPT], vth[FBC_VPT];
  int vs0[FBC_VPT], vs1[FBC_VPT];
  bool vv[FBC_VPT];
#pragma unroll
  for (int k = 0; k < FBC_VPT; ++k) {
    const int v = v0 + tid + k * FBC_THREADS;
    vv[k] = v < v1;
    vx[k] = vw1[k] = vw2[k] = vz[k] = vth[k] = 0.f;
    vs0[k] = vs1[k] = 0;
    if (vv[k]) {
      vx[k] = g.x[vb + v]; vw1[k] = g.w1[vb + v]; vw2[k] = g.w2[vb + v];
      vz[k] = g.z[vb + v];
      vth[k] = tl * g.wt[vb + v];
      vs0[k] = row[v] - slot0;
      vs1[k] = row[v + 1] - slot0;
    }
  }
  __syncthreads();  // mbarrier init visible to all threads of the CTA
  if (bar_bytes) {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(mbar)
          : "memory");
    }
  }
  fbc_cluster_sync();  // every CTA's tile is resident before any remote read

  for (int it = 0; it < iters; ++it) {
    // ---- dual half-step: edge threads, gather bar (local + DSMEM), push K^T q contributions ----
#pragma unroll
    for (int k = 0; k < FBC_EPT; ++k) {
      if (ev[k]) {
        const float4 bi = fbc_ld_cluster(a_bi[k]);
        const float4 bj = fbc_ld_cluster(a_bj[k]);
        float t = bi.x - bj.x;
        t = fmaf(-edx[k], bi.y, t);
        t = fmaf(-edy[k], bi.z, t);
        const float k1 = ea[k] * t;
        const float k2 = ebt[k] * (bi.y - bj.y);
        const float k3 = ebt[k] * (bi.z - bj.z);
        q1[k] = fb_clamp1(fmaf(sigma, k1, q1[k]));
        q2[k] = fb_clamp1(fmaf(sigma, k2, q2[k]));
        q3[k] = fb_clamp1(fmaf(sigma, k3, q3[k]));
        const float a1 = ea[k] * q1[k];
        fbc_st_cluster(a_si[k], make_float4(a1, fmaf(ebt[k], q2[k], -(edx[k] * a1)),
                                            fmaf(ebt[k], q3[k], -(edy[k] * a1)), 0.f));
        fbc_st_cluster(a_sj[k], make_float4(-a1, -(ebt[k] * q2[k]), -(ebt[k] * q3[k]), 0.f));
      }
    }
    fbc_cluster_sync();
    // ---- primal half-step: vertex threads, local slot gather in CSR order ---------------------
#pragma unroll
    for (int k = 0; k < FBC_VPT; ++k) {
      if (vv[k]) {
        float gx = 0.f, g1 = 0.f, g2 = 0.f;
        for (int r = vs0[k]; r < vs1[k]; ++r) {
          const float4 c = s_slot[r];
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
        s_bar[tid + k * FBC_THREADS] =
            make_float4(fmaf(theta, xn - xo, xn), fmaf(theta, w1n - w1o, w1n),
                        fmaf(theta, w2n - w2o, w2n), 0.f);
      }
    }
    fbc_cluster_sync();
  }

  // ---- write back: registers -> global (coalesced), extragradient tile via TMA bulk store ----
#pragma unroll
  for (int k = 0; k < FBC_EPT; ++k)
    if (ev[k]) g.q4[eb + e0 + tid + k * FBC_THREADS] = make_float4(q1[k], q2[k], q3[k], 0.f);
#pragma unroll
  for (int k = 0; k < FBC_VPT; ++k)
    if (vv[k]) {
      const size_t v = vb + v0 + tid + k * FBC_THREADS;
      g.x[v] = vx[k]; g.w1[v] = vw1[k]; g.w2[v] = vw2[k];
    }
  if (tid == 0 && bar_bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g.vbar + vb + v0),
                 "r"(fbc_smem_u32(s_bar)), "r"(bar_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
static inline size_t fbc_smem_bytes(int capV, int capI) {
  return 16 * ((size_t)capV + (size_t)capI) + 64;
}

// Balanced contiguous partition into C vertex ranges; returns false when a range breaks the
// per-thread register budget or the shared-memory capacity.
static bool fbc_partition(const ClusterPlan::Topo& t, int C, std::vector<int>& vpart,
                          std::vector<int>& epart, int& capV, int& capI) {
  const int V = t.V, E = t.E;
  vpart.assign(C + 1, V);
  epart.assign(C + 1, E);
  vpart[0] = 0;
  epart[0] = 0;
  // first edge owned by each vertex (edges sorted by source)
  std::vector<int> first(V + 1, E);
  for (int e = E - 1; e >= 0; --e) first[t.eij[e].x] = e;
  for (int v = V - 1; v >= 0; --v) first[v] = std::min(first[v], first[v + 1]);
  // cost model: a vertex costs its slot gather (degree) + 4, an owned edge costs 8
  std::vector<double> pre(V + 1, 0.0);
  for (int v = 0; v < V; ++v) {
    const int deg = t.row[v + 1] - t.row[v];
    const int owned = first[v + 1] - first[v];
    pre[v + 1] = pre[v] + 4.0 + deg + 8.0 * owned;
  }
  int v = 0;
  for (int r = 1; r < C; ++r) {
    const double target = pre[V] * r / C;
    while (v < V && pre[v] < target) ++v;
    vpart[r] = v;
  }
  capV = capI = 0;
  for (int r = 0; r < C; ++r) {
    epart[r] = first[vpart[r]];
    epart[r + 1] = first[vpart[r + 1]];
    const int Vc = vpart[r + 1] - vpart[r];
    const int Ec = epart[r + 1] - epart[r];
    const int Ic = t.row[vpart[r + 1]] - t.row[vpart[r]];
    if (Vc > FBC_VPT * FBC_THREADS || Ec > FBC_EPT * FBC_THREADS) return false;
    capV = std::max(capV, Vc);
    capI = std::max(capI, Ic);
  }
  return fbc_smem_bytes(capV, capI) <= FBC_SMEM_LIMIT;
}

static int cluster_plan_build(fb_ctx* c, int s, int V, int E, const int2* eij, const int32_t* row,
                              const int32_t* inc) {
  if (!c->plan) {
    c->plan = new ClusterPlan();
    c->plan->topo.resize(c->S);
    if (dalloc(&c->plan->eplan, (size_t)c->S * c->maxE) != cudaSuccess ||
        dalloc(&c->plan->vpart, (size_t)c->S * (FBC_MAXC + 1)) != cudaSuccess ||
        dalloc(&c->plan->epart, (size_t)c->S * (FBC_MAXC + 1)) != cudaSuccess)
      FB_FAIL(c, FB_E_NOMEM, "cluster plan allocation failed");
  }
  ClusterPlan::Topo& t = c->plan->topo[s];
  t.V = V;
  t.E = E;
  t.eij.assign(eij, eij + E);
  t.row.assign(row, row + V + 1);
  t.inc.assign(inc, inc + 2 * (size_t)E);
  t.dirty = true;
  t.needC = 0;
  std::vector<int> vp, ep;
  for (int C = 1; C <= FBC_MAXC; C *= 2) {
    int capV, capI;
    if (fbc_partition(t, C, vp, ep, capV, capI)) {
      t.needC = C;
      break;
    }
  }
  return FB_OK;
}

static bool cluster_plan_ready(fb_ctx* c) {
  if (!c->plan) return false;
  for (int s = 0; s < c->S; ++s)
    if (c->plan->topo[s].V > 0 && c->plan->topo[s].needC == 0) return false;
  return true;
}

static void cluster_plan_free(fb_ctx* c) {
  if (!c->plan) return;
  cudaFree(c->plan->eplan);
  cudaFree(c->plan->vpart);
  cudaFree(c->plan->epart);
  delete c->plan;
  c->plan = nullptr;
}

// (Re)build the device plans of every dirty stream for cluster size C.
static int fbc_upload_plans(fb_ctx* c, int C) {
  ClusterPlan* P = c->plan;
  if (P->C != C)
    for (auto& t : P->topo) t.dirty = true;
  P->C = C;
  std::vector<int> vp, ep;
  std::vector<int4> eplan;
  std::vector<int32_t> pad(FBC_MAXC + 1);
  for (int s = 0; s < c->S; ++s) {
    ClusterPlan::Topo& t = P->topo[s];
    if (!t.dirty) continue;
    if (t.V == 0) {  // empty stream: its cluster exits on nV == 0, but keep the tables defined
      std::fill(pad.begin(), pad.end(), 0);
      FB_CUDA(c, cudaMemcpyAsync(P->vpart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, c->stream));
      FB_CUDA(c, cudaMemcpyAsync(P->epart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, c->stream));
      FB_CUDA(c, cudaStreamSynchronize(c->stream));
      t.capV = t.capI = 0;
      t.dirty = false;
      continue;
    }
    int capV = 0, capI = 0;
    if (!fbc_partition(t, C, vp, ep, capV, capI))
      FB_FAIL(c, FB_E_STATE, "cluster plan: partition infeasible");
    t.capV = capV;
    t.capI = capI;
    // rank of every vertex
    std::vector<int> rk(t.V);
    for (int r = 0; r < C; ++r)
      for (int v = vp[r]; v < vp[r + 1]; ++v) rk[v] = r;
    eplan.assign(t.E, make_int4(0, 0, 0, 0));
    for (int v = 0; v < t.V; ++v) {
      const int r = rk[v], base = t.row[vp[r]];
      for (int k = t.row[v]; k < t.row[v + 1]; ++k) {
        const int code = t.inc[k], e = code >> 1, local = k - base;
        if ((code & 1) == 0) {
          eplan[e].x = v - vp[r];
          eplan[e].z = local;
        } else {
          eplan[e].y = (r << 20) | (v - vp[r]);
          eplan[e].w = (r << 20) | local;
        }
      }
    }
    cudaStream_t st = c->stream;
    if (t.E)
      FB_CUDA(c, cudaMemcpyAsync(P->eplan + (size_t)s * c->maxE, eplan.data(), sizeof(int4) * t.E, cudaMemcpyHostToDevice, st));
    for (int r = 0; r <= FBC_MAXC; ++r) pad[r] = vp[std::min(r, C)];
    FB_CUDA(c, cudaMemcpyAsync(P->vpart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, st));
    for (int r = 0; r <= FBC_MAXC; ++r) pad[r] = ep[std::min(r, C)];
    FB_CUDA(c, cudaMemcpyAsync(P->epart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaStreamSynchronize(st));  // staging vectors are reused per stream
    t.dirty = false;
  }
  P->capV = P->capI = 1;
  for (auto& t : P->topo) {
    if (t.V == 0) continue;
    P->capV = std::max(P->capV, t.capV);
    P->capI = std::max(P->capI, t.capI);
  }
  return FB_OK;
}

static int solve_cluster(fb_ctx* c, int iters, const fb_nltgv2_params* p) {
  ClusterPlan* P = c->plan;
  int C = 1;
  bool any = false;
  for (auto& t : P->topo)
    if (t.V > 0) {
      C = std::max(C, t.needC);
      any = true;
    }
  if (!any) return FB_OK;
  int rc = fbc_upload_plans(c, C);
  if (rc) return rc;
  const size_t smem = fbc_smem_bytes(P->capV, P->capI);
  FB_CUDA(c, cudaFuncSetAttribute(k_nltgv2_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (C > 8) FB_CUDA(c, cudaFuncSetAttribute(k_nltgv2_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  ClusterArgs a;
  a.g = graph_view(c);
  a.eplan = P->eplan;
  a.vpart = P->vpart;
  a.epart = P->epart;
  a.capV = P->capV;
  a.capI = P->capI;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(c->S * C));
  cfg.blockDim = dim3(FBC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float tl = p->step_x * p->data_factor;
  FB_CUDA(c, cudaLaunchKernelEx(&cfg, k_nltgv2_cluster, a, iters, p->step_q, p->step_x, tl, p->theta, p->x_min, p->x_max));
  c->launches++;
  return FB_OK;
}
