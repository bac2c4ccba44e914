// nltgv2_cluster.cuh -- persistent thread-block-cluster NLTGV2-L1 solver (variant 2).
//
// One cluster of C CTAs (C in {1,2,4,8,16}) owns one stream's graph for ALL iterations of a solve:
//   * the graph is cut into C contiguous vertex ranges; an edge belongs to the CTA of its source
//     vertex (edges are sorted by (i,j), so each CTA's edges are one contiguous range too);
//   * per-edge state (q, alpha, beta, dx, dy, addresses) and per-vertex state (x, w, z, threshold)
//     live in REGISTERS of the owning thread for the whole solve;
//   * shared memory holds the extragradient points (16 B / vertex: the CTA's own vertices followed
//     by HALO copies of the remote targets of its cut edges) and one 16 B slot per vertex-edge
//     incidence that receives the edge's K^T q contribution, in the vertex's CSR order so the
//     accumulation order equals the streaming kernel's (bit-identical results);
//   * CTAs exchange only what crosses the cut, as asynchronous DSMEM stores that signal the
//     receiver's mbarrier (st.async ... mbarrier::complete_tx::bytes): target contributions flow
//     edge-owner -> vertex-owner in the dual half-step, refreshed halo points flow back after the
//     primal half-step.  There is NO cluster-wide barrier inside the iteration loop: a CTA waits
//     only for the bytes it actually consumes, plus two CTA-local bar.syncs per iteration.
// HBM is touched once per solve (state in, state out) instead of once per iteration; the
// extragradient tile is staged in/out of shared memory with TMA bulk copies (cp.async.bulk).
#pragma once

#include <algorithm>

#include "common.cuh"
#include "nltgv2.cuh"

#ifndef FBC_THREADS
#define FBC_THREADS 512
#define FBC_EPT 4   // edges per thread (register resident)
#define FBC_VPT 2   // vertices per thread
#endif
#define FBC_MAXC 16
#define FBC_SMEM_LIMIT (227 * 1024)

struct ClusterPlan {
  int C = 0;                       // cluster size the device plan was built for (0 = none)
  int capV = 0, capH = 0, capI = 0;  // per-CTA capacities (max over streams and ranks): layout
  int4* eplan = nullptr;     // [S*maxE] {i_local, j_index, slot_i, rank<<20|slot_j}
  int2* pplan = nullptr;     // [S*maxE] overflow push list {local vertex, rank<<20|halo index}
  int4* vplan = nullptr;     // [S*maxV] per vertex (processing order): {push target 0, push target 1 (-1 = none), slot begin, slot end}
  int32_t* eorig = nullptr;  // [S*maxE] processing order -> edge id (cut edges first per CTA)
  int32_t* vorig = nullptr;  // [S*maxV] processing order -> vertex id (halo-feeding vertices first)
  int32_t* hplan = nullptr;  // [S*maxE] halo list: stream-local vertex id of every halo entry
  int32_t* vpart = nullptr;  // [S*(FBC_MAXC+1)] vertex range boundaries per rank
  int32_t* epart = nullptr;  // [S*(FBC_MAXC+1)] edge range boundaries per rank
  int4* cinfo = nullptr;     // [S*FBC_MAXC] {halo count, incoming remote slots, push begin, push end}
  int32_t* hpart = nullptr;  // [S*(FBC_MAXC+1)] halo list range per rank
  // host copies of every stream's topology so plans can be rebuilt when C changes
  struct Topo {
    int V = 0, E = 0;
    std::vector<int2> eij;
    std::vector<int32_t> row, inc;
    int needC = 1;      // smallest feasible cluster size for this graph (0 = does not fit, -1 = not computed yet)
    bool dirty = true;  // device plan out of date
    int capV = 0, capH = 0, capI = 0;
    int partC = 0;               // cluster size `part` was computed for (0 = stale)
    struct FbcPartData { std::vector<int> vpart, epart; int capV = 0, capH = 0, capI = 0; } part;
  };
  std::vector<Topo> topo;
  size_t smem_set = 0;  // dynamic shared memory size the kernel attribute is currently set to
  int max_active[FBC_MAXC + 1] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};  // co-resident clusters per size
  bool nonportable_set = false;
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
__device__ __forceinline__ float4 fbc_lds(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void fbc_sts(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
// asynchronous 16 B store into a peer CTA's shared memory, completing 16 tx-bytes on its mbarrier
__device__ __forceinline__ void fbc_st_async(uint32_t raddr, float4 v, uint32_t rmbar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::
          "r"(raddr),
      "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rmbar)
      : "memory");
}
__device__ __forceinline__ void fbc_mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void fbc_mbar_expect(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void fbc_mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
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
  const int2* pplan;
  const int4* vplan;
  const int32_t* eorig;
  const int32_t* vorig;
  const int32_t* hplan;
  const int32_t* vpart;
  const int32_t* epart;
  const int32_t* hpart;
  const int4* cinfo;
  int capV, capH, capI;
};

__global__ void __launch_bounds__(FBC_THREADS, 1)
k_nltgv2_cluster(ClusterArgs a, int iters, float sigma, float tau, float tl, float theta,
                 float xmin, float xmax) {
  extern __shared__ __align__(128) uint8_t fbc_smem[];
  float4* s_bar = reinterpret_cast<float4*>(fbc_smem);  // [capV own | capH halo]
  float4* s_slot = s_bar + a.capV + a.capH;             // [capI + 1 dummy]
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_slot + a.capI + 1);  // [0] TMA, [1] halo (A), [2] slots (B)
  const GraphView& g = a.g;
  const int tid = threadIdx.x;
  const uint32_t C = fbc_cluster_nctarank(), rank = fbc_cluster_ctarank();
  const int s = blockIdx.x / C;
  if (g.nV[s] == 0 || (g.only >= 0 && s != g.only)) return;  // uniform over the cluster
  const int32_t* vp = a.vpart + (size_t)s * (FBC_MAXC + 1);
  const int32_t* ep = a.epart + (size_t)s * (FBC_MAXC + 1);
  const int32_t* hp = a.hpart + (size_t)s * (FBC_MAXC + 1);
  const int v0 = vp[rank], v1 = vp[rank + 1], e0 = ep[rank], e1 = ep[rank + 1];
  const int Vc = v1 - v0;
  const int4 ci = a.cinfo[(size_t)s * FBC_MAXC + rank];
  const uint32_t haloBytes = 16u * (uint32_t)ci.x, slotBytes = 16u * (uint32_t)ci.y;
  const size_t vb = (size_t)s * g.maxV, eb = (size_t)s * g.maxE;

  const uint32_t mb_tma = fbc_smem_u32(s_mbar), mb_halo = mb_tma + 8, mb_slot = mb_tma + 16;
  const uint32_t bar_base = fbc_smem_u32(s_bar), slot_base = fbc_smem_u32(s_slot);
  const uint32_t bar_bytes = (uint32_t)Vc * 16u;

  // ---- stage this CTA's extragradient tile with one TMA bulk copy; arm the exchange barriers ----
  if (tid == 0) {
    fbc_mbar_init(mb_tma, 1);
    fbc_mbar_init(mb_halo, 1);
    fbc_mbar_init(mb_slot, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (bar_bytes) {
      fbc_mbar_expect(mb_tma, bar_bytes);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
              "r"(bar_base),
          "l"(g.vbar + vb + v0), "r"(bar_bytes), "r"(mb_tma)
          : "memory");
    }
    if (haloBytes && iters > 1) fbc_mbar_expect(mb_halo, haloBytes);  // halo phase 0
    if (slotBytes) fbc_mbar_expect(mb_slot, slotBytes);               // slot phase 0
  }
  // halo copies of remote targets: first value straight from global memory
  {
    const int32_t* hl = a.hplan + eb + hp[rank];
    for (int h = tid; h < ci.x; h += FBC_THREADS) s_bar[a.capV + h] = g.vbar[vb + hl[h]];
  }

  // ---- register-resident per-edge and per-vertex state (coalesced global loads) -------------
  // indices into s_bar / s_slot are kept as 16-bit pairs; only remote addresses need 32 bits
  float q1[FBC_EPT], q2[FBC_EPT], q3[FBC_EPT], ea[FBC_EPT], ebt[FBC_EPT], edx[FBC_EPT], edy[FBC_EPT];
  int e_bi[FBC_EPT], e_bj[FBC_EPT], e_si[FBC_EPT], e_id[FBC_EPT];
  uint32_t a_sj[FBC_EPT], a_mb[FBC_EPT];  // target slot: local index (a_mb == 0) or remote address
  // Processing order (eorig/vorig): a CTA's cut edges come first, so row k=0 of the unrolled edge
  // loop sends the remote contributions at the very start of the dual half-step; the vertices that
  // feed other CTAs' halos come first in the primal half-step.  The receiver then finds its bytes
  // already landed when it gets to them: the exchange overlaps with the interior work.
#pragma unroll
  for (int k = 0; k < FBC_EPT; ++k) {
    const int e = e0 + tid + k * FBC_THREADS;
    e_id[k] = -1;
    // idle lanes run the same arithmetic on zero weights and write to the dummy slot, so the
    // unrolled edges of a thread form independent, branch-free instruction streams (ILP)
    q1[k] = q2[k] = q3[k] = ea[k] = ebt[k] = edx[k] = edy[k] = 0.f;
    e_bi[k] = e_bj[k] = 0;
    e_si[k] = a.capI;
    a_sj[k] = (uint32_t)a.capI;
    a_mb[k] = 0u;
    if (e < e1) {
      const int4 pl = a.eplan[eb + e];
      const int eo = a.eorig[eb + e];
      e_id[k] = eo;
      const float4 c = g.ec[eb + eo];
      const float4 q = g.q4[eb + eo];
      ea[k] = c.x; ebt[k] = c.y; edx[k] = c.z; edy[k] = c.w;
      q1[k] = q.x; q2[k] = q.y; q3[k] = q.z;
      e_bi[k] = pl.x;
      e_bj[k] = pl.y;
      e_si[k] = pl.z;
      const uint32_t jr = (uint32_t)pl.w >> 20, sj = (uint32_t)pl.w & 0xfffffu;
      if (jr == rank) {
        a_sj[k] = sj;
      } else {
        a_sj[k] = fbc_mapa(slot_base + 16u * sj, jr);
        a_mb[k] = fbc_mapa(mb_slot, jr);
      }
    }
  }
  float vx[FBC_VPT], vw1[FBC_VPT], vw2[FBC_VPT], vz[FBC_VPT], vth[FBC_VPT];
  int vs0[FBC_VPT], vs1[FBC_VPT], v_li[FBC_VPT];
  uint32_t p_a0[FBC_VPT], p_m0[FBC_VPT], p_a1[FBC_VPT], p_m1[FBC_VPT];  // halo copies to refresh
#pragma unroll
  for (int k = 0; k < FBC_VPT; ++k) {
    const int vo = v0 + tid + k * FBC_THREADS;
    vx[k] = vw1[k] = vw2[k] = vz[k] = vth[k] = 0.f;
    vs0[k] = 0;
    vs1[k] = -1;  // marks "no vertex"
    v_li[k] = 0;
    p_a0[k] = p_m0[k] = p_a1[k] = p_m1[k] = 0u;
    if (vo < v1) {
      const int v = a.vorig[vb + vo];
      v_li[k] = v - v0;
      vx[k] = g.x[vb + v]; vw1[k] = g.w1[vb + v]; vw2[k] = g.w2[vb + v];
      vz[k] = g.z[vb + v];
      vth[k] = tl * g.wt[vb + v];
      const int4 pt = a.vplan[vb + vo];
      vs0[k] = pt.z;
      vs1[k] = pt.w;
      if (pt.x >= 0) {
        p_a0[k] = fbc_mapa(bar_base + 16u * ((uint32_t)pt.x & 0xfffffu), (uint32_t)pt.x >> 20);
        p_m0[k] = fbc_mapa(mb_halo, (uint32_t)pt.x >> 20);
      }
      if (pt.y >= 0) {
        p_a1[k] = fbc_mapa(bar_base + 16u * ((uint32_t)pt.y & 0xfffffu), (uint32_t)pt.y >> 20);
        p_m1[k] = fbc_mapa(mb_halo, (uint32_t)pt.y >> 20);
      }
    }
  }
  __syncthreads();  // mbarrier init + halo fill visible to all threads of the CTA
  if (bar_bytes) fbc_mbar_wait(mb_tma, 0);
  fbc_cluster_sync();  // every CTA's barriers are initialised before any remote store can target them

  const int2* pl = a.pplan + eb;
  for (int it = 0; it < iters; ++it) {
    const bool more = it + 1 < iters;
    // ---- dual half-step ---------------------------------------------------------------------
    if (it > 0 && haloBytes) fbc_mbar_wait(mb_halo, (uint32_t)(it - 1) & 1u);  // refreshed halo landed
#pragma unroll
    for (int k = 0; k < FBC_EPT; ++k) {
      {
        const float4 bi = s_bar[e_bi[k]];
        const float4 bj = s_bar[e_bj[k]];
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
        const float4 cs = make_float4(a1, fmaf(ebt[k], q2[k], -(edx[k] * a1)),
                                      fmaf(ebt[k], q3[k], -(edy[k] * a1)), 0.f);
        const float4 ct = make_float4(-a1, -(ebt[k] * q2[k]), -(ebt[k] * q3[k]), 0.f);
        if (e_id[k] >= 0) {  // predicated stores: idle lanes compute on zeros but write nothing
          s_slot[e_si[k]] = cs;
          if (a_mb[k] == 0u) s_slot[a_sj[k]] = ct;
          else fbc_st_async(a_sj[k], ct, a_mb[k]);
        }
      }
    }
    __syncthreads();  // local slot writes visible; every thread is past the halo wait, so the
                      // next halo phase may be armed without a late waiter seeing the parity wrap
    if (tid == 0 && it > 0 && haloBytes && more) fbc_mbar_expect(mb_halo, haloBytes);
    if (slotBytes) fbc_mbar_wait(mb_slot, (uint32_t)it & 1u);  // remote contributions have landed
    // ---- primal half-step: vertex threads, local slot gather in CSR order ---------------------
    {
      // the slot gathers of a thread's vertices advance together (one slot of each per trip) so
      // their dependent FADD chains interleave; per vertex the order is still CSR order
      float gx[FBC_VPT], g1[FBC_VPT], g2[FBC_VPT];
      int dmax = 0;
#pragma unroll
      for (int k = 0; k < FBC_VPT; ++k) {
        gx[k] = g1[k] = g2[k] = 0.f;
        dmax = max(dmax, vs1[k] - vs0[k]);
      }
      for (int j = 0; j < dmax; ++j) {
#pragma unroll
        for (int k = 0; k < FBC_VPT; ++k) {
          if (vs0[k] + j < vs1[k]) {
            const float4 c = s_slot[vs0[k] + j];
            gx[k] += c.x;
            g1[k] += c.y;
            g2[k] += c.z;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < FBC_VPT; ++k) {
        if (vs1[k] >= 0) {
          const float xo = vx[k], w1o = vw1[k], w2o = vw2[k];
          const float xp = fmaf(-tau, gx[k], xo);
          const float w1n = fmaf(-tau, g1[k], w1o);
          const float w2n = fmaf(-tau, g2[k], w2o);
          const float d = xp - vz[k];
          float xn = (d > vth[k]) ? (xp - vth[k]) : ((d < -vth[k]) ? (xp + vth[k]) : vz[k]);
          xn = fminf(fmaxf(xn, xmin), xmax);
          vx[k] = xn; vw1[k] = w1n; vw2[k] = w2n;
          const float4 nb = make_float4(fmaf(theta, xn - xo, xn), fmaf(theta, w1n - w1o, w1n),
                                        fmaf(theta, w2n - w2o, w2n), 0.f);
          s_bar[v_li[k]] = nb;
          // the owner refreshes the halo copies held by the CTAs whose cut edges point at this vertex
          if (more) {
            if (p_m0[k]) fbc_st_async(p_a0[k], nb, p_m0[k]);
            if (p_m1[k]) fbc_st_async(p_a1[k], nb, p_m1[k]);
          }
        }
      }
    }
    __syncthreads();  // own points visible to the edge threads (and to the overflow push below)
    if (tid == 0 && slotBytes && more) fbc_mbar_expect(mb_slot, slotBytes);
    if (more) {  // vertices with more than two consumer CTAs (rare): remaining copies
      for (int p = ci.z + tid; p < ci.w; p += FBC_THREADS) {
        const int2 e = pl[p];
        const uint32_t pr = (uint32_t)e.y >> 20;
        fbc_st_async(fbc_mapa(bar_base + 16u * ((uint32_t)e.y & 0xfffffu), pr), s_bar[e.x], fbc_mapa(mb_halo, pr));
      }
    }
  }

  // ---- write back: registers -> global (coalesced), extragradient tile via TMA bulk store ----
#pragma unroll
  for (int k = 0; k < FBC_EPT; ++k)
    if (e_id[k] >= 0) g.q4[eb + e_id[k]] = make_float4(q1[k], q2[k], q3[k], 0.f);
#pragma unroll
  for (int k = 0; k < FBC_VPT; ++k)
    if (vs1[k] >= 0) {
      const size_t v = vb + v0 + v_li[k];
      g.x[v] = vx[k]; g.w1[v] = vw1[k]; g.w2[v] = vw2[k];
    }
  if (tid == 0 && bar_bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g.vbar + vb + v0),
                 "r"(bar_base), "r"(bar_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  fbc_cluster_sync();  // no CTA retires while a peer could still address its shared memory
}

// ---------------------------------------------------------------------------------- host side
static inline size_t fbc_smem_bytes(int capV, int capH, int capI) {
  return 16 * ((size_t)capV + (size_t)capH + (size_t)capI + 1) + 64;  // +1: dummy slot of idle lanes
}

typedef ClusterPlan::Topo::FbcPartData FbcPart;

// Balanced contiguous partition into C vertex ranges; returns false when a range breaks the
// per-thread register budget or the shared-memory capacity.
static bool fbc_partition(const ClusterPlan::Topo& t, int C, FbcPart& P) {
  const int V = t.V, E = t.E;
  P.vpart.assign(C + 1, V);
  P.epart.assign(C + 1, E);
  P.vpart[0] = 0;
  P.epart[0] = 0;
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
    P.vpart[r] = v;
  }
  P.capV = P.capI = P.capH = 0;
  std::vector<int> seen(V, -1);
  for (int r = 0; r < C; ++r) {
    P.epart[r] = first[P.vpart[r]];
    P.epart[r + 1] = first[P.vpart[r + 1]];
    const int Vc = P.vpart[r + 1] - P.vpart[r];
    const int Ec = P.epart[r + 1] - P.epart[r];
    int Ic = 0;  // slot blocks are padded to an odd number of 16 B records (bank-conflict-free gathers)
    for (int v = P.vpart[r]; v < P.vpart[r + 1]; ++v) Ic += (t.row[v + 1] - t.row[v]) | 1;
    if (Vc > FBC_VPT * FBC_THREADS || Ec > FBC_EPT * FBC_THREADS) return false;
    int H = 0;
    for (int e = P.epart[r]; e < P.epart[r + 1]; ++e) {
      const int j = t.eij[e].y;
      if (j >= P.vpart[r + 1] && seen[j] != r) {
        seen[j] = r;
        ++H;
      }
    }
    P.capV = std::max(P.capV, Vc);
    P.capI = std::max(P.capI, Ic);
    P.capH = std::max(P.capH, H);
  }
  return fbc_smem_bytes(P.capV, P.capH, P.capI) <= FBC_SMEM_LIMIT;
}

static int cluster_plan_build(fb_ctx* c, int s, int V, int E, const int2* eij, const int32_t* row,
                              const int32_t* inc) {
  if (!c->plan) {
    c->plan = new ClusterPlan();
    c->plan->topo.resize(c->S);
    const size_t S = c->S;
    if (dalloc(&c->plan->eplan, S * c->maxE) != cudaSuccess ||
        dalloc(&c->plan->pplan, S * c->maxE) != cudaSuccess ||
        dalloc(&c->plan->vplan, S * c->maxV) != cudaSuccess ||
        dalloc(&c->plan->eorig, S * c->maxE) != cudaSuccess ||
        dalloc(&c->plan->vorig, S * c->maxV) != cudaSuccess ||
        dalloc(&c->plan->hplan, S * c->maxE) != cudaSuccess ||
        dalloc(&c->plan->vpart, S * (FBC_MAXC + 1)) != cudaSuccess ||
        dalloc(&c->plan->epart, S * (FBC_MAXC + 1)) != cudaSuccess ||
        dalloc(&c->plan->hpart, S * (FBC_MAXC + 1)) != cudaSuccess ||
        dalloc(&c->plan->cinfo, S * FBC_MAXC) != cudaSuccess)
      FB_FAIL(c, FB_E_NOMEM, "cluster plan allocation failed");
  }
  ClusterPlan::Topo& t = c->plan->topo[s];
  t.V = V;
  t.E = E;
  t.eij.assign(eij, eij + E);
  t.row.assign(row, row + V + 1);
  t.inc.assign(inc, inc + 2 * (size_t)E);
  t.dirty = true;
  t.needC = -1;  // computed on first use: fb_update re-sets the graph every frame and never asks
  t.partC = 0;
  return FB_OK;
}

static void fbc_need(ClusterPlan::Topo& t) {
  if (t.needC >= 0) return;
  t.needC = 0;
  FbcPart P;
  for (int C = 1; C <= FBC_MAXC; C *= 2) {
    if (fbc_partition(t, C, P)) {
      t.needC = C;
      break;
    }
  }
}

static bool cluster_plan_ready(fb_ctx* c) {
  if (!c->plan) return false;
  for (int s = 0; s < c->S; ++s)
    if (c->plan->topo[s].V > 0) {
      fbc_need(c->plan->topo[s]);
      if (c->plan->topo[s].needC == 0) return false;
    }
  return true;
}

static void cluster_plan_free(fb_ctx* c) {
  if (!c->plan) return;
  cudaFree(c->plan->eplan);
  cudaFree(c->plan->pplan);
  cudaFree(c->plan->vplan);
  cudaFree(c->plan->eorig);
  cudaFree(c->plan->vorig);
  cudaFree(c->plan->hplan);
  cudaFree(c->plan->vpart);
  cudaFree(c->plan->epart);
  cudaFree(c->plan->hpart);
  cudaFree(c->plan->cinfo);
  delete c->plan;
  c->plan = nullptr;
}

// (Re)build the device plans of every dirty stream for cluster size C.  Because the shared-memory
// layout (capV, capH) is baked into the j indices, a change of the context-wide capacities
// invalidates every stream's plan.
static int fbc_upload_plans(fb_ctx* c, int C) {
  ClusterPlan* P = c->plan;
  if (P->C != C)
    for (auto& t : P->topo) t.dirty = true;
  P->C = C;
  // pass 1: partitions + capacities
  int capV = 1, capH = 0, capI = 1;
  for (int s = 0; s < c->S; ++s) {
    ClusterPlan::Topo& t = P->topo[s];
    if (t.V == 0) continue;
    if (t.partC != C) {  // partitions are cached per topology: recomputed only after fb_graph_set
      if (!fbc_partition(t, C, t.part)) FB_FAIL(c, FB_E_STATE, "cluster plan: partition infeasible");
      t.partC = C;
    }
    capV = std::max(capV, t.part.capV);
    capH = std::max(capH, t.part.capH);
    capI = std::max(capI, t.part.capI);
  }
  if (fbc_smem_bytes(capV, capH, capI) > FBC_SMEM_LIMIT)
    FB_FAIL(c, FB_E_STATE, "cluster plan: shared-memory layout exceeds 227 KB");
  if (capV != P->capV || capH != P->capH || capI != P->capI)
    for (auto& t : P->topo) t.dirty = true;
  P->capV = capV;
  P->capH = capH;
  P->capI = capI;
  // pass 2: per-stream tables
  std::vector<int4> eplan, cinfo(FBC_MAXC);
  std::vector<int2> pplan;
  std::vector<int32_t> hplan, pad(FBC_MAXC + 1), hpart(FBC_MAXC + 1);
  cudaStream_t st = c->stream;
  for (int s = 0; s < c->S; ++s) {
    ClusterPlan::Topo& t = P->topo[s];
    if (!t.dirty) continue;
    std::fill(cinfo.begin(), cinfo.end(), make_int4(0, 0, 0, 0));
    if (t.V == 0) {  // empty stream: its cluster exits on nV == 0, but keep the tables defined
      std::fill(pad.begin(), pad.end(), 0);
      FB_CUDA(c, cudaMemcpyAsync(P->vpart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, st));
      FB_CUDA(c, cudaMemcpyAsync(P->epart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, st));
      FB_CUDA(c, cudaMemcpyAsync(P->hpart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, st));
      FB_CUDA(c, cudaMemcpyAsync(P->cinfo + (size_t)s * FBC_MAXC, cinfo.data(), sizeof(int4) * FBC_MAXC, cudaMemcpyHostToDevice, st));
      FB_CUDA(c, cudaStreamSynchronize(st));
      t.dirty = false;
      continue;
    }
    const std::vector<int>& vp = t.part.vpart;
    const std::vector<int>& ep = t.part.epart;
    std::vector<int> rk(t.V);
    for (int r = 0; r < C; ++r)
      for (int v = vp[r]; v < vp[r + 1]; ++v) rk[v] = r;
    // halo lists: per rank, the sorted unique remote targets of its owned edges
    hplan.clear();
    std::vector<int> hidx(t.V, -1);  // halo index of vertex j within the rank being processed
    std::vector<std::vector<int2>> push(C);  // push[y] = {local vertex in y, rank<<20 | halo index}
    for (int r = 0; r < C; ++r) {
      hpart[r] = (int)hplan.size();
      std::vector<int> hl;
      for (int e = ep[r]; e < ep[r + 1]; ++e) {
        const int j = t.eij[e].y;
        if (j >= vp[r + 1]) hl.push_back(j);
      }
      std::sort(hl.begin(), hl.end());
      hl.erase(std::unique(hl.begin(), hl.end()), hl.end());
      for (size_t h = 0; h < hl.size(); ++h) {
        hidx[hl[h]] = (int)h;
        const int y = rk[hl[h]];
        push[y].push_back(make_int2(hl[h] - vp[y], (r << 20) | (capV + (int)h)));
        hplan.push_back(hl[h]);
      }
      cinfo[r].x = (int)hl.size();
      // edge plans of rank r (needs hidx of this rank)
      if (r == 0) eplan.assign(t.E, make_int4(0, 0, 0, 0));
      for (int e = ep[r]; e < ep[r + 1]; ++e) {
        const int i = t.eij[e].x, j = t.eij[e].y;
        eplan[e].x = i - vp[r];
        eplan[e].y = (j < vp[r + 1]) ? (j - vp[r]) : (capV + hidx[j]);
      }
    }
    for (int r = C; r <= FBC_MAXC; ++r) hpart[r] = (int)hplan.size();
    // processing order of the vertices: per CTA by descending degree, so the lanes of a warp run
    // slot loops of equal length; slot blocks follow that order and are padded to an odd number
    // of 16 B records, which makes the per-lane stride conflict-free across the 32 banks
    std::vector<int32_t> vorig(t.V), sbase(t.V);
    for (int r = 0; r < C; ++r) {
      std::vector<int> vs;
      for (int v = vp[r]; v < vp[r + 1]; ++v) vs.push_back(v);
      std::stable_sort(vs.begin(), vs.end(), [&](int x, int y) {
        return (t.row[x + 1] - t.row[x]) > (t.row[y + 1] - t.row[y]);
      });
      int base = 0;
      for (size_t k = 0; k < vs.size(); ++k) {
        vorig[vp[r] + k] = vs[k];
        sbase[vs[k]] = base;
        base += (t.row[vs[k] + 1] - t.row[vs[k]]) | 1;
      }
    }
    // slots: position of every incidence within its vertex's block
    for (int v = 0; v < t.V; ++v) {
      const int r = rk[v];
      for (int k = t.row[v]; k < t.row[v + 1]; ++k) {
        const int code = t.inc[k], e = code >> 1, local = sbase[v] + (k - t.row[v]);
        if ((code & 1) == 0) {
          eplan[e].z = local;
        } else {
          eplan[e].w = (r << 20) | local;
          if (rk[t.eij[e].x] != r) cinfo[r].y++;  // incoming remote contribution
        }
      }
    }
    // the first two consumers of a vertex are refreshed by its owner thread (vplan); the rest
    // (a vertex referenced from more than two other CTAs) go to the overflow list (pplan)
    pplan.clear();
    std::vector<int4> vplan(t.V, make_int4(-1, -1, 0, 0));
    for (int v = 0; v < t.V; ++v) {
      vplan[v].z = sbase[v];
      vplan[v].w = sbase[v] + (t.row[v + 1] - t.row[v]);
    }
    for (int r = 0; r < C; ++r) {
      cinfo[r].z = (int)pplan.size();
      for (const int2& e : push[r]) {
        int4& slot = vplan[vp[r] + e.x];
        if (slot.x < 0) slot.x = e.y;
        else if (slot.y < 0) slot.y = e.y;
        else pplan.push_back(e);
      }
      cinfo[r].w = (int)pplan.size();
    }
    // processing order of the edges: per CTA, cut edges first (their remote stores leave early)
    std::vector<int32_t> eorig(t.E);
    std::vector<int4> eperm(t.E);
    std::vector<int4> vperm(t.V);
    for (int r = 0; r < C; ++r) {
      int pos = ep[r];
      for (int pass = 0; pass < 2; ++pass)
        for (int e = ep[r]; e < ep[r + 1]; ++e) {
          const bool cut = ((uint32_t)eplan[e].w >> 20) != (uint32_t)r;
          if (cut == (pass == 0)) eorig[pos++] = e;
        }
    }
    for (int e = 0; e < t.E; ++e) eperm[e] = eplan[eorig[e]];
    for (int v = 0; v < t.V; ++v) vperm[v] = vplan[vorig[v]];
    FB_CUDA(c, cudaMemcpyAsync(P->vplan + (size_t)s * c->maxV, vperm.data(), sizeof(int4) * t.V, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(P->vorig + (size_t)s * c->maxV, vorig.data(), sizeof(int32_t) * t.V, cudaMemcpyHostToDevice, st));
    if (t.E) FB_CUDA(c, cudaMemcpyAsync(P->eorig + (size_t)s * c->maxE, eorig.data(), sizeof(int32_t) * t.E, cudaMemcpyHostToDevice, st));
    eplan.swap(eperm);
    if ((int)pplan.size() > c->maxE || (int)hplan.size() > c->maxE)
      FB_FAIL(c, FB_E_NOMEM, "cluster plan: halo tables exceed capacity");
    if (t.E)
      FB_CUDA(c, cudaMemcpyAsync(P->eplan + (size_t)s * c->maxE, eplan.data(), sizeof(int4) * t.E, cudaMemcpyHostToDevice, st));
    if (!pplan.empty())
      FB_CUDA(c, cudaMemcpyAsync(P->pplan + (size_t)s * c->maxE, pplan.data(), sizeof(int2) * pplan.size(), cudaMemcpyHostToDevice, st));
    if (!hplan.empty())
      FB_CUDA(c, cudaMemcpyAsync(P->hplan + (size_t)s * c->maxE, hplan.data(), sizeof(int32_t) * hplan.size(), cudaMemcpyHostToDevice, st));
    for (int r = 0; r <= FBC_MAXC; ++r) pad[r] = vp[std::min(r, C)];
    FB_CUDA(c, cudaMemcpyAsync(P->vpart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, st));
    for (int r = 0; r <= FBC_MAXC; ++r) pad[r] = ep[std::min(r, C)];
    FB_CUDA(c, cudaMemcpyAsync(P->epart + (size_t)s * (FBC_MAXC + 1), pad.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(P->hpart + (size_t)s * (FBC_MAXC + 1), hpart.data(), sizeof(int32_t) * (FBC_MAXC + 1), cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(P->cinfo + (size_t)s * FBC_MAXC, cinfo.data(), sizeof(int4) * FBC_MAXC, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaStreamSynchronize(st));  // staging vectors are reused per stream
    t.dirty = false;
  }
  return FB_OK;
}

static int solve_cluster(fb_ctx* c, int iters, const fb_nltgv2_params* p, int only = -1) {
  ClusterPlan* P = c->plan;
  int C = 1;
  bool any = false;
  for (auto& t : P->topo)
    if (t.V > 0) {
      fbc_need(t);
      C = std::max(C, t.needC);
      any = true;
    }
  if (!any) return FB_OK;
  // Cluster size: the largest C (<= 16, non-portable above 8) for which all active streams' clusters
  // are co-resident, so few streams spread over more SMs (per-CTA work shrinks, the exchange cost
  // stays: 115 us at C=8 vs 92 us at C=16 for one C2 graph) while many streams keep C small enough
  // to run in one wave.  Graphs are never cut below ~256 vertices per CTA.
  int n_act = 0, maxV = 0;
  for (auto& t : P->topo) {
    if (t.V > 0) ++n_act;
    maxV = std::max(maxV, t.V);
  }
  if (only >= 0) n_act = 1;
  if (c->cluster_min > 1) {
    C = std::max(C, c->cluster_min);  // FB_CLUSTER_MIN: forced lower bound (tuning / tests)
  } else {
    const int cmax = std::min(FBC_MAXC, std::max(C, maxV / 256));
    for (int cand = cmax; cand > C; --cand) {
      if (P->max_active[cand] < 0) {
        cudaLaunchConfig_t q{};
        q.gridDim = dim3((unsigned)(cand * 64));
        q.blockDim = dim3(FBC_THREADS);
        q.dynamicSmemBytes = 96 * 1024;  // registers, not shared memory, limit residency (1 CTA / SM)
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = (unsigned)cand;
        qa[0].val.clusterDim.y = 1;
        qa[0].val.clusterDim.z = 1;
        q.attrs = qa;
        q.numAttrs = 1;
        if (P->smem_set < 96 * 1024) {
          cudaFuncSetAttribute(k_nltgv2_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
          P->smem_set = 96 * 1024;
        }
        if (cand > 8 && !P->nonportable_set) {
          cudaFuncSetAttribute(k_nltgv2_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
          P->nonportable_set = true;
        }
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, k_nltgv2_cluster, &q) != cudaSuccess) {
          cudaGetLastError();
          n = 0;
        }
        P->max_active[cand] = n;
      }
      if (P->max_active[cand] >= n_act) {
        FbcPart probe;
        bool fits = true;
        for (auto& t : P->topo)
          if (t.V > 0 && t.partC != cand && !fbc_partition(t, cand, probe)) fits = false;
        if (fits) {
          C = cand;
          break;
        }
      }
    }
  }
  c->last_cluster = C;
  int rc = fbc_upload_plans(c, C);
  if (rc) return rc;
  const size_t smem = fbc_smem_bytes(P->capV, P->capH, P->capI);
  if (smem != P->smem_set) {
    FB_CUDA(c, cudaFuncSetAttribute(k_nltgv2_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P->smem_set = smem;
  }
  if (C > 8 && !P->nonportable_set) {
    FB_CUDA(c, cudaFuncSetAttribute(k_nltgv2_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    P->nonportable_set = true;
  }
  ClusterArgs a;
  a.g = graph_view(c);
  a.g.only = only;
  a.eplan = P->eplan;
  a.pplan = P->pplan;
  a.vplan = P->vplan;
  a.eorig = P->eorig;
  a.vorig = P->vorig;
  a.hplan = P->hplan;
  a.vpart = P->vpart;
  a.epart = P->epart;
  a.hpart = P->hpart;
  a.cinfo = P->cinfo;
  a.capV = P->capV;
  a.capH = P->capH;
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
  ProfScope ps(c, FB_PROF_SOLVE);  // events hug the launch: host-side planning is not kernel time
  FB_CUDA(c, cudaLaunchKernelEx(&cfg, k_nltgv2_cluster, a, iters, p->step_q, p->step_x, tl, p->theta, p->x_min, p->x_max));
  c->launches++;
  return FB_OK;
}
