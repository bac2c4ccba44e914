// nltgv2_grid.cuh -- grid-resident NLTGV2-L1 solver (variant 3).
//
// The whole batch of graphs stays on chip for ALL iterations of a solve, spread over as many CTAs as
// the device keeps co-resident (cooperative launch, up to two CTAs per SM, no cluster-size or GPC
// limit), so graphs of any practical size run persistently (C4: 20k vertices over 148+ CTAs):
//   * each stream's vertices are cut into `nper` compact parts by recursive coordinate bisection
//     of the pixel positions (balanced by degree); a part is owned by one CTA;
//   * a CTA holds EVERY edge incident to its vertices -- cut edges are held (and computed) by both
//     sides.  Both copies see bit-identical inputs and run the same instruction sequence, so they
//     stay bit-identical; only the copy in the source vertex's CTA is written back.  All K^T q
//     contributions a vertex needs are therefore produced inside its own CTA: the only thing that
//     crosses CTAs is the extragradient point of boundary vertices, ONCE per iteration;
//   * that exchange is a tagged 128-bit mailbox in L2: the owner publishes (xb, w1b, w2b, tag)
//     with one st.relaxed.gpu.b128, readers poll the same 16 bytes with ld.relaxed.gpu.b128 until
//     the tag of the iteration shows up.  Data and flag travel in one single-copy-atomic access:
//     no fence, no flag round trip, no cluster barrier, no grid barrier.  Mailboxes are double
//     buffered by iteration parity (a writer can be at most one iteration ahead of a reader);
//   * per-edge state (q, alpha, beta, dx, dy) and per-vertex state (x, w, z, threshold) live in
//     registers for the whole solve; shared memory holds the extragradient points (own + halo) and
//     one 16 B slot per vertex-edge incidence in the vertex's CSR order, so the summation order
//     (ascending edge id) and hence every bit of the result equals the streaming kernels'.
// HBM is touched once per solve (state in, state out).
#pragma once

#include <algorithm>
#include <numeric>

#include "common.cuh"
#include "nltgv2.cuh"
#include "nltgv2_cluster.cuh"

#define FBG_THREADS 256
#define FBG_EPT 4            // edges per thread (register resident)
#define FBG_VPT 2            // vertices per thread
#define FBG_MAXP 512         // parts per stream (table stride)
#define FBG_MAX_ITERS 16383  // tag space per launch
#define FBG_SMEM_LIMIT (100 * 1024)
#define FBG_SPIN_LIMIT (1u << 21)  // mailbox polls before a reader gives up (watchdog, ~0.3 s)

struct GridPlan {
  int nper = 0;              // parts per stream the device tables are built for
  int2* eplan = nullptr;     // [S*2*maxE] {bi | bj<<16, si | sj<<16}: s_bar / s_slot entry indices
  int32_t* eid = nullptr;    // [S*2*maxE] edge id, bit 31 set on the copy that is NOT written back
  int4* vplan = nullptr;     // [S*maxV] {vertex id, slot begin, slot end, 1 = boundary (published)}
  int32_t* hplan = nullptr;  // [S*2*maxE] halo lists: stream-local vertex ids
  int4* cinfo = nullptr;     // [S*FBG_MAXP*2] {vBeg, nOwn, eBeg, nEdge}, {hBeg, nHalo, nSlot, 0}
  float4* pub = nullptr;     // [2][S*maxV] tagged mailboxes (parity-major)
  int* err = nullptr;        // mapped host flag: set by the watchdog
  uint32_t seq = 0;          // launch counter -> tag base
  int max_blocks = -1;       // co-resident CTAs of k_nltgv2_grid on this device
  size_t smem_set = 0;
  int budget_env = 0;        // FB_GRID_CTAS: total CTA budget override
  int sms = 0;
  int capBar = 0;            // layout of the last prepared launch
  struct Topo {
    std::vector<float2> pos;
    bool dirty = true;
    int planned = 0;   // nper the cached tables below were built for (0 = none)
    bool feasible = false;
    int capBar = 0, capSlot = 0;
    std::vector<int2> eplan;
    std::vector<int32_t> eid, hplan;
    std::vector<int4> vplan, cinfo;
  };
  std::vector<Topo> topo;
};

// ---------------------------------------------------------------------------------- device side
__device__ __forceinline__ uint4 fbg_ld_mailbox(const float4* p) {
  unsigned long long lo, hi;
  asm volatile(
      "{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}"
      : "=l"(lo), "=l"(hi)
      : "l"(p)
      : "memory");
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}
__device__ __forceinline__ void fbg_st_mailbox(float4* p, float a, float b, float c, uint32_t tag) {
  const unsigned long long lo =
      (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
  const unsigned long long hi = (unsigned long long)__float_as_uint(c) | ((unsigned long long)tag << 32);
  asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}" ::"l"(p),
               "l"(lo), "l"(hi)
               : "memory");
}
// Poll one mailbox until it carries `tag`; returns the point.  A reader that never sees its tag
// (a bug, or a grid that is not co-resident) raises the context's error flag instead of hanging.
__device__ __forceinline__ float4 fbg_poll(const float4* p, uint32_t tag, int* err, bool& dead) {
  uint4 v = fbg_ld_mailbox(p);
  if (v.w != tag && !dead) {
    uint32_t spins = 0;
    do {
      v = fbg_ld_mailbox(p);
      if (++spins > FBG_SPIN_LIMIT) {
        dead = true;
        *err = 1;
        break;
      }
    } while (v.w != tag);
  }
  return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), 0.f);
}

struct GridArgs {
  GraphView g;
  const int2* eplan;
  const int32_t* eid;
  const int4* vplan;
  const int32_t* hplan;
  const int4* cinfo;
  float4* pub;
  int* err;
  int nper;
  int capBar;      // s_slot starts capBar records after s_bar
  size_t pstride;  // S*maxV: distance between the two mailbox banks
};

__global__ void __launch_bounds__(FBG_THREADS, 2)
k_nltgv2_grid(GridArgs a, int iters, float sigma, float tau, float tl, float theta, float xmin,
              float xmax, uint32_t tag0) {
  extern __shared__ __align__(16) uint8_t fbg_smem[];
  float4* s_bar = reinterpret_cast<float4*>(fbg_smem);  // [nOwn own | nHalo halo]
  float4* s_slot = s_bar + a.capBar;                    // [nSlot + 1 dummy]
  const GraphView& g = a.g;
  const int tid = threadIdx.x;
  const int s = (g.only >= 0) ? g.only : (int)blockIdx.x / a.nper;
  const int r = (int)blockIdx.x % a.nper;
  if (g.nV[s] == 0) return;
  const int4 c0 = a.cinfo[((size_t)s * FBG_MAXP + r) * 2], c1 = a.cinfo[((size_t)s * FBG_MAXP + r) * 2 + 1];
  const int nOwn = c0.y, nEdge = c0.w, nHalo = c1.y;
  if (nOwn == 0) return;  // an empty part owns nothing and feeds nobody
  const size_t vb = (size_t)s * g.maxV, eb = (size_t)s * g.maxE;
  const int2* epl = a.eplan + 2 * eb + c0.z;
  const int32_t* eidl = a.eid + 2 * eb + c0.z;
  const int4* vpl = a.vplan + vb + c0.x;
  const int32_t* hl = a.hplan + 2 * eb + c1.x;
  const size_t pstride = a.pstride;  // second mailbox bank (odd iterations)
  float4* pub0 = a.pub + vb;

  // ---- register-resident per-edge and per-vertex state ----------------------------------------
  float q1[FBG_EPT], q2[FBG_EPT], q3[FBG_EPT], ea[FBG_EPT], ebt[FBG_EPT], edx[FBG_EPT], edy[FBG_EPT];
  uint32_t e_b[FBG_EPT], e_s[FBG_EPT];  // packed 16-bit entry indices: (bi, bj) and (si, sj)
  int e_id[FBG_EPT];
#pragma unroll
  for (int k = 0; k < FBG_EPT; ++k) {
    const int idx = tid + k * FBG_THREADS;
    e_id[k] = -1;
    // idle lanes run the same arithmetic on zero weights and store nothing (branch-free ILP)
    q1[k] = q2[k] = q3[k] = ea[k] = ebt[k] = edx[k] = edy[k] = 0.f;
    e_b[k] = 0u;
    e_s[k] = (uint32_t)c1.z | ((uint32_t)c1.z << 16);
    if (idx < nEdge) {
      const int2 pl = epl[idx];
      const int id = eidl[idx];
      e_id[k] = id;
      const float4 c = g.ec[eb + (id & 0x7fffffff)];
      const float4 q = g.q4[eb + (id & 0x7fffffff)];
      ea[k] = c.x; ebt[k] = c.y; edx[k] = c.z; edy[k] = c.w;
      q1[k] = q.x; q2[k] = q.y; q3[k] = q.z;
      e_b[k] = (uint32_t)pl.x;
      e_s[k] = (uint32_t)pl.y;
    }
  }
  float vx[FBG_VPT], vw1[FBG_VPT], vw2[FBG_VPT], vz[FBG_VPT], vth[FBG_VPT];
  int vs0[FBG_VPT], vs1[FBG_VPT], v_id[FBG_VPT];  // v_id: vertex id, bit 30 = boundary; -1 = none
#pragma unroll
  for (int k = 0; k < FBG_VPT; ++k) {
    const int idx = tid + k * FBG_THREADS;
    vx[k] = vw1[k] = vw2[k] = vz[k] = vth[k] = 0.f;
    vs0[k] = 0;
    vs1[k] = 0;
    v_id[k] = -1;
    if (idx < nOwn) {
      const int4 pt = vpl[idx];
      const int v = pt.x;
      v_id[k] = v | (pt.w ? 0x40000000 : 0);
      vx[k] = g.x[vb + v]; vw1[k] = g.w1[vb + v]; vw2[k] = g.w2[vb + v];
      vz[k] = g.z[vb + v];
      vth[k] = tl * g.wt[vb + v];
      vs0[k] = pt.y;
      vs1[k] = pt.z;
      s_bar[idx] = g.vbar[vb + v];
    }
  }
  // halo: vertex ids of the first two entries per thread stay in registers; the first value comes
  // straight from global memory (written by earlier kernels of the stream)
  int hv0 = -1, hv1 = -1;
  for (int h = tid; h < nHalo; h += FBG_THREADS) {
    const int hv = hl[h];
    if (h == tid) hv0 = hv;
    else if (h == tid + FBG_THREADS) hv1 = hv;
    s_bar[nOwn + h] = g.vbar[vb + hv];
  }
  bool dead = false;

  for (int it = 0; it < iters; ++it) {
    const bool more = it + 1 < iters;
    // ---- halo refresh: poll the tagged mailboxes of iteration it-1 ---------------------------
    if (it > 0 && hv0 >= 0) {
      const uint32_t tag = tag0 + (uint32_t)(it - 1);
      const float4* pb = pub0 + (size_t)((it - 1) & 1) * pstride;
      s_bar[nOwn + tid] = fbg_poll(pb + hv0, tag, a.err, dead);
      if (hv1 >= 0) {
        s_bar[nOwn + tid + FBG_THREADS] = fbg_poll(pb + hv1, tag, a.err, dead);
        for (int h = tid + 2 * FBG_THREADS; h < nHalo; h += FBG_THREADS)
          s_bar[nOwn + h] = fbg_poll(pb + hl[h], tag, a.err, dead);
      }
    }
    __syncthreads();  // own points (primal of it-1) and halo points visible to the edge threads
    // ---- dual half-step: every edge incident to this CTA's vertices ---------------------------
#pragma unroll
    for (int k = 0; k < FBG_EPT; ++k) {
      const float4 bi = s_bar[e_b[k] & 0xffffu];
      const float4 bj = s_bar[e_b[k] >> 16];
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
      if (e_id[k] != -1) {  // the endpoint owned by another CTA maps to the dummy slot
        s_slot[e_s[k] & 0xffffu] = cs;
        s_slot[e_s[k] >> 16] = ct;
      }
    }
    __syncthreads();  // slots complete
    // ---- primal half-step: slot gather in CSR order, prox, box, extragradient, publish --------
    {
      float gx[FBG_VPT], g1[FBG_VPT], g2[FBG_VPT];
      int dmax = 0;
#pragma unroll
      for (int k = 0; k < FBG_VPT; ++k) {
        gx[k] = g1[k] = g2[k] = 0.f;
        dmax = max(dmax, vs1[k] - vs0[k]);
      }
      for (int j = 0; j < dmax; ++j) {
#pragma unroll
        for (int k = 0; k < FBG_VPT; ++k) {
          if (vs0[k] + j < vs1[k]) {
            const float4 c = s_slot[vs0[k] + j];
            gx[k] += c.x;
            g1[k] += c.y;
            g2[k] += c.z;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < FBG_VPT; ++k) {
        if (v_id[k] >= 0) {
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
          if (more) {
            if (v_id[k] & 0x40000000)
              fbg_st_mailbox(pub0 + (size_t)(it & 1) * pstride + (v_id[k] & 0x3fffffff), nb.x, nb.y, nb.z,
                             tag0 + (uint32_t)it);
            s_bar[tid + k * FBG_THREADS] = nb;
          } else {
            g.vbar[vb + (v_id[k] & 0x3fffffff)] = nb;
          }
        }
      }
    }
  }

  // ---- write back: registers -> global -----------------------------------------------------------
#pragma unroll
  for (int k = 0; k < FBG_EPT; ++k)
    if (e_id[k] >= 0) g.q4[eb + e_id[k]] = make_float4(q1[k], q2[k], q3[k], 0.f);
#pragma unroll
  for (int k = 0; k < FBG_VPT; ++k)
    if (v_id[k] >= 0) {
      const size_t v = vb + (v_id[k] & 0x3fffffff);
      g.x[v] = vx[k]; g.w1[v] = vw1[k]; g.w2[v] = vw2[k];
    }
}

// ---------------------------------------------------------------------------------- host side
static inline size_t fbg_smem_bytes(int capBar, int capSlot) {
  return 16 * ((size_t)capBar + (size_t)capSlot + 1);
}

// Recursive coordinate bisection: ids[lo,hi) -> parts [p0, p0+np), split along the longer extent
// at the weighted position that gives each side its share of the parts.
static void fbg_rcb(const float2* pos, const int* wgt, int* ids, int lo, int hi, int p0, int np, int* part) {
  if (np <= 1 || hi - lo <= 1) {
    for (int k = lo; k < hi; ++k) part[ids[k]] = p0;
    return;
  }
  float x0 = 1e30f, x1 = -1e30f, y0 = 1e30f, y1 = -1e30f;
  long long tot = 0;
  for (int k = lo; k < hi; ++k) {
    const float2 p = pos[ids[k]];
    x0 = std::min(x0, p.x); x1 = std::max(x1, p.x);
    y0 = std::min(y0, p.y); y1 = std::max(y1, p.y);
    tot += wgt[ids[k]];
  }
  const bool ax = (x1 - x0) >= (y1 - y0);
  std::sort(ids + lo, ids + hi, [&](int u, int v) {
    const float cu = ax ? pos[u].x : pos[u].y, cv = ax ? pos[v].x : pos[v].y;
    return cu < cv || (cu == cv && u < v);
  });
  const int nl = np / 2;
  const long long target = tot * nl / np;
  long long acc = 0;
  int m = lo;
  while (m < hi - 1 && acc + wgt[ids[m]] <= target) acc += wgt[ids[m++]];
  if (m == lo) m = lo + 1;
  fbg_rcb(pos, wgt, ids, lo, m, p0, nl, part);
  fbg_rcb(pos, wgt, ids, m, hi, p0 + nl, np - nl, part);
}

// Build the host tables of one stream for `nper` parts.  Returns false when a part exceeds the
// per-CTA register or shared-memory capacity (the caller then tries more parts or another variant).
static bool fbg_build(GridPlan::Topo& g, const ClusterPlan::Topo& t, int nper) {
  const int V = t.V, E = t.E;
  g.planned = nper;
  g.feasible = false;
  g.capBar = g.capSlot = 0;
  g.cinfo.assign((size_t)2 * nper, make_int4(0, 0, 0, 0));
  g.eplan.clear(); g.eid.clear(); g.hplan.clear();
  g.vplan.assign(V, make_int4(0, 0, 0, 0));
  if (V == 0) {
    g.feasible = true;
    return true;
  }
  std::vector<int> deg(V), wgt(V), ids(V), part(V, 0);
  for (int v = 0; v < V; ++v) {
    deg[v] = t.row[v + 1] - t.row[v];
    wgt[v] = 2 + deg[v];
  }
  std::iota(ids.begin(), ids.end(), 0);
  fbg_rcb(g.pos.data(), wgt.data(), ids.data(), 0, V, 0, nper, part.data());
  // boundary vertices; CSR position of every edge at its two endpoints
  std::vector<uint8_t> bnd(V, 0);
  std::vector<int> psrc(E), pdst(E);
  for (int e = 0; e < E; ++e)
    if (part[t.eij[e].x] != part[t.eij[e].y]) bnd[t.eij[e].x] = bnd[t.eij[e].y] = 1;
  for (int v = 0; v < V; ++v)
    for (int k = t.row[v]; k < t.row[v + 1]; ++k) {
      const int code = t.inc[k];
      if (code & 1) pdst[code >> 1] = k - t.row[v];
      else psrc[code >> 1] = k - t.row[v];
    }
  // processing order per part: boundary vertices first (published early), then descending degree
  // (lanes of a warp run slot loops of equal length); slot blocks are padded to an odd number of
  // 16 B records so a warp's gathers spread over all banks
  std::vector<int> cnt(nper + 1, 0);
  for (int v = 0; v < V; ++v) cnt[part[v] + 1]++;
  for (int r = 0; r < nper; ++r) cnt[r + 1] += cnt[r];
  std::vector<int> order(V);
  {
    std::vector<int> fill(cnt.begin(), cnt.begin() + nper);
    for (int v = 0; v < V; ++v) order[fill[part[v]]++] = v;
  }
  std::vector<int> lidx(V), sbase(V), nslot(nper, 0);
  for (int r = 0; r < nper; ++r) {
    std::sort(order.begin() + cnt[r], order.begin() + cnt[r + 1], [&](int u, int v) {
      if (bnd[u] != bnd[v]) return bnd[u] > bnd[v];
      if (deg[u] != deg[v]) return deg[u] > deg[v];
      return u < v;
    });
    int base = 0;
    for (int k = cnt[r]; k < cnt[r + 1]; ++k) {
      const int v = order[k];
      lidx[v] = k - cnt[r];
      sbase[v] = base;
      base += deg[v] | 1;
      g.vplan[k] = make_int4(v, sbase[v], sbase[v] + deg[v], bnd[v]);
    }
    nslot[r] = base;
    const int nOwn = cnt[r + 1] - cnt[r];
    if (nOwn > FBG_VPT * FBG_THREADS || base + 1 > 0xffff) return false;
    g.cinfo[2 * r].x = cnt[r];
    g.cinfo[2 * r].y = nOwn;
    g.cinfo[2 * r + 1].z = base;
    g.capSlot = std::max(g.capSlot, base);
  }
  // edge lists per part: interior edges first, then cut edges (held by both sides)
  std::vector<int> ecnt(nper + 1, 0);
  for (int e = 0; e < E; ++e) {
    const int ri = part[t.eij[e].x], rj = part[t.eij[e].y];
    ecnt[ri + 1]++;
    if (rj != ri) ecnt[rj + 1]++;
  }
  for (int r = 0; r < nper; ++r) ecnt[r + 1] += ecnt[r];
  const int total = ecnt[nper];
  g.eplan.assign(total, make_int2(0, 0));
  g.eid.assign(total, 0);
  std::vector<int> efill(ecnt.begin(), ecnt.begin() + nper);
  std::vector<int> elist(total);
  for (int pass = 0; pass < 2; ++pass)
    for (int e = 0; e < E; ++e) {
      const int ri = part[t.eij[e].x], rj = part[t.eij[e].y];
      if ((ri != rj) != (pass == 1)) continue;
      elist[efill[ri]++] = e;
      if (rj != ri) elist[efill[rj]++] = e | (int)0x80000000;
    }
  std::vector<int> hidx(V, -1), hstamp(V, -1);
  for (int r = 0; r < nper; ++r) {
    const int nOwn = cnt[r + 1] - cnt[r], nE = ecnt[r + 1] - ecnt[r];
    if (nE > FBG_EPT * FBG_THREADS) return false;
    const int hbeg = (int)g.hplan.size();
    int nh = 0;
    for (int k = ecnt[r]; k < ecnt[r + 1]; ++k) {
      const int code = elist[k], e = code & 0x7fffffff;
      const int i = t.eij[e].x, j = t.eij[e].y;
      int b[2], sl[2];
      const int end[2] = {i, j};
      for (int side = 0; side < 2; ++side) {
        const int v = end[side];
        if (part[v] == r) {
          b[side] = lidx[v];
          sl[side] = sbase[v] + (side == 0 ? psrc[e] : pdst[e]);
        } else {
          if (hstamp[v] != r) {
            hstamp[v] = r;
            hidx[v] = nh++;
            g.hplan.push_back(v);
          }
          b[side] = nOwn + hidx[v];
          sl[side] = nslot[r];  // dummy slot
        }
      }
      if (nOwn + nh > 0xffff) return false;
      g.eplan[k] = make_int2(b[0] | (b[1] << 16), sl[0] | (sl[1] << 16));
      g.eid[k] = code;
    }
    g.cinfo[2 * r].z = ecnt[r];
    g.cinfo[2 * r].w = nE;
    g.cinfo[2 * r + 1].x = hbeg;
    g.cinfo[2 * r + 1].y = nh;
    g.capBar = std::max(g.capBar, nOwn + nh);
  }
  if (fbg_smem_bytes(g.capBar, g.capSlot) > FBG_SMEM_LIMIT) return false;
  g.feasible = true;
  return true;
}

static int grid_plan_init(fb_ctx* c) {
  if (c->gplan) return FB_OK;
  GridPlan* P = new GridPlan();
  c->gplan = P;
  P->topo.resize(c->S);
  const size_t S = c->S;
  if (dalloc(&P->eplan, S * 2 * c->maxE) != cudaSuccess || dalloc(&P->eid, S * 2 * c->maxE) != cudaSuccess ||
      dalloc(&P->vplan, S * c->maxV) != cudaSuccess || dalloc(&P->hplan, S * 2 * c->maxE) != cudaSuccess ||
      dalloc(&P->cinfo, S * FBG_MAXP * 2) != cudaSuccess || dalloc(&P->pub, 2 * S * c->maxV) != cudaSuccess ||
      cudaHostAlloc((void**)&P->err, sizeof(int), cudaHostAllocMapped) != cudaSuccess)
    FB_FAIL(c, FB_E_NOMEM, "grid plan allocation failed");
  *P->err = 0;
  FB_CUDA(c, cudaMemsetAsync(P->pub, 0, sizeof(float4) * 2 * S * c->maxV, c->stream));
  if (const char* e = getenv("FB_GRID_CTAS")) P->budget_env = atoi(e);
  return FB_OK;
}

// fb_graph_set hook: remember the vertex positions (the rest of the topology is shared with the
// cluster plan's host copy) and invalidate the stream's tables.
static int grid_plan_set(fb_ctx* c, int s, int V, const float* pos) {
  int rc = grid_plan_init(c);
  if (rc) return rc;
  GridPlan::Topo& g = c->gplan->topo[s];
  g.pos.resize(V);
  for (int v = 0; v < V; ++v) g.pos[v] = make_float2(pos[2 * v], pos[2 * v + 1]);
  g.dirty = true;
  g.planned = 0;
  return FB_OK;
}

static void grid_plan_free(fb_ctx* c) {
  GridPlan* P = c->gplan;
  if (!P) return;
  cudaFree(P->eplan); cudaFree(P->eid); cudaFree(P->vplan); cudaFree(P->hplan); cudaFree(P->cinfo);
  cudaFree(P->pub);
  if (P->err) cudaFreeHost(P->err);
  delete P;
  c->gplan = nullptr;
}

// True when the watchdog of an earlier variant-3 launch fired (checked after stream syncs).
static bool grid_watchdog_fired(const fb_ctx* c) { return c->gplan && c->gplan->err && *c->gplan->err != 0; }

// Pick the parts per stream, (re)build and upload the tables.  Returns FB_OK with *nper_out = 0
// when the batch does not fit the grid-resident solver (the caller falls back to another variant).
static int grid_prepare(fb_ctx* c, int only, int* nper_out, size_t* smem_out) {
  *nper_out = 0;
  if (!c->gplan || !c->plan) return FB_OK;
  GridPlan* P = c->gplan;
  int n_act = 0, maxV = 0, maxE = 0;
  for (int s = 0; s < c->S; ++s) {
    if (only >= 0 && s != only) continue;
    if (c->hV[s] > 0) ++n_act;
    maxV = std::max(maxV, c->hV[s]);
    maxE = std::max(maxE, c->hE[s]);
  }
  if (n_act == 0) return FB_OK;
  if (P->max_blocks < 0) {
    // co-residency bound of the cooperative launch, at the shared-memory ceiling the plans may use
    cudaFuncSetAttribute(k_nltgv2_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, FBG_SMEM_LIMIT);
    P->smem_set = FBG_SMEM_LIMIT;
    int per_sm = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_nltgv2_grid, FBG_THREADS, FBG_SMEM_LIMIT) != cudaSuccess) {
      cudaGetLastError();
      per_sm = 0;
    }
    P->max_blocks = per_sm * sms;
    P->sms = sms;
  }
  int budget = P->max_blocks;
  if (P->budget_env > 0) budget = std::min(budget, P->budget_env);
  const int n_div = only >= 0 ? 1 : c->S;  // CTAs of empty streams are launched too (they exit at once)
  if (budget < n_div) return FB_OK;
  // parts per stream: all co-resident CTAs shared by the active streams, but never below ~96
  // vertices per part (the exchange then dominates) and at least what the capacities need
  int nper = std::min(budget / n_div, FBG_MAXP);
  nper = std::max(1, std::min(nper, std::max(1, maxV / (P->budget_env > 0 ? 16 : 96))));
  const int need = std::max(fb_div_up(maxV, FBG_VPT * FBG_THREADS), fb_div_up(maxE + maxE / 4, FBG_EPT * FBG_THREADS));
  nper = std::max(nper, need);
  bool ok = false;
  for (; nper * n_div <= budget && nper <= FBG_MAXP; nper += std::max(1, nper / 8)) {
    ok = true;
    for (int s = 0; s < c->S && ok; ++s) {
      if ((only >= 0 && s != only) || c->hV[s] == 0) continue;
      GridPlan::Topo& g = P->topo[s];
      if (g.planned != nper) {
        fbg_build(g, c->plan->topo[s], nper);
        g.dirty = true;
      }
      ok = g.feasible;
    }
    if (ok) break;
  }
  if (!ok) return FB_OK;
  int capBar = 1, capSlot = 1;
  cudaStream_t st = c->stream;
  for (int s = 0; s < c->S; ++s) {
    if ((only >= 0 && s != only) || c->hV[s] == 0) continue;
    GridPlan::Topo& g = P->topo[s];
    capBar = std::max(capBar, g.capBar);
    capSlot = std::max(capSlot, g.capSlot);
    if (!g.dirty) continue;
    const size_t vb = (size_t)s * c->maxV, eb2 = (size_t)s * 2 * c->maxE;
    if (g.eplan.size() > 2 * (size_t)c->maxE || g.hplan.size() > 2 * (size_t)c->maxE)
      FB_FAIL(c, FB_E_NOMEM, "grid plan: tables exceed capacity");
    FB_CUDA(c, cudaMemcpyAsync(P->vplan + vb, g.vplan.data(), sizeof(int4) * g.vplan.size(), cudaMemcpyHostToDevice, st));
    if (!g.eplan.empty()) {
      FB_CUDA(c, cudaMemcpyAsync(P->eplan + eb2, g.eplan.data(), sizeof(int2) * g.eplan.size(), cudaMemcpyHostToDevice, st));
      FB_CUDA(c, cudaMemcpyAsync(P->eid + eb2, g.eid.data(), sizeof(int32_t) * g.eid.size(), cudaMemcpyHostToDevice, st));
    }
    if (!g.hplan.empty())
      FB_CUDA(c, cudaMemcpyAsync(P->hplan + eb2, g.hplan.data(), sizeof(int32_t) * g.hplan.size(), cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(P->cinfo + (size_t)s * FBG_MAXP * 2, g.cinfo.data(), sizeof(int4) * g.cinfo.size(), cudaMemcpyHostToDevice, st));
    g.dirty = false;  // host tables stay alive in the Topo: no staging lifetime issue
  }
  P->nper = nper;
  P->capBar = capBar;
  *nper_out = nper;
  *smem_out = fbg_smem_bytes(capBar, capSlot);
  return FB_OK;
}

static int solve_grid(fb_ctx* c, int iters, const fb_nltgv2_params* p, int nper, size_t smem, int only = -1) {
  GridPlan* P = c->gplan;
  if (iters > FBG_MAX_ITERS) FB_FAIL(c, FB_E_ARG, "fb_nltgv2_solve: variant 3 supports at most 16383 iterations per call");
  // tags are unique per launch; when the 32-bit tag space is used up the mailboxes are cleared
  if (P->seq >= 0xffffffffu / (FBG_MAX_ITERS + 1) - 1) {
    FB_CUDA(c, cudaMemsetAsync(P->pub, 0, sizeof(float4) * 2 * (size_t)c->S * c->maxV, c->stream));
    P->seq = 0;
  }
  P->seq++;
  const uint32_t tag0 = P->seq * (uint32_t)(FBG_MAX_ITERS + 1);
  GridArgs a;
  a.g = graph_view(c);
  a.g.only = only;
  a.eplan = P->eplan; a.eid = P->eid; a.vplan = P->vplan; a.hplan = P->hplan; a.cinfo = P->cinfo;
  a.pub = P->pub;
  a.err = P->err;
  a.nper = nper;
  a.capBar = P->capBar;
  a.pstride = (size_t)c->S * c->maxV;
  c->last_cluster = nper;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((only >= 0 ? 1 : c->S) * nper));
  cfg.blockDim = dim3(FBG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the mailbox readers spin
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float tl = p->step_x * p->data_factor;
  ProfScope ps(c, FB_PROF_SOLVE);
  FB_CUDA(c, cudaLaunchKernelEx(&cfg, k_nltgv2_grid, a, iters, p->step_q, p->step_x, tl, p->theta, p->x_min,
                                p->x_max, tag0));
  c->launches++;
  return FB_OK;
}

// ---------------------------------------------------------------------------------- plan verifier
// Host-only structural check of the variant-3 tables for one graph (no device needed): used by the
// CPU test-suite to validate the partitioner against the invariants the kernel relies on.
// stats[8] = {max own vertices, max edges per part, max halo, duplicated (cut) edges, max slots,
//             shared memory bytes, boundary vertices, parts}.
static int fbg_verify(int V, int E, const float* pos, const int32_t* ij, int nper, int32_t* stats, std::string& why) {
  ClusterPlan::Topo t;
  t.V = V;
  t.E = E;
  t.eij.resize(E);
  t.row.assign(V + 1, 0);
  t.inc.resize(2 * (size_t)E);
  for (int e = 0; e < E; ++e) {
    t.eij[e] = make_int2(ij[2 * e], ij[2 * e + 1]);
    t.row[ij[2 * e] + 1]++;
    t.row[ij[2 * e + 1] + 1]++;
  }
  for (int v = 0; v < V; ++v) t.row[v + 1] += t.row[v];
  {
    std::vector<int32_t> fill(t.row.begin(), t.row.begin() + V);
    for (int e = 0; e < E; ++e) {
      t.inc[fill[t.eij[e].x]++] = (e << 1);
      t.inc[fill[t.eij[e].y]++] = (e << 1) | 1;
    }
  }
  GridPlan::Topo g;
  g.pos.resize(V);
  for (int v = 0; v < V; ++v) g.pos[v] = make_float2(pos[2 * v], pos[2 * v + 1]);
  if (!fbg_build(g, t, nper)) {
    why = "partition infeasible for this part count";
    return 1;
  }
  std::vector<int> owner(V, -1), lidx(V, -1), written(E, 0);
  int maxOwn = 0, maxEdge = 0, maxHalo = 0, dup = 0, maxSlot = 0, nb = 0;
  for (int r = 0; r < nper; ++r) {
    const int4 c0 = g.cinfo[2 * r], c1 = g.cinfo[2 * r + 1];
    for (int k = 0; k < c0.y; ++k) {
      const int4 pv = g.vplan[c0.x + k];
      if (pv.x < 0 || pv.x >= V || owner[pv.x] != -1) { why = "vertex owned twice or out of range"; return 2; }
      owner[pv.x] = r;
      lidx[pv.x] = k;
      if (pv.z - pv.y != t.row[pv.x + 1] - t.row[pv.x]) { why = "slot block size != degree"; return 3; }
      if (pv.z > c1.z) { why = "slot block beyond the part's slot count"; return 3; }
      nb += pv.w ? 1 : 0;
    }
    maxOwn = std::max(maxOwn, c0.y);
    maxEdge = std::max(maxEdge, c0.w);
    maxHalo = std::max(maxHalo, c1.y);
    maxSlot = std::max(maxSlot, c1.z);
  }
  for (int v = 0; v < V; ++v)
    if (owner[v] < 0) { why = "vertex without owner"; return 2; }
  for (int r = 0; r < nper; ++r) {
    const int4 c0 = g.cinfo[2 * r], c1 = g.cinfo[2 * r + 1];
    std::vector<int> hit(c1.z + 1, 0);
    for (int k = 0; k < c0.w; ++k) {
      const int code = g.eid[c0.z + k], e = code & 0x7fffffff;
      if (e >= E) { why = "edge id out of range"; return 4; }
      const int2 pl = g.eplan[c0.z + k];
      const int end[2] = {t.eij[e].x, t.eij[e].y};
      const int b[2] = {pl.x & 0xffff, (int)((uint32_t)pl.x >> 16)};
      const int sl[2] = {pl.y & 0xffff, (int)((uint32_t)pl.y >> 16)};
      bool any_own = false;
      for (int side = 0; side < 2; ++side) {
        const int v = end[side];
        if (owner[v] == r) {
          any_own = true;
          if (b[side] != lidx[v]) { why = "own endpoint index mismatch"; return 5; }
          // slot = block base + CSR position of this incidence
          const int4 pv = g.vplan[c0.x + lidx[v]];
          int posn = -1;
          for (int q = t.row[v]; q < t.row[v + 1]; ++q)
            if (t.inc[q] == ((e << 1) | side)) posn = q - t.row[v];
          if (posn < 0 || sl[side] != pv.y + posn) { why = "slot != base + CSR position"; return 6; }
          hit[sl[side]]++;
        } else {
          const int h = b[side] - c0.y;
          if (h < 0 || h >= c1.y || g.hplan[c1.x + h] != v) { why = "halo index mismatch"; return 7; }
          if (sl[side] != c1.z) { why = "remote endpoint must map to the dummy slot"; return 8; }
          const int4 pv = g.vplan[g.cinfo[2 * owner[v]].x + lidx[v]];
          if (!pv.w) { why = "halo vertex not published by its owner"; return 9; }
        }
      }
      if (!any_own) { why = "edge without an own endpoint"; return 10; }
      const bool wb = code >= 0;
      if (wb != (owner[end[0]] == r)) { why = "write-back copy must live with the source vertex"; return 11; }
      if (wb) written[e]++;
      else ++dup;
    }
    for (int k = 0; k < c0.y; ++k) {
      const int4 pv = g.vplan[c0.x + k];
      for (int q = pv.y; q < pv.z; ++q)
        if (hit[q] != 1) { why = "slot not written exactly once"; return 12; }
    }
  }
  for (int e = 0; e < E; ++e)
    if (written[e] != 1) { why = "edge not written back exactly once"; return 13; }
  if (stats) {
    stats[0] = maxOwn; stats[1] = maxEdge; stats[2] = maxHalo; stats[3] = dup; stats[4] = maxSlot;
    stats[5] = (int32_t)fbg_smem_bytes(g.capBar, g.capSlot); stats[6] = nb; stats[7] = nper;
  }
  return 0;
}
