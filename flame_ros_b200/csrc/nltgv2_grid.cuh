// nltgv2_grid.cuh -- resident NLTGV2-L1 solver with ONE halo exchange per iteration (variant 3).
//
// The batch of graphs stays on chip for ALL iterations of a solve:
//   * each stream's vertices are cut into `nper` compact parts by recursive coordinate bisection
//     of the pixel positions (balanced by degree); a part is owned by one CTA;
//   * a CTA holds EVERY edge incident to its vertices -- cut edges are held (and computed) by both
//     sides.  Both copies see bit-identical inputs and run the same instruction sequence, so they
//     stay bit-identical; only the copy in the source vertex's CTA is written back.  All K^T q
//     contributions a vertex needs are therefore produced inside its own CTA: the only thing that
//     crosses CTAs is the extragradient point of boundary vertices, ONCE per iteration
//     (variant 2 exchanges twice: contributions one way, refreshed halo points back);
//   * a thread owns ONE vertex and that vertex's out-edges (edges (v,w), v < w: up to FBG_FAST of
//     them in register rows).  Its vertex state (x, w, z, threshold, extragradient point) and its
//     edges' state (q, alpha, beta, dx, dy) live in registers for the whole solve, so the dual
//     update of an out-edge reads only the TARGET's point from shared memory and writes only the
//     target's K^T q contribution to a slot; the source's contribution is re-derived from q in
//     registers.  Per vertex and iteration that is ~160 B of shared-memory traffic instead of
//     ~300 B for an edge-parallel layout -- shared-memory bandwidth bounds this kernel, not HBM;
//   * a vertex sums its contributions in CSR order (ascending edge id): edges are sorted by (i,j)
//     with i < j, so that order is [in-edges, from slots] [out-edges, from registers]
//     [out-edges beyond the register rows, from slots] -- every bit of the result equals the
//     streaming kernels';
//   * edges whose source is not a thread's own vertex (cut edges seen from the target's CTA,
//     out-edges beyond FBG_FAST) are "generic": one per thread, both points through shared memory.
// Two transports for the halo, same kernel body (template parameter):
//   CLUSTER  one thread-block cluster (<= 16 CTAs x 512 threads) per stream: the owner thread pushes
//            the new point straight into the consumers' shared memory with
//            st.async...mbarrier::complete_tx::bytes; a CTA waits on its own mbarrier (one per
//            parity) for exactly the bytes it consumes.  ~215 cycles per hand-over.
//   L2       any number of co-resident CTAs (cooperative launch, 2 x 256 threads per SM) for
//            graphs that need more than 16 CTAs (C4: 20k vertices): the owner publishes
//            (xb, w1b, w2b, tag) with one st.relaxed.gpu.b128, readers poll the same 16 bytes with
//            ld.relaxed.gpu.b128 until the tag of the iteration shows up.  Data and flag travel in
//            one single-copy-atomic access: no fence, no flag round trip, no grid barrier.
//            Measured floor ~1700 cycles per lockstep iteration (profiles/r1_mailbox_latency.md).
// HBM is touched once per solve (state in, state out).
#pragma once

#include <algorithm>
#include <numeric>
#include <set>

#include "common.cuh"
#include "nltgv2.cuh"
#include "nltgv2_cluster.cuh"

#define FBG_FAST 4             // register rows: out-edges of the thread's own vertex
#define FBG_THREADS_L2 256     // L2 transport: two CTAs per SM
#define FBG_THREADS_CL 512     // cluster transport: largest CTA (one per SM); see FBG_CL for the others
#define FBG_NCL 3
#define FBG_MAXP 512           // parts per stream (table stride)
#define FBG_MAXC 16            // largest cluster
#define FBG_MAX_ITERS 16383    // tag space per launch (L2 transport)
#define FBG_SMEM_LIMIT_L2 (100 * 1024)
#define FBG_SMEM_LIMIT_CL (200 * 1024)
#define FBG_SPIN_LIMIT (1u << 21)  // mailbox polls before a reader gives up (watchdog, ~0.3 s)

struct GridPlan {
  uint32_t* fplan = nullptr; // [S*FBG_FAST*maxV] fast rows: bj | sj<<16 (row-major: row k, vertex slot)
  int32_t* feid = nullptr;   // [S*FBG_FAST*maxV] fast rows: edge id, -1 = idle
  int2* gplan = nullptr;     // [S*2*maxE] generic edges {bi | bj<<16, si | sj<<16}: s_bar / s_slot entry indices
  int32_t* geid = nullptr;   // [S*2*maxE] generic edges: edge id, bit 31 set on a copy that is NOT written back
  int4* vplan = nullptr;     // [S*maxV] {vertex id, n_target | n_overflow<<8 | entry<<16, push begin | end<<16, 0}
  int32_t* hplan = nullptr;  // [S*2*maxE] halo lists: stream-local vertex ids
  int2* pplan = nullptr;     // [S*2*maxE] push lists: {consumer part, entry index in its s_bar}
  int4* cinfo = nullptr;     // [S*FBG_MAXP*3] {vBeg, nOwn, gBeg, nGen}, {hBeg, nHalo, nSlot, in-edge rows | stride<<8}, {pBeg, nPush, first halo-reading thread, 0}
  float4* pub = nullptr;     // [2][S*maxV] tagged mailboxes (parity-major), L2 transport
  int* err = nullptr;        // mapped host flag: set by the watchdog
  uint32_t seq = 0;          // launch counter -> tag base
  int max_blocks = -1;       // co-resident CTAs of the L2-transport kernel on this device
  int max_clusters[FBG_NCL][FBG_MAXC + 1];  // co-resident clusters per CTA shape and cluster size (-1 = not queried)
  size_t smem_set[1 + FBG_NCL] = {0, 0, 0, 0};  // [0] L2 kernel, [1 + k] cluster kernel of shape k
  int cl_cfg = 0;            // CTA shape (index into FBG_CL) of the last prepared cluster launch
  int threads_env = 0;       // FB_GRID_THREADS: forced CTA size of the cluster transport
  int budget_env = 0;        // FB_GRID_CTAS: CTA budget of the L2 transport
  int cluster_env = 0;       // FB_GRID_CLUSTER: forced cluster size
  int mode_env = 0;          // FB_GRID_MODE: 1 = cluster only, 2 = L2 only
  // layout of the last prepared launch
  int nper = 0, capBar = 0, capSlot = 0, capPush = 0;
  bool cluster = false;
  // decision cache: valid while no topology changed (version) and the same streams are solved
  uint64_t version = 1, dec_version = 0;
  int dec_only = -2, dec_nper = 0;
  bool dec_cluster = false;
  int dec_cfg = 0;
  size_t dec_smem = 0;
  struct Topo {
    std::vector<float2> pos;
    bool dirty = true;
    int planned = 0;          // nper the cached tables below were built for (0 = none)
    int planned_threads = 0;  // capacity class they were checked against
    bool feasible = false;
    int capBar = 0, capSlot = 0, capPush = 0;
    std::vector<uint32_t> fplan;   // [FBG_FAST][V]
    std::vector<int32_t> feid;     // [FBG_FAST][V]
    std::vector<int2> gplan, pplan;
    std::vector<int32_t> geid, hplan;
    std::vector<int4> vplan, cinfo;
  };
  std::vector<Topo> topo;
  GridPlan() { std::fill(&max_clusters[0][0], &max_clusters[0][0] + FBG_NCL * (FBG_MAXC + 1), -1); }
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
// Predicated 16 B shared-memory store (no branch): rows without a local target store nothing.
__device__ __forceinline__ void fbg_sts_if(uint32_t addr, float4 v, uint32_t pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"r"(addr),
      "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(pred)
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
  const uint32_t* fplan;
  const int32_t* feid;
  const int2* gplan;
  const int32_t* geid;
  const int4* vplan;
  const int32_t* hplan;
  const int2* pplan;
  const int4* cinfo;
  float4* pub;
  int* err;
  int nper;
  int capBar;      // records per s_bar bank
  int capSlot;     // slot records (+1 dummy)
  size_t pstride;  // S*maxV: distance between the two mailbox banks
  long long* trace;  // FBG_TRACE builds: [CTA][iteration][6] clock64 stamps of thread 0 (else unused)
};

struct FbgFast {  // register-resident out-edges of the thread's own vertex (source = own vertex)
  float q1[FBG_FAST], q2[FBG_FAST], q3[FBG_FAST], a[FBG_FAST], b[FBG_FAST], dx[FBG_FAST], dy[FBG_FAST];
  uint32_t idx[FBG_FAST];  // target's s_bar entry | target's slot << 16 (dummy slot when the target is remote)
  uint32_t store;          // bit k: row k has a local target slot; bits 8 / 9: the generic edge's source / target slot
};
struct FbgGen {  // register-resident generic edge (both endpoints through shared memory)
  float q1, q2, q3, a, b, dx, dy;
  uint32_t bar, slot;  // (bi, bj) and (si, sj), 16 bits each
};

// Dual half-step of one thread: the first R out-edges of its vertex (register rows) and, when GEN,
// its generic edge.  For an out-edge the source's extragradient point is in registers; only the
// target's is read from shared memory, only the target's contribution is written to a slot (the
// source's is re-derived from q in the primal half-step).  All loads are issued before the first
// dependent instruction and the stores are predicated, not branched around (idle rows and remote
// targets store nothing), so the rows are independent branch-free instruction streams that overlap.
template <int R, bool GEN>
__device__ __forceinline__ void fbg_dual(FbgFast& F, FbgGen& G, float xb, float w1b, float w2b, uint32_t bar_rd,
                                         uint32_t slot_base, float sigma) {
  float4 bj[R > 0 ? R : 1], gi, gj;
#pragma unroll
  for (int k = 0; k < R; ++k) bj[k] = fbc_lds(bar_rd + ((F.idx[k] & 0xffffu) << 4));
  if (GEN) {
    gi = fbc_lds(bar_rd + ((G.bar & 0xffffu) << 4));
    gj = fbc_lds(bar_rd + ((G.bar >> 16) << 4));
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    float t = xb - bj[k].x;
    t = fmaf(-F.dx[k], w1b, t);
    t = fmaf(-F.dy[k], w2b, t);
    const float k1 = F.a[k] * t;
    const float k2 = F.b[k] * (w1b - bj[k].y);
    const float k3 = F.b[k] * (w2b - bj[k].z);
    F.q1[k] = fb_clamp1(fmaf(sigma, k1, F.q1[k]));
    F.q2[k] = fb_clamp1(fmaf(sigma, k2, F.q2[k]));
    F.q3[k] = fb_clamp1(fmaf(sigma, k3, F.q3[k]));
  }
  if (GEN) {  // the source is remote (cut edge seen from the target's CTA) or beyond the register rows
    float t = gi.x - gj.x;
    t = fmaf(-G.dx, gi.y, t);
    t = fmaf(-G.dy, gi.z, t);
    const float k1 = G.a * t;
    const float k2 = G.b * (gi.y - gj.y);
    const float k3 = G.b * (gi.z - gj.z);
    G.q1 = fb_clamp1(fmaf(sigma, k1, G.q1));
    G.q2 = fb_clamp1(fmaf(sigma, k2, G.q2));
    G.q3 = fb_clamp1(fmaf(sigma, k3, G.q3));
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const float a1 = F.a[k] * F.q1[k];
    fbg_sts_if(slot_base + ((F.idx[k] >> 16) << 4), make_float4(-a1, -(F.b[k] * F.q2[k]), -(F.b[k] * F.q3[k]), 0.f),
               F.store & (1u << k));
  }
  if (GEN) {
    const float a1 = G.a * G.q1;
    fbg_sts_if(slot_base + ((G.slot & 0xffffu) << 4),
               make_float4(a1, fmaf(G.b, G.q2, -(G.dx * a1)), fmaf(G.b, G.q3, -(G.dy * a1)), 0.f), F.store & 0x100u);
    fbg_sts_if(slot_base + ((G.slot >> 16) << 4), make_float4(-a1, -(G.b * G.q2), -(G.b * G.q3), 0.f), F.store & 0x200u);
  }
}
template <bool GEN>
__device__ __forceinline__ void fbg_dual_rows(int rows, FbgFast& F, FbgGen& G, float xb, float w1b, float w2b,
                                              uint32_t bar_rd, uint32_t slot_base, float sigma) {
  switch (rows) {  // warp-uniform
    case 0: fbg_dual<0, GEN>(F, G, xb, w1b, w2b, bar_rd, slot_base, sigma); break;
    case 1: fbg_dual<1, GEN>(F, G, xb, w1b, w2b, bar_rd, slot_base, sigma); break;
    case 2: fbg_dual<2, GEN>(F, G, xb, w1b, w2b, bar_rd, slot_base, sigma); break;
    case 3: fbg_dual<3, GEN>(F, G, xb, w1b, w2b, bar_rd, slot_base, sigma); break;
    default: fbg_dual<FBG_FAST, GEN>(F, G, xb, w1b, w2b, bar_rd, slot_base, sigma); break;
  }
}

template <int THREADS, int MINB, bool CLUSTER>
__global__ void __launch_bounds__(THREADS, MINB)
k_nltgv2_grid(GridArgs a, int iters, float sigma, float tau, float tl, float theta, float xmin,
              float xmax, uint32_t tag0) {
  extern __shared__ __align__(16) uint8_t fbg_smem[];
  float4* s_bar = reinterpret_cast<float4*>(fbg_smem);  // 2 banks x [nEnt own entries | nHalo halo]
  float4* s_slot = s_bar + 2 * a.capBar;                // [nSlot + 1 dummy]
  uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_slot + a.capSlot + 1);  // [2] one per parity
  uint2* s_push = reinterpret_cast<uint2*>(s_mbar + 2);  // {remote s_bar address (bank 0), remote mbarrier 0}
  const GraphView& g = a.g;
  const int tid = threadIdx.x;
  const int s = (g.only >= 0) ? g.only : (int)blockIdx.x / a.nper;
  const int r = CLUSTER ? (int)fbc_cluster_ctarank() : (int)blockIdx.x % a.nper;
  if (g.nV[s] == 0) return;  // uniform over the stream's CTAs
  const int4* ci = a.cinfo + ((size_t)s * FBG_MAXP + r) * 3;
  const int4 c0 = ci[0], c1 = ci[1], c2 = ci[2];
  const int nOwn = c0.y, nGen = c0.w, nHalo = c1.y, nPush = c2.y, haloLo = c2.z;
  if (nOwn == 0) {  // an empty part owns nothing and feeds nobody
    if (CLUSTER) {
      fbc_cluster_sync();
      fbc_cluster_sync();
    }
    return;
  }
  const size_t vb = (size_t)s * g.maxV, eb = (size_t)s * g.maxE;
  const int32_t* hl = a.hplan + 2 * eb + c1.x;
  float4* pub0 = a.pub + vb;
  const uint32_t bar_base = fbc_smem_u32(s_bar), slot_base = fbc_smem_u32(s_slot);
  const uint32_t bank_bytes = 16u * (uint32_t)a.capBar;
  const uint32_t mb0 = fbc_smem_u32(s_mbar);
  const uint32_t haloBytes = 16u * (uint32_t)nHalo;
  const uint32_t dummy = (uint32_t)c1.z;
  const int rowsIn = c1.w & 0xff, stride = c1.w >> 8, nEnt = stride - 1;  // slot rows of in-edges, row stride, own entries

  // ---- register-resident state: the thread's vertex, its out-edges, one generic edge ------------
  float vx = 0.f, vw1 = 0.f, vw2 = 0.f, vz = 0.f, vth = 0.f, xb = 0.f, w1b = 0.f, w2b = 0.f;
  int v_id = -1;             // vertex id, bit 30 = boundary; -1 = none
  uint32_t v_sl = 0u;        // n_target | n_overflow << 8 | entry << 16 (slot rows and s_bar entry of this vertex)
  uint32_t v_push = 0u;      // push list range begin | end << 16 (cluster transport)
  uint32_t fvalid = 0u;      // bit k: fast row k holds an edge
  FbgFast F;
#pragma unroll
  for (int k = 0; k < FBG_FAST; ++k) {
    F.q1[k] = F.q2[k] = F.q3[k] = F.a[k] = F.b[k] = F.dx[k] = F.dy[k] = 0.f;
    F.idx[k] = dummy << 16;
  }
  F.store = 0u;
  // The part's slice of the vertex table (vertex id, CSR slot rows and s_bar entry, push range: 16 B per
  // vertex, contiguous) comes in as ONE TMA bulk copy into the slot area, which is idle until the first
  // dual half-step; completion is signalled on an mbarrier.  (Tiny graphs whose slot area is smaller
  // than the slice read the table directly.)
  __shared__ __align__(8) uint64_t s_tma;
  const bool tma = nOwn <= a.capSlot;  // uniform over the CTA
  if (tma) {
    const uint32_t mbt = fbc_smem_u32(&s_tma);
    if (tid == 0) {
      fbc_mbar_init(mbt, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fbc_mbar_expect(mbt, 16u * (uint32_t)nOwn);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot_base),
                   "l"(a.vplan + vb + c0.x), "r"(16u * (uint32_t)nOwn), "r"(mbt)
                   : "memory");
    }
    __syncthreads();  // the barrier is initialised before anyone waits on it
    fbc_mbar_wait(mbt, 0);
  }
  if (tid < nOwn) {
    const int4 pt = tma ? reinterpret_cast<const int4*>(s_slot)[tid] : a.vplan[vb + c0.x + tid];
    const int v = pt.x;
    v_sl = (uint32_t)pt.y;
    v_push = (uint32_t)pt.z;
    v_id = v | (((uint32_t)pt.z >> 16) != ((uint32_t)pt.z & 0xffffu) ? 0x40000000 : 0);
    vx = g.x[vb + v]; vw1 = g.w1[vb + v]; vw2 = g.w2[vb + v];
    vz = g.z[vb + v];
    vth = tl * g.wt[vb + v];
    const float4 b0 = g.vbar[vb + v];
    xb = b0.x; w1b = b0.y; w2b = b0.z;
    s_bar[v_sl >> 16] = b0;  // bank 0: the points iteration 0 reads
#pragma unroll
    for (int k = 0; k < FBG_FAST; ++k) {
      const size_t fi = ((size_t)s * FBG_FAST + k) * g.maxV + c0.x + tid;
      const int id = a.feid[fi];
      if (id >= 0) {
        fvalid |= 1u << k;
        F.idx[k] = a.fplan[fi];
        if ((F.idx[k] >> 16) != dummy) F.store |= 1u << k;
        const float4 c = g.ec[eb + id];
        const float4 q = g.q4[eb + id];
        F.a[k] = c.x; F.b[k] = c.y; F.dx[k] = c.z; F.dy[k] = c.w;
        F.q1[k] = q.x; F.q2[k] = q.y; F.q3[k] = q.z;
      }
    }
  }
  FbgGen G;
  G.q1 = G.q2 = G.q3 = G.a = G.b = G.dx = G.dy = 0.f;
  G.bar = 0u;
  G.slot = dummy | (dummy << 16);
  int g_id = -1;
  const int gi = THREADS - 1 - tid;  // generic edges sit at the back of the CTA (see warpH below)
  if (gi < nGen) {
    const int2 pl = a.gplan[2 * eb + c0.z + gi];
    g_id = a.geid[2 * eb + c0.z + gi];
    const float4 c = g.ec[eb + (g_id & 0x7fffffff)];
    const float4 q = g.q4[eb + (g_id & 0x7fffffff)];
    G.a = c.x; G.b = c.y; G.dx = c.z; G.dy = c.w;
    G.q1 = q.x; G.q2 = q.y; G.q3 = q.z;
    G.bar = (uint32_t)pl.x;
    G.slot = (uint32_t)pl.y;
    if ((G.slot & 0xffffu) != dummy) F.store |= 0x100u;
    if ((G.slot >> 16) != dummy) F.store |= 0x200u;
  }
  // halo: the first value comes straight from global memory (written by earlier kernels of the
  // stream); L2 transport keeps the vertex ids of a thread's first two entries in registers
  int hv0 = -1, hv1 = -1;
  for (int h = tid; h < nHalo; h += THREADS) {
    const int hv = hl[h];
    if (h == tid) hv0 = hv;
    else if (h == tid + THREADS) hv1 = hv;
    s_bar[nEnt + h] = g.vbar[vb + hv];
  }
  if (CLUSTER) {
    if (tid == 0) {
      fbc_mbar_init(mb0, 1);
      fbc_mbar_init(mb0 + 8, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      // points of iteration it-1 land in bank it&1 and complete mbarrier it&1
      if (nHalo && iters > 1) fbc_mbar_expect(mb0 + 8, haloBytes);
      if (nHalo && iters > 2) fbc_mbar_expect(mb0, haloBytes);
    }
    const int2* ppl = a.pplan + 2 * eb + c2.x;
    for (int i = tid; i < nPush; i += THREADS) {
      const int2 e = ppl[i];
      s_push[i] = make_uint2(fbc_mapa(bar_base + 16u * (uint32_t)e.y, (uint32_t)e.x), fbc_mapa(mb0, (uint32_t)e.x));
    }
    __syncthreads();
    fbc_cluster_sync();  // every CTA's barriers are initialised before any remote store can target them
  }
  bool dead = false;
  // warp-uniform work extents: fast rows any lane of the warp uses, generic row, vertex row
  const int rowsF = __reduce_max_sync(0xffffffffu, __popc(fvalid));
  const bool warpG = (THREADS - 1 - (tid | 31)) < nGen, warpV = (tid & ~31) < nOwn;
  // only the LAST warps (boundary vertices, generic edges) ever read a halo entry: the others start the
  // dual half-step without waiting for the hand-over.  The warps on the exchange's critical path are
  // the high-numbered ones on purpose: the SMSP arbiter issues the highest warp id first.
  const bool warpH = (tid | 31) >= haloLo;
  // slots are slot-major: record (row p, entry n) at p * stride + n.  Entries are a permutation of the
  // thread index inside each aligned group of 8 and the stride is odd, so a warp's gather of one
  // row is conflict-free; rows [0, nT) hold the in-edges' contributions, rows [rowsIn, rowsIn + nO)
  // those of out-edges beyond the register rows (rare)
  const int nT = (int)(v_sl & 0xffu), nO = (int)((v_sl >> 8) & 0xffu), ent = (int)(v_sl >> 16);
  const int rowsT = __reduce_max_sync(0xffffffffu, nT), rowsO = __reduce_max_sync(0xffffffffu, nO);

// Phase trace (scripts/grid_trace.py; build with FB_NVCC_EXTRA=-DFBG_TRACE): thread 0 of every CTA
// stamps clock64 at the phase boundaries of the first 64 iterations.  Compiled out otherwise.
#ifdef FBG_TRACE
#define FBG_STAMP(k) if (a.trace && tid == 0 && it < 64) a.trace[((size_t)blockIdx.x * 64 + it) * 6 + (k)] = clock64();
#else
#define FBG_STAMP(k)
#endif
  for (int it = 0; it < iters; ++it) {
    const bool more = it + 1 < iters;
    FBG_STAMP(0)
    const uint32_t rd_off = (it & 1) ? bank_bytes : 0u, wr_off = bank_bytes - rd_off;
    // ---- halo of iteration it-1 -----------------------------------------------------------------
    if (!CLUSTER && it > 0 && hv0 >= 0) {
      const uint32_t tag = tag0 + (uint32_t)(it - 1);
      const float4* pb = pub0 + (size_t)((it - 1) & 1) * a.pstride;
      float4* hb = s_bar + ((it & 1) ? a.capBar : 0) + nEnt;
      hb[tid] = fbg_poll(pb + hv0, tag, a.err, dead);
      if (hv1 >= 0) {
        hb[tid + THREADS] = fbg_poll(pb + hv1, tag, a.err, dead);
        for (int h = tid + 2 * THREADS; h < nHalo; h += THREADS) hb[h] = fbg_poll(pb + hl[h], tag, a.err, dead);
      }
    }
    FBG_STAMP(1)
    __syncthreads();  // own points (primal of it-1) and polled halo points visible to the edge threads
    FBG_STAMP(2)
    if (CLUSTER && it > 0 && nHalo && warpH) {
      // pushed halo points: only the warps that read them wait; the barrier of this parity is re-armed
      // for iteration it+2 by one of the waiting threads once it has seen the phase complete (waits are by parity, so a
      // slower warp still finds this phase completed)
      fbc_mbar_wait(mb0 + 8u * (uint32_t)(it & 1), (uint32_t)(((it + 1) >> 1) - 1) & 1u);
      if (tid == haloLo && it + 2 < iters) fbc_mbar_expect(mb0 + 8u * (uint32_t)(it & 1), haloBytes);
    }
    // ---- dual half-step ---------------------------------------------------------------------------
    if (warpG) fbg_dual_rows<true>(rowsF, F, G, xb, w1b, w2b, bar_base + rd_off, slot_base, sigma);
    else fbg_dual_rows<false>(rowsF, F, G, xb, w1b, w2b, bar_base + rd_off, slot_base, sigma);
    FBG_STAMP(3)
    __syncthreads();  // slots complete
    FBG_STAMP(4)
    // ---- primal half-step: CSR order = target-role slots, own out-edges, overflow slots -----------
    if (warpV) {
      float gx = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll 4
      for (int j = 0; j < rowsT; ++j) {  // warp-uniform trip count, per-lane predicate
        if (j < nT) {
          const float4 c = s_slot[j * stride + ent];
          gx += c.x;
          g1 += c.y;
          g2 += c.z;
        }
      }
#pragma unroll
      for (int k = 0; k < FBG_FAST; ++k) {
        if (fvalid & (1u << k)) {  // the source's contribution, same expression the slot would have held
          const float a1 = F.a[k] * F.q1[k];
          gx += a1;
          g1 += fmaf(F.b[k], F.q2[k], -(F.dx[k] * a1));
          g2 += fmaf(F.b[k], F.q3[k], -(F.dy[k] * a1));
        }
      }
#pragma unroll 1
      for (int o = 0; o < rowsO; ++o) {  // rare: out-edges beyond the register rows
        if (o < nO) {
          const float4 c = s_slot[(rowsIn + o) * stride + ent];
          gx += c.x;
          g1 += c.y;
          g2 += c.z;
        }
      }
      if (v_id >= 0) {
        const float xo = vx, w1o = vw1, w2o = vw2;
        const float xp = fmaf(-tau, gx, xo);
        const float w1n = fmaf(-tau, g1, w1o);
        const float w2n = fmaf(-tau, g2, w2o);
        const float d = xp - vz;
        float xn = (d > vth) ? (xp - vth) : ((d < -vth) ? (xp + vth) : vz);
        xn = fminf(fmaxf(xn, xmin), xmax);
        vx = xn; vw1 = w1n; vw2 = w2n;
        xb = fmaf(theta, xn - xo, xn);
        w1b = fmaf(theta, w1n - w1o, w1n);
        w2b = fmaf(theta, w2n - w2o, w2n);
        const float4 nb = make_float4(xb, w1b, w2b, 0.f);
        if (more) {
          if (v_id & 0x40000000) {  // boundary vertex: hand the point to the CTAs across the cut
            if (CLUSTER) {
              const uint32_t moff = 8u * (uint32_t)((it + 1) & 1);
#pragma unroll 1
              for (uint32_t p = v_push & 0xffffu; p < (v_push >> 16); ++p) {
                const uint2 e = s_push[p];
                fbc_st_async(e.x + wr_off, nb, e.y + moff);
              }
            } else {
              fbg_st_mailbox(pub0 + (size_t)(it & 1) * a.pstride + (v_id & 0x3fffffff), nb.x, nb.y, nb.z,
                             tag0 + (uint32_t)it);
            }
          }
          fbc_sts(bar_base + wr_off + 16u * (uint32_t)ent, nb);
        } else {
          g.vbar[vb + (v_id & 0x3fffffff)] = nb;
        }
      }
    }
    FBG_STAMP(5)
  }

  // ---- write back: registers -> global -----------------------------------------------------------
  if (v_id >= 0) {
    const size_t v = vb + (v_id & 0x3fffffff);
    g.x[v] = vx; g.w1[v] = vw1; g.w2[v] = vw2;
#pragma unroll
    for (int k = 0; k < FBG_FAST; ++k)
      if (fvalid & (1u << k)) {
        const int id = a.feid[((size_t)s * FBG_FAST + k) * g.maxV + c0.x + tid];
        g.q4[eb + id] = make_float4(F.q1[k], F.q2[k], F.q3[k], 0.f);
      }
  }
  if (g_id >= 0) g.q4[eb + g_id] = make_float4(G.q1, G.q2, G.q3, 0.f);
  if (CLUSTER) fbc_cluster_sync();  // no CTA retires while a peer could still address its shared memory
}

// CTA shapes of the cluster transport: 512 threads x 1 CTA per SM (126 registers), or smaller CTAs of
// which two share an SM (<= 96 / 80 registers), so that two streams' phases overlap on one SM.
struct FbgClCfg { int threads, per_sm; size_t smem_limit; };
static const FbgClCfg FBG_CL[FBG_NCL] = {{512, 1, FBG_SMEM_LIMIT_CL}, {384, 2, FBG_SMEM_LIMIT_L2}, {320, 2, FBG_SMEM_LIMIT_L2}};
static const void* fbg_cl_kernel(int k) {
  switch (k) {
    case 1: return (const void*)k_nltgv2_grid<384, 2, true>;
    case 2: return (const void*)k_nltgv2_grid<320, 2, true>;
    default: return (const void*)k_nltgv2_grid<512, 1, true>;
  }
}

// ---------------------------------------------------------------------------------- host side
static inline size_t fbg_smem_bytes(int capBar, int capSlot, int capPush) {
  return 16 * (2 * (size_t)capBar + (size_t)capSlot + 1) + 16 + 8 * (size_t)capPush;
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

// Shared-memory wavefronts one quarter-warp (8 lanes x 16 B) needs for a set of records: the largest
// number of DIFFERENT records falling into one 16-byte bank group (equal records merge).
static inline int fbg_group_cost(const int* rec, int n) {
  int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 0;
  for (int a = 0; a < n; ++a) {
    bool dup = false;
    for (int b = 0; b < a; ++b) dup |= rec[b] == rec[a];
    if (!dup) worst = std::max(worst, ++cnt[rec[a] & 7]);
  }
  return worst;
}
// Smooth surrogate of the above for the local search: pairs of different records in one bank group.
static inline int fbg_group_pairs(const int* rec, int n) {
  int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pairs = 0;
  for (int a = 0; a < n; ++a) {
    bool dup = false;
    for (int b = 0; b < a; ++b) dup |= rec[b] == rec[a];
    if (!dup) pairs += cnt[rec[a] & 7]++;
  }
  return pairs;
}

// Build the host tables of one stream for `nper` parts and a CTA of `threads` threads (one vertex
// and one generic edge per thread).  Returns false when a part exceeds the per-CTA capacity (the
// caller then tries more parts, the other transport or another variant).
static bool fbg_build(GridPlan::Topo& g, const ClusterPlan::Topo& t, int nper, int threads, size_t smem_limit) {
  const int V = t.V, E = t.E;
  g.planned = nper;
  g.planned_threads = threads;
  g.feasible = false;
  g.capBar = g.capSlot = g.capPush = 0;
  g.cinfo.assign((size_t)3 * nper, make_int4(0, 0, 0, 0));
  g.gplan.clear(); g.geid.clear(); g.hplan.clear(); g.pplan.clear();
  g.vplan.assign(V, make_int4(0, 0, 0, 0));
  g.fplan.assign((size_t)FBG_FAST * V, 0u);
  g.feid.assign((size_t)FBG_FAST * V, -1);
  if (V == 0) {
    g.feasible = true;
    return true;
  }
  // CSR rows hold the target-role incidences (edges (u,v), u < v) before the source-role ones
  // (edges (v,w)): edges are sorted by (i,j) with i < j.  The kernel's summation order relies on it.
  std::vector<int> nin(V, 0), deg(V), wgt(V), ids(V), part(V, 0);
  for (int v = 0; v < V; ++v) {
    deg[v] = t.row[v + 1] - t.row[v];
    bool src_seen = false;
    for (int k = t.row[v]; k < t.row[v + 1]; ++k) {
      if (t.inc[k] & 1) {
        if (src_seen) return false;
        nin[v]++;
      } else {
        src_seen = true;
      }
    }
    wgt[v] = 2 + deg[v];
  }
  std::iota(ids.begin(), ids.end(), 0);
  fbg_rcb(g.pos.data(), wgt.data(), ids.data(), 0, V, 0, nper, part.data());
  std::vector<int> pdst(E);  // position of edge e among the target-role incidences of its target
  std::vector<uint8_t> bnd(V, 0);
  for (int e = 0; e < E; ++e)
    if (part[t.eij[e].x] != part[t.eij[e].y]) bnd[t.eij[e].x] = bnd[t.eij[e].y] = 1;
  for (int v = 0; v < V; ++v)
    for (int k = t.row[v]; k < t.row[v] + nin[v]; ++k) pdst[t.inc[k] >> 1] = k - t.row[v];
  // thread order per part: interior vertices, then boundary vertices (the high-numbered warps are
  // issued first by the SMSP arbiter and carry the exchange), each by in-degree and out-degree -- the lanes of a
  // warp then run equally many slot rows and register rows.  Shared-memory records are addressed through an ENTRY index that
  // is a permutation of the thread index inside each aligned group of 8 (chosen below), slots are
  // slot-major with an odd row stride: record (row p, entry n) at p * stride + n.
  std::vector<int> cnt(nper + 1, 0);
  for (int v = 0; v < V; ++v) cnt[part[v] + 1]++;
  for (int r = 0; r < nper; ++r) cnt[r + 1] += cnt[r];
  std::vector<int> order(V);
  {
    std::vector<int> fill(cnt.begin(), cnt.begin() + nper);
    for (int v = 0; v < V; ++v) order[fill[part[v]]++] = v;
  }
  std::vector<int> lidx(V), ent(V), nent(nper), stride(nper), rows_in(nper, 0), nslot(nper, 0);
  for (int r = 0; r < nper; ++r) {
    std::sort(order.begin() + cnt[r], order.begin() + cnt[r + 1], [&](int u, int v) {
      if (bnd[u] != bnd[v]) return bnd[u] < bnd[v];  // boundary vertices last: few warps run the hand-over
      if (nin[u] != nin[v]) return nin[u] > nin[v];
      if (deg[u] != deg[v]) return deg[u] > deg[v];
      return u < v;
    });
    const int nOwn = cnt[r + 1] - cnt[r];
    int rows_ov = 0;
    for (int k = cnt[r]; k < cnt[r + 1]; ++k) {
      const int v = order[k];
      const int novf = std::max(0, deg[v] - nin[v] - FBG_FAST);
      if (nin[v] > 255 || novf > 255) return false;
      lidx[v] = ent[v] = k - cnt[r];
      rows_in[r] = std::max(rows_in[r], nin[v]);
      rows_ov = std::max(rows_ov, novf);
    }
    nent[r] = (nOwn + 7) & ~7;
    stride[r] = nent[r] + 1;  // odd: the rows of one entry fall into different bank groups
    nslot[r] = (rows_in[r] + rows_ov) * stride[r];  // the dummy record follows the last row
    if (nOwn > threads || nslot[r] + 1 > 0xffff) return false;
    g.cinfo[3 * r].x = cnt[r];
    g.cinfo[3 * r].y = nOwn;
    g.cinfo[3 * r + 1].z = nslot[r];
    g.cinfo[3 * r + 1].w = rows_in[r] | (stride[r] << 8);
    g.capSlot = std::max(g.capSlot, nslot[r]);
  }
  // halo entries (sorted vertex ids) follow the own entries; push lists name them for the owners
  std::vector<std::vector<int>> halo(nper);
  std::vector<std::vector<int4>> push(nper);  // per owner part: {owner-local thread, consumer part, consumer entry}
  for (int e = 0; e < E; ++e) {
    const int i = t.eij[e].x, j = t.eij[e].y, ri = part[i], rj = part[j];
    if (ri == rj) continue;
    halo[ri].push_back(j);
    halo[rj].push_back(i);
  }
  for (int r = 0; r < nper; ++r) {
    std::vector<int>& h = halo[r];
    std::sort(h.begin(), h.end());
    h.erase(std::unique(h.begin(), h.end()), h.end());
    if (nent[r] + (int)h.size() > 0xffff) return false;
    g.cinfo[3 * r + 1].x = (int)g.hplan.size();
    g.cinfo[3 * r + 1].y = (int)h.size();
    for (size_t k = 0; k < h.size(); ++k) {
      g.hplan.push_back(h[k]);
      push[part[h[k]]].push_back(make_int4(lidx[h[k]], r, nent[r] + (int)k, 0));
    }
    g.capBar = std::max(g.capBar, nent[r] + (int)h.size());
  }
  // register rows: the first FBG_FAST out-edges of every vertex in ascending edge id
  std::vector<int> fe((size_t)FBG_FAST * V, -1), nfast(V, 0);
  for (int e = 0; e < E; ++e) {
    const int i = t.eij[e].x;
    if (nfast[i] < FBG_FAST) fe[(size_t)nfast[i]++ * V + cnt[part[i]] + lidx[i]] = e;
  }
  // ---- entry permutation: inside every aligned group of 8 threads, swap entries while that lowers
  // the bank conflicts of the register rows' target-side accesses (LDS of the target's point: bank
  // group = entry & 7; STS of its contribution: bank group = (row + entry) & 7 with the odd stride).
  // Own stores and slot gathers stay conflict-free because each group of 8 keeps 8 distinct entries.
  for (int r = 0; r < nper; ++r) {
    const int nOwn = cnt[r + 1] - cnt[r];
    if (nOwn <= 8) continue;
    const int ngrp = (nOwn + 7) / 8;
    // groups = (quarter-warp, row); members = (target vertex, slot row) with the target in this part
    std::vector<std::vector<int>> memb((size_t)ngrp * FBG_FAST);  // packed: target local thread | row << 16
    std::vector<std::vector<int>> of_vertex(nOwn);             // groups a local vertex is a target in
    for (int k = 0; k < FBG_FAST; ++k)
      for (int tl = 0; tl < nOwn; ++tl) {
        const int e = fe[(size_t)k * V + cnt[r] + tl];
        if (e < 0) continue;
        const int j = t.eij[e].y;
        if (part[j] != r) continue;  // halo targets keep their fixed entries: left out of the search
        const int gi = (tl / 8) * FBG_FAST + k;
        memb[gi].push_back(lidx[j] | (pdst[e] << 16));
        of_vertex[lidx[j]].push_back(gi);
      }
    std::vector<int> e_of(nOwn);  // local thread -> entry
    for (int tl = 0; tl < nOwn; ++tl) e_of[tl] = tl;
    auto cost = [&](int gi) {
      int ld[8], st[8];
      const std::vector<int>& m = memb[gi];
      const int n = std::min((int)m.size(), 8);
      for (int a = 0; a < n; ++a) {
        const int en = e_of[m[a] & 0xffff], row = m[a] >> 16;
        ld[a] = en;
        st[a] = row * stride[r] + en;
      }
      return fbg_group_pairs(ld, n) + fbg_group_pairs(st, n);
    };
    for (int sweep = 0; sweep < 2; ++sweep) {
      bool improved = false;
      for (int q = 0; q < ngrp; ++q) {
        const int lo = q * 8, hi = std::min(nOwn, lo + 8);
        for (int u = lo; u < hi; ++u)
          for (int v = u + 1; v < hi; ++v) {
            if (of_vertex[u].empty() && of_vertex[v].empty()) continue;
            int before = 0, after = 0;
            for (int gi : of_vertex[u]) before += cost(gi);
            for (int gi : of_vertex[v]) before += cost(gi);
            std::swap(e_of[u], e_of[v]);
            for (int gi : of_vertex[u]) after += cost(gi);
            for (int gi : of_vertex[v]) after += cost(gi);
            if (after < before) improved = true;
            else std::swap(e_of[u], e_of[v]);
          }
      }
      if (!improved) break;
    }
    for (int tl = 0; tl < nOwn; ++tl) ent[order[cnt[r] + tl]] = e_of[tl];
  }
  auto entry = [&](int r, int v) -> int {  // s_bar entry of vertex v as seen from part r
    if (part[v] == r) return ent[v];
    const std::vector<int>& h = halo[r];
    return nent[r] + (int)(std::lower_bound(h.begin(), h.end(), v) - h.begin());
  };
  for (int v = 0; v < V; ++v) {
    const int novf = std::max(0, deg[v] - nin[v] - FBG_FAST);
    g.vplan[cnt[part[v]] + lidx[v]] = make_int4(v, nin[v] | (novf << 8) | (ent[v] << 16), 0, 0);
  }
  // ---- edge tables
  std::vector<std::vector<int2>> gpl(nper);
  std::vector<std::vector<int>> gid(nper);
  std::vector<int> nf2(V, 0), novf_seen(V, 0);
  for (int e = 0; e < E; ++e) {
    const int i = t.eij[e].x, j = t.eij[e].y, ri = part[i], rj = part[j];
    // the copy in the source's part: a register row of i's thread, or (beyond FBG_FAST) a generic edge
    const int sj_i = (rj == ri) ? pdst[e] * stride[ri] + ent[j] : nslot[ri];
    const int bj_i = entry(ri, j);
    if (nf2[i] < FBG_FAST) {
      const size_t fi = (size_t)nf2[i]++ * V + cnt[ri] + lidx[i];
      g.fplan[fi] = (uint32_t)bj_i | ((uint32_t)sj_i << 16);
      g.feid[fi] = e;
    } else {
      const int si = (rows_in[ri] + novf_seen[i]++) * stride[ri] + ent[i];
      gpl[ri].push_back(make_int2(ent[i] | (bj_i << 16), si | (sj_i << 16)));
      gid[ri].push_back(e);
    }
    // cut edge: the copy in the target's part reads the source from the halo, feeds only the target
    if (rj != ri) {
      gpl[rj].push_back(make_int2(entry(rj, i) | (ent[j] << 16), nslot[rj] | ((pdst[e] * stride[rj] + ent[j]) << 16)));
      gid[rj].push_back(e | (int)0x80000000);
    }
  }
  // first thread that reads a halo entry: boundary vertices with a remote out-neighbour in a register
  // row sit at the end of the own threads, generic edges (cut edges seen from the target side;
  // overflow edges ride along) occupy the last threads of the CTA
  for (int r = 0; r < nper; ++r) {
    int first = threads - (int)gpl[r].size();
    for (int tl = 0; tl < cnt[r + 1] - cnt[r] && tl < first; ++tl)
      for (int k = 0; k < FBG_FAST; ++k) {
        const size_t fi = (size_t)k * V + cnt[r] + tl;
        if (g.feid[fi] >= 0 && (int)(g.fplan[fi] & 0xffffu) >= nent[r]) first = std::min(first, tl);
      }
    g.cinfo[3 * r + 2].z = std::max(0, std::min(first, threads - 1));
  }
  for (int r = 0; r < nper; ++r) {
    if ((int)gpl[r].size() > threads) return false;
    g.cinfo[3 * r].z = (int)g.gplan.size();
    g.cinfo[3 * r].w = (int)gpl[r].size();
    g.gplan.insert(g.gplan.end(), gpl[r].begin(), gpl[r].end());
    g.geid.insert(g.geid.end(), gid[r].begin(), gid[r].end());
  }
  // push lists (cluster transport): per owner part, grouped by vertex in thread order
  for (int r = 0; r < nper; ++r) {
    std::vector<int4>& pl = push[r];
    std::stable_sort(pl.begin(), pl.end(), [](const int4& x, const int4& y) { return x.x < y.x; });
    if (pl.size() > 0xffff) return false;
    const int pbeg = (int)g.pplan.size();
    size_t k = 0;
    for (int li = 0; li < cnt[r + 1] - cnt[r]; ++li) {
      const size_t b = k;
      while (k < pl.size() && pl[k].x == li) {
        g.pplan.push_back(make_int2(pl[k].y, pl[k].z));
        ++k;
      }
      g.vplan[cnt[r] + li].z = (int)(b | (k << 16));
    }
    g.cinfo[3 * r + 2].x = pbeg;
    g.cinfo[3 * r + 2].y = (int)pl.size();
    g.capPush = std::max(g.capPush, (int)pl.size());
  }
  if (fbg_smem_bytes(g.capBar, g.capSlot, g.capPush) > smem_limit) return false;
  g.feasible = true;
  return true;
}

static int grid_plan_init(fb_ctx* c) {
  if (c->gplan) return FB_OK;
  GridPlan* P = new GridPlan();
  c->gplan = P;
  P->topo.resize(c->S);
  const size_t S = c->S;
  if (dalloc(&P->fplan, S * FBG_FAST * c->maxV) != cudaSuccess || dalloc(&P->feid, S * FBG_FAST * c->maxV) != cudaSuccess ||
      dalloc(&P->gplan, S * 2 * c->maxE) != cudaSuccess || dalloc(&P->geid, S * 2 * c->maxE) != cudaSuccess ||
      dalloc(&P->vplan, S * c->maxV) != cudaSuccess || dalloc(&P->hplan, S * 2 * c->maxE) != cudaSuccess ||
      dalloc(&P->pplan, S * 2 * c->maxE) != cudaSuccess ||
      dalloc(&P->cinfo, S * FBG_MAXP * 3) != cudaSuccess || dalloc(&P->pub, 2 * S * c->maxV) != cudaSuccess ||
      cudaHostAlloc((void**)&P->err, sizeof(int), cudaHostAllocMapped) != cudaSuccess)
    FB_FAIL(c, FB_E_NOMEM, "grid plan allocation failed");
  *P->err = 0;
  FB_CUDA(c, cudaMemsetAsync(P->pub, 0, sizeof(float4) * 2 * S * c->maxV, c->stream));
  if (const char* e = getenv("FB_GRID_CTAS")) P->budget_env = atoi(e);
  if (const char* e = getenv("FB_GRID_CLUSTER")) P->cluster_env = std::max(0, std::min(FBG_MAXC, atoi(e)));
  if (const char* e = getenv("FB_GRID_THREADS")) P->threads_env = atoi(e);
  if (const char* e = getenv("FB_GRID_MODE")) P->mode_env = (e[0] == 'c') ? 1 : (e[0] == 'l' ? 2 : 0);
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
  c->gplan->version++;
  return FB_OK;
}

static void grid_plan_free(fb_ctx* c) {
  GridPlan* P = c->gplan;
  if (!P) return;
  cudaFree(P->fplan); cudaFree(P->feid); cudaFree(P->gplan); cudaFree(P->geid); cudaFree(P->vplan);
  cudaFree(P->hplan); cudaFree(P->pplan);
  cudaFree(P->cinfo); cudaFree(P->pub);
  if (P->err) cudaFreeHost(P->err);
  delete P;
  c->gplan = nullptr;
}

// True when the watchdog of an earlier variant-3 launch fired (checked after stream syncs).
static bool grid_watchdog_fired(const fb_ctx* c) { return c->gplan && c->gplan->err && *c->gplan->err != 0; }

static inline bool fbg_active(const fb_ctx* c, int s, int only) { return (only < 0 || s == only) && c->hV[s] > 0; }

// (Re)build every active stream's tables for (nper, threads); true when all of them fit.
static bool fbg_plan_all(fb_ctx* c, int only, int nper, int threads, size_t smem_limit) {
  GridPlan* P = c->gplan;
  for (int s = 0; s < c->S; ++s) {
    if (!fbg_active(c, s, only)) continue;
    GridPlan::Topo& g = P->topo[s];
    if (g.planned != nper || g.planned_threads != threads) {
      fbg_build(g, c->plan->topo[s], nper, threads, smem_limit);
      g.dirty = true;
    }
    if (!g.feasible) return false;
  }
  return true;
}

static int fbg_max_clusters(GridPlan* P, int k, int C) {
  if (P->max_clusters[k][C] >= 0) return P->max_clusters[k][C];
  const void* kern = fbg_cl_kernel(k);
  if (P->smem_set[1 + k] < FBG_CL[k].smem_limit) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FBG_CL[k].smem_limit);
    P->smem_set[1 + k] = FBG_CL[k].smem_limit;
  }
  if (C > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t q{};
  q.gridDim = dim3((unsigned)(C * 64));
  q.blockDim = dim3((unsigned)FBG_CL[k].threads);
  q.dynamicSmemBytes = 64 * 1024;  // registers, not shared memory, bound residency
  cudaLaunchAttribute qa[1];
  qa[0].id = cudaLaunchAttributeClusterDimension;
  qa[0].val.clusterDim.x = (unsigned)C;
  qa[0].val.clusterDim.y = 1;
  qa[0].val.clusterDim.z = 1;
  q.attrs = qa;
  q.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  P->max_clusters[k][C] = n;
  return n;
}

// Pick the transport and the parts per stream, (re)build and upload the tables.  Returns FB_OK with
// *nper_out = 0 when the batch does not fit (the caller falls back to another variant).
static int grid_prepare(fb_ctx* c, int iters, int only, int* nper_out, size_t* smem_out) {
  *nper_out = 0;
  if (!c->gplan || !c->plan) return FB_OK;
  GridPlan* P = c->gplan;
  if (P->dec_version == P->version && P->dec_only == only && (P->dec_cluster || iters <= FBG_MAX_ITERS)) {
    P->nper = P->dec_nper;  // nothing changed since the last launch: tables are on the device
    P->cluster = P->dec_cluster;
    P->cl_cfg = P->dec_cfg;
    *nper_out = P->dec_nper;
    *smem_out = P->dec_smem;
    return FB_OK;
  }
  int n_act = 0, maxV = 0, maxE = 0;
  for (int s = 0; s < c->S; ++s) {
    if (!fbg_active(c, s, only)) continue;
    ++n_act;
    maxV = std::max(maxV, c->hV[s]);
    maxE = std::max(maxE, c->hE[s]);
  }
  if (n_act == 0) return FB_OK;
  int nper = 0;
  bool cluster = false;
  // ---- cluster transport: the largest cluster size (<= 16) for which all active streams'
  // clusters are co-resident (one wave), never below ~128 vertices per CTA
  if (P->mode_env != 2) {
    // CTA shape: FB_GRID_THREADS forces one; otherwise one 512-thread CTA per SM while all streams'
    // clusters are co-resident (measured, C2: 8 streams 74 us against 94 us with two 320-thread CTAs
    // per SM), and the two-per-SM shapes for larger batches (12 streams: 98 us against 135 us in two
    // waves of 512-thread clusters)
    int order[FBG_NCL], norder = 0;
    for (int k = 0; k < FBG_NCL; ++k)
      if (P->threads_env == FBG_CL[k].threads) order[norder++] = k;
    if (norder == 0) {
      order[norder++] = 0;
      order[norder++] = 2;
      order[norder++] = 1;
    }
    for (int oi = 0; oi < norder && !cluster; ++oi) {
      const int k = order[oi], thr = FBG_CL[k].threads;
      const int need = std::max(1, fb_div_up(maxV, thr));
      if (need > FBG_MAXC) continue;
      int cmax = std::min(FBG_MAXC, std::max(need, maxV / 128));
      if (P->cluster_env > 0) cmax = std::max(need, std::min(FBG_MAXC, P->cluster_env));
      for (int cand = cmax; cand >= need && cand >= 1 && !cluster; --cand) {
        // co-residency first (cheap, cached per shape and size), then the tables
        if (P->cluster_env == 0 && fbg_max_clusters(P, k, cand) < n_act) continue;
        if (!fbg_plan_all(c, only, cand, thr, FBG_CL[k].smem_limit)) continue;
        nper = cand;
        cluster = true;
        P->cl_cfg = k;
      }
    }
    // more streams than co-resident clusters of any size: the smallest feasible cluster of the largest
    // CTA shape, in waves (streams are independent, so clusters need not run at the same time)
    for (int oi = 0; oi < norder && !cluster; ++oi) {
      const int k = order[oi], thr = FBG_CL[k].threads;
      for (int cand = std::max(1, fb_div_up(maxV, thr)); cand <= FBG_MAXC && !cluster; ++cand) {
        if (!fbg_plan_all(c, only, cand, thr, FBG_CL[k].smem_limit)) continue;
        nper = cand;
        cluster = true;
        P->cl_cfg = k;
      }
    }
  }
  // ---- L2 transport: all co-resident CTAs shared by the streams of the launch
  if (!cluster && P->mode_env != 1 && iters <= FBG_MAX_ITERS) {
    auto kern = k_nltgv2_grid<FBG_THREADS_L2, 2, false>;
    if (P->max_blocks < 0) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FBG_SMEM_LIMIT_L2);
      P->smem_set[0] = FBG_SMEM_LIMIT_L2;
      int per_sm = 0, sms = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FBG_THREADS_L2, FBG_SMEM_LIMIT_L2) != cudaSuccess) {
        cudaGetLastError();
        per_sm = 0;
      }
      P->max_blocks = per_sm * sms;
    }
    int budget = P->max_blocks;
    if (P->budget_env > 0) budget = std::min(budget, P->budget_env);
    const int n_div = only >= 0 ? 1 : c->S;  // CTAs of empty streams are launched too (they exit at once)
    nper = 0;
    if (budget >= n_div) {
      // never below ~96 vertices per part (the exchange then dominates), at least what the capacities need
      nper = std::min(budget / n_div, FBG_MAXP);
      nper = std::max(1, std::min(nper, std::max(1, maxV / (P->budget_env > 0 ? 16 : 96))));
      const int need = fb_div_up(maxV, FBG_THREADS_L2);
      nper = std::max(nper, need);
      bool ok = false;
      for (; nper * n_div <= budget && nper <= FBG_MAXP; nper += std::max(1, nper / 8))
        if ((ok = fbg_plan_all(c, only, nper, FBG_THREADS_L2, FBG_SMEM_LIMIT_L2))) break;
      if (!ok) nper = 0;
    }
  }
  if (nper == 0) {  // does not fit: remember that too (the caller falls back to another variant)
    P->dec_version = P->version;
    P->dec_only = only;
    P->dec_nper = 0;
    P->dec_cluster = true;
    P->dec_smem = 0;
    return FB_OK;
  }
  int capBar = 1, capSlot = 1, capPush = 0;
  cudaStream_t st = c->stream;
  for (int s = 0; s < c->S; ++s) {
    if (!fbg_active(c, s, only)) continue;
    GridPlan::Topo& g = P->topo[s];
    capBar = std::max(capBar, g.capBar);
    capSlot = std::max(capSlot, g.capSlot);
    capPush = std::max(capPush, g.capPush);
    if (!g.dirty) continue;
    const size_t vb = (size_t)s * c->maxV, eb2 = (size_t)s * 2 * c->maxE;
    if (g.gplan.size() > 2 * (size_t)c->maxE || g.hplan.size() > 2 * (size_t)c->maxE || g.pplan.size() > 2 * (size_t)c->maxE)
      FB_FAIL(c, FB_E_NOMEM, "grid plan: tables exceed capacity");
    const size_t Vs = g.vplan.size();
    FB_CUDA(c, cudaMemcpyAsync(P->vplan + vb, g.vplan.data(), sizeof(int4) * Vs, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < FBG_FAST; ++k) {  // row-major on the device with the context's vertex capacity as stride
      const size_t dv = ((size_t)s * FBG_FAST + k) * c->maxV;
      FB_CUDA(c, cudaMemcpyAsync(P->fplan + dv, g.fplan.data() + k * Vs, sizeof(uint32_t) * Vs, cudaMemcpyHostToDevice, st));
      FB_CUDA(c, cudaMemcpyAsync(P->feid + dv, g.feid.data() + k * Vs, sizeof(int32_t) * Vs, cudaMemcpyHostToDevice, st));
    }
    if (!g.gplan.empty()) {
      FB_CUDA(c, cudaMemcpyAsync(P->gplan + eb2, g.gplan.data(), sizeof(int2) * g.gplan.size(), cudaMemcpyHostToDevice, st));
      FB_CUDA(c, cudaMemcpyAsync(P->geid + eb2, g.geid.data(), sizeof(int32_t) * g.geid.size(), cudaMemcpyHostToDevice, st));
    }
    if (!g.hplan.empty())
      FB_CUDA(c, cudaMemcpyAsync(P->hplan + eb2, g.hplan.data(), sizeof(int32_t) * g.hplan.size(), cudaMemcpyHostToDevice, st));
    if (!g.pplan.empty())
      FB_CUDA(c, cudaMemcpyAsync(P->pplan + eb2, g.pplan.data(), sizeof(int2) * g.pplan.size(), cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(P->cinfo + (size_t)s * FBG_MAXP * 3, g.cinfo.data(), sizeof(int4) * g.cinfo.size(), cudaMemcpyHostToDevice, st));
    g.dirty = false;  // host tables stay alive in the Topo: no staging lifetime issue
  }
  P->nper = nper;
  P->cluster = cluster;
  P->capBar = capBar;
  P->capSlot = capSlot;
  P->capPush = capPush;
  *nper_out = nper;
  *smem_out = fbg_smem_bytes(capBar, capSlot, capPush);
  P->dec_version = P->version;
  P->dec_only = only;
  P->dec_nper = nper;
  P->dec_cluster = cluster;
  P->dec_cfg = P->cl_cfg;
  P->dec_smem = *smem_out;
  return FB_OK;
}

static int solve_grid(fb_ctx* c, int iters, const fb_nltgv2_params* p, int nper, size_t smem, int only = -1) {
  GridPlan* P = c->gplan;
  // L2 transport: tags are unique per launch; when the 32-bit tag space is used up the mailboxes are cleared
  if (P->seq >= 0xffffffffu / (FBG_MAX_ITERS + 1) - 1) {
    FB_CUDA(c, cudaMemsetAsync(P->pub, 0, sizeof(float4) * 2 * (size_t)c->S * c->maxV, c->stream));
    P->seq = 0;
  }
  P->seq++;
  const uint32_t tag0 = P->seq * (uint32_t)(FBG_MAX_ITERS + 1);
  GridArgs a;
  a.g = graph_view(c);
  a.g.only = only;
  a.fplan = P->fplan; a.feid = P->feid; a.gplan = P->gplan; a.geid = P->geid;
  a.vplan = P->vplan; a.hplan = P->hplan; a.pplan = P->pplan;
  a.cinfo = P->cinfo;
  a.pub = P->pub;
  a.err = P->err;
  a.nper = nper;
  a.capBar = P->capBar;
  a.capSlot = P->capSlot;
  a.pstride = (size_t)c->S * c->maxV;
  a.trace = nullptr;
#ifdef FBG_TRACE
  static long long* d_trace = nullptr;
  const size_t trace_n = (size_t)((only >= 0 ? 1 : c->S) * nper) * 64 * 6;
  if (getenv("FB_GRID_TRACE")) {
    if (!d_trace) cudaMalloc((void**)&d_trace, sizeof(long long) * 1024 * 64 * 6);
    cudaMemsetAsync(d_trace, 0, sizeof(long long) * trace_n, c->stream);
    a.trace = d_trace;
  }
#endif
  c->last_cluster = nper;
  c->last_transport = P->cluster ? 1 : 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((only >= 0 ? 1 : c->S) * nper));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const float tl = p->step_x * p->data_factor;
  if (P->cluster) {
    const int k = P->cl_cfg;
    const void* kern = fbg_cl_kernel(k);
    if (P->smem_set[1 + k] < smem) {
      FB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FBG_CL[k].smem_limit));
      P->smem_set[1 + k] = FBG_CL[k].smem_limit;
    }
    if (nper > 8) FB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cfg.blockDim = dim3((unsigned)FBG_CL[k].threads);
    attr[0].id = cudaLaunchAttributeClusterDimension;  // streams are independent: clusters may run in waves
    attr[0].val.clusterDim.x = (unsigned)nper;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    float sq = p->step_q, sx = p->step_x, th = p->theta, x0 = p->x_min, x1 = p->x_max, tlv = tl;
    uint32_t tg = tag0;
    int itv = iters;
    void* args[] = {&a, &itv, &sq, &sx, &tlv, &th, &x0, &x1, &tg};
    ProfScope ps(c, FB_PROF_SOLVE);
    FB_CUDA(c, cudaLaunchKernelExC(&cfg, kern, args));
    c->last_threads = FBG_CL[k].threads;
  } else {
    auto kern = k_nltgv2_grid<FBG_THREADS_L2, 2, false>;
    cfg.blockDim = dim3(FBG_THREADS_L2);
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the mailbox readers spin
    attr[0].val.cooperative = 1;
    ProfScope ps(c, FB_PROF_SOLVE);
    FB_CUDA(c, cudaLaunchKernelEx(&cfg, kern, a, iters, p->step_q, p->step_x, tl, p->theta, p->x_min, p->x_max, tag0));
    c->last_threads = FBG_THREADS_L2;
  }
#ifdef FBG_TRACE
  if (a.trace) {  // debugging aid: dump the stamps of this launch (synchronises)
    std::vector<long long> h(trace_n);
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(h.data(), d_trace, sizeof(long long) * trace_n, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("FB_GRID_TRACE"), "wb")) {
      fwrite(h.data(), sizeof(long long), trace_n, f);
      fclose(f);
    }
  }
#endif
  c->launches++;
  return FB_OK;
}

// ---------------------------------------------------------------------------------- plan verifier
// Host-only structural check of the variant-3 tables for one graph (no device needed): used by the
// CPU test-suite to validate the partitioner against the invariants the kernel relies on.
// stats[8] = {max own vertices, max generic edges per part, max halo, cut edges (held twice),
//             max slots, shared memory bytes, boundary vertices, edges beyond the register rows}.
static int fbg_verify(int V, int E, const float* pos, const int32_t* ij, int nper, int threads, size_t smem_limit,
                      int32_t* stats, std::string& why) {
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
  if (!fbg_build(g, t, nper, threads, smem_limit)) {
    why = "partition infeasible for this part count";
    return 1;
  }
  std::vector<int> owner(V, -1), lidx(V, -1), written(E, 0), ent_of(V, -1);
  std::vector<std::set<int>> ent_used(nper);
  int maxOwn = 0, maxGen = 0, maxHalo = 0, dup = 0, maxSlot = 0, nb = 0, novf = 0;
  for (int r = 0; r < nper; ++r) {
    const int4 c0 = g.cinfo[3 * r], c1 = g.cinfo[3 * r + 1];
    if (c0.y > threads || c0.w > threads) { why = "part exceeds the CTA's thread count"; return 2; }
    for (int k = 0; k < c0.y; ++k) {
      const int4 pv = g.vplan[c0.x + k];
      if (pv.x < 0 || pv.x >= V || owner[pv.x] != -1) { why = "vertex owned twice or out of range"; return 2; }
      owner[pv.x] = r;
      lidx[pv.x] = k;
      const int nt = pv.y & 0xff, no = (pv.y >> 8) & 0xff, en = (int)((uint32_t)pv.y >> 16);
      const int rows_in = c1.w & 0xff, stride = c1.w >> 8;
      if (en >= stride - 1 || en / 8 != k / 8) { why = "entry must be a permutation inside the thread's group of 8"; return 16; }
      if (ent_used[r].count(en)) { why = "entry used twice"; return 16; }
      ent_used[r].insert(en);
      ent_of[pv.x] = en;
      if (nt > rows_in) { why = "in-degree beyond the part's in-edge rows"; return 3; }
      int nin = 0;
      for (int q = t.row[pv.x]; q < t.row[pv.x + 1]; ++q) nin += t.inc[q] & 1;
      const int nout = t.row[pv.x + 1] - t.row[pv.x] - nin;
      if (nt != nin || no != std::max(0, nout - FBG_FAST)) { why = "slot block does not match the in-degree / overflow"; return 3; }
      if ((rows_in + no) * stride > c1.z) { why = "slot rows beyond the part's slot count"; return 3; }
      nb += ((uint32_t)pv.z >> 16) != ((uint32_t)pv.z & 0xffffu) ? 1 : 0;
      novf += no;
    }
    maxOwn = std::max(maxOwn, c0.y);
    maxGen = std::max(maxGen, c0.w);
    maxHalo = std::max(maxHalo, c1.y);
    maxSlot = std::max(maxSlot, c1.z);
  }
  for (int v = 0; v < V; ++v)
    if (owner[v] < 0) { why = "vertex without owner"; return 2; }
  // what an (entry, slot) pair must be for endpoint v of edge e seen from part r
  auto check_end = [&](int r, int e, int v, bool is_target, int ovf_pos, int b, int sl, bool slot_used, int& code) -> bool {
    const int4 c1 = g.cinfo[3 * r + 1];
    const int rows_in = c1.w & 0xff, stride = c1.w >> 8;
    if (owner[v] == r) {
      if (b != ent_of[v]) { code = 5; return false; }
      if (!slot_used) return true;
      int want;
      if (is_target) {  // position among the target-role incidences (CSR order)
        int posn = -1;
        for (int q = t.row[v]; q < t.row[v + 1]; ++q)
          if (t.inc[q] == ((e << 1) | 1)) posn = q - t.row[v];
        want = posn * stride + ent_of[v];
        if (posn < 0 || posn >= rows_in) { code = 6; return false; }
      } else {
        want = (rows_in + ovf_pos) * stride + ent_of[v];
      }
      if (sl != want) { code = 6; return false; }
    } else {
      const int h = b - (stride - 1);
      if (h < 0 || h >= c1.y || g.hplan[c1.x + h] != v) { code = 7; return false; }
      if (slot_used && sl != c1.z) { code = 8; return false; }
      // the owner's push list must name exactly this halo entry
      const int4 pv = g.vplan[g.cinfo[3 * owner[v]].x + lidx[v]];
      const int pbeg = g.cinfo[3 * owner[v] + 2].x;
      bool found = false;
      for (uint32_t q = (uint32_t)pv.z & 0xffffu; q < ((uint32_t)pv.z >> 16); ++q) {
        const int2 pe = g.pplan[pbeg + q];
        if (pe.x == r && pe.y == b) found = true;
      }
      if (!found) { code = 9; return false; }
    }
    return true;
  };
  static const char* msg[] = {"", "", "", "", "", "own endpoint index mismatch", "slot != row * stride + entry",
                              "halo index mismatch", "remote endpoint must map to the dummy slot",
                              "halo vertex not pushed/published by its owner"};
  std::vector<std::vector<int>> hit(nper);
  for (int r = 0; r < nper; ++r) hit[r].assign(g.cinfo[3 * r + 1].z + 1, 0);
  // register rows: out-edges of each vertex in ascending edge id
  for (int r = 0; r < nper; ++r) {
    const int4 c0 = g.cinfo[3 * r];
    for (int k = 0; k < c0.y; ++k) {
      const int v = g.vplan[c0.x + k].x;
      int nin = 0;
      for (int q = t.row[v]; q < t.row[v + 1]; ++q) nin += t.inc[q] & 1;
      const int nout = t.row[v + 1] - t.row[v] - nin;
      for (int f = 0; f < FBG_FAST; ++f) {
        const size_t fi = (size_t)f * V + c0.x + k;
        const int e = g.feid[fi];
        if (f >= nout) {
          if (e != -1) { why = "register row beyond the out-degree must be idle"; return 15; }
          continue;
        }
        const int want_e = t.inc[t.row[v] + nin + f] >> 1;  // f-th source-role incidence
        if (e != want_e) { why = "register rows must hold the out-edges in ascending edge id"; return 15; }
        const int j = t.eij[e].y;
        const int b = (int)(g.fplan[fi] & 0xffffu), sl = (int)(g.fplan[fi] >> 16);
        int code = 0;
        if (!check_end(r, e, j, true, 0, b, sl, true, code)) { why = msg[code]; return code; }
        if (owner[j] == r) hit[r][sl]++;
        else if (k < g.cinfo[3 * r + 2].z) { why = "halo read by a thread below the halo-reading bound"; return 17; }
        written[e]++;
      }
    }
  }
  // generic edges
  std::vector<int> ovf_seen(V, 0);
  for (int r = 0; r < nper; ++r) {
    const int4 c0 = g.cinfo[3 * r], c1 = g.cinfo[3 * r + 1];
    for (int k = 0; k < c0.w; ++k) {
      const int code_e = g.geid[c0.z + k], e = code_e & 0x7fffffff;
      if (e >= E) { why = "edge id out of range"; return 4; }
      const int2 pl = g.gplan[c0.z + k];
      const int i = t.eij[e].x, j = t.eij[e].y;
      const int bi = pl.x & 0xffff, bj = (int)((uint32_t)pl.x >> 16), si = pl.y & 0xffff, sj = (int)((uint32_t)pl.y >> 16);
      const bool wb = code_e >= 0;
      int code = 0;
      if (threads - 1 - k < g.cinfo[3 * r + 2].z) { why = "generic edge on a thread below the halo-reading bound"; return 17; }
      if (wb) {  // overflowed out-edge of an own vertex
        if (owner[i] != r) { why = "write-back copy must live with the source vertex"; return 11; }
        if (!check_end(r, e, i, false, ovf_seen[i], bi, si, true, code)) { why = msg[code]; return code; }
        ovf_seen[i]++;
        hit[r][si]++;
        if (!check_end(r, e, j, true, 0, bj, sj, true, code)) { why = msg[code]; return code; }
        if (owner[j] == r) hit[r][sj]++;
        written[e]++;
      } else {  // cut edge seen from the target's part
        if (owner[j] != r || owner[i] == r) { why = "non-write-back copy must be a cut edge in the target's part"; return 11; }
        if (!check_end(r, e, i, false, 0, bi, si, true, code)) { why = msg[code]; return code; }
        if (!check_end(r, e, j, true, 0, bj, sj, true, code)) { why = msg[code]; return code; }
        hit[r][sj]++;
        ++dup;
      }
      (void)c1;
    }
  }
  for (int r = 0; r < nper; ++r) {
    const int4 c0 = g.cinfo[3 * r];
    for (int k = 0; k < c0.y; ++k) {
      const int4 pv = g.vplan[c0.x + k];
      const int4 c1 = g.cinfo[3 * r + 1];
      const int rows_in = c1.w & 0xff, stride = c1.w >> 8, en = (int)((uint32_t)pv.y >> 16);
      for (int q = 0; q < (pv.y & 0xff); ++q)
        if (hit[r][q * stride + en] != 1) { why = "slot not written exactly once"; return 12; }
      for (int q = 0; q < ((pv.y >> 8) & 0xff); ++q)
        if (hit[r][(rows_in + q) * stride + en] != 1) { why = "overflow slot not written exactly once"; return 12; }
    }
  }
  for (int e = 0; e < E; ++e)
    if (written[e] != 1) { why = "edge not written back exactly once"; return 13; }
  {  // every push entry corresponds to one halo entry
    size_t np = 0, nh = 0;
    for (int r = 0; r < nper; ++r) { np += g.cinfo[3 * r + 2].y; nh += g.cinfo[3 * r + 1].y; }
    if (np != nh || np != g.pplan.size()) { why = "push lists and halo lists differ in size"; return 14; }
  }
  // shared-memory wavefronts of the register rows' target-side accesses (LDS.128 of the target's
  // point, STS.128 of its contribution): a quarter-warp (8 lanes x 16 B) is one wavefront when its
  // lanes touch 8 different 16-byte bank groups (equal addresses broadcast / merge)
  long long wf_ideal = 0, wf_ld = 0, wf_st = 0;
  for (int r = 0; r < nper; ++r) {
    const int4 c0 = g.cinfo[3 * r];
    for (int f = 0; f < FBG_FAST; ++f)
      for (int q0 = 0; q0 < c0.y; q0 += 8) {
        int cl[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int seen_l[8], seen_s[8], nl = 0, ns = 0;
        bool any = false;
        for (int k = q0; k < std::min(q0 + 8, c0.y); ++k) {
          const size_t fi = (size_t)f * V + c0.x + k;
          if (g.feid[fi] < 0) continue;
          any = true;
          const int b = (int)(g.fplan[fi] & 0xffffu), sl = (int)(g.fplan[fi] >> 16);
          bool dup_l = false, dup_s = false;
          for (int m = 0; m < nl; ++m) dup_l |= seen_l[m] == b;
          for (int m = 0; m < ns; ++m) dup_s |= seen_s[m] == sl;
          if (!dup_l) { seen_l[nl++] = b; cl[b & 7]++; }
          if (!dup_s) { seen_s[ns++] = sl; cs[sl & 7]++; }
        }
        if (!any) continue;
        wf_ideal++;
        wf_ld += *std::max_element(cl, cl + 8);
        wf_st += *std::max_element(cs, cs + 8);
      }
  }
  if (stats) {
    stats[8] = (int32_t)wf_ideal; stats[9] = (int32_t)wf_ld; stats[10] = (int32_t)wf_st; stats[11] = 0;
    stats[0] = maxOwn; stats[1] = maxGen; stats[2] = maxHalo; stats[3] = dup; stats[4] = maxSlot;
    stats[5] = (int32_t)fbg_smem_bytes(g.capBar, g.capSlot, g.capPush); stats[6] = nb; stats[7] = novf;
  }
  return 0;
}
