// nltgv2_tile.cuh -- tile-resident NLTGV2-L1 solver planned ON THE DEVICE (variant 5).
//
// flame::Flame::update re-triangulates every frame and ~2 % of the edges change, so the per-topology
// tables of variants 2 / 3 (built on the host in ~1 ms) can never be reused; variant 4 needs no tables
// but moves ~1.9 MB per iteration through the L2 port of ONE GPC (a cluster lives in one GPC) and is
// bound by it: 270 us per 50 iterations at 5.6k vertices (profiles/r2_update_launch_table_v1.md).
// This variant gets the locality of variants 2 / 3 from a plan that costs one small kernel:
//   k_tile_assign  (one CTA)  a balanced 4 x 4 k-d split of the vertex positions (4 vertical strips
//                  of equal count, each cut into 4 parts of equal count): tile and in-tile index of
//                  every vertex.  Tiles are compact, so ~90 % of the edges stay inside a tile.
//   k_nltgv2_tile  one cluster of 16 CTAs per stream, CTA = tile.  The prologue derives everything
//                  else in parallel from the CSR arrays already on the device (own edges = the
//                  contiguous out-edge ranges of the own vertices, slot of every incidence = CSR
//                  position), then all iterations run with the state in registers and the exchange in
//                  shared memory: an edge thread reads the two extragradient points (the target's from
//                  a halo record of its own tile when the target lives in another one: the owner pushes
//                  it there after every primal half-step) and stores the two K^T q contributions into
//                  the CSR slots of its endpoints (the target's through DSMEM, st.async + mbarrier
//                  complete_tx); a vertex thread sums its slots in CSR order.  Point-to-point mbarriers,
//                  two CTA barriers per iteration, no global memory traffic inside the loop.
// Arithmetic and summation order are those of nltgv2.cuh: results are bit-identical.
#pragma once

#include "common.cuh"
#include "delaunay_gpu.cuh"
#include "nltgv2.cuh"
#include "nltgv2_cluster.cuh"

#define FBT_THREADS 512
#define FBT_VPT 2
#define FBT_EPT 5
#define FBT_VCAP (FBT_THREADS * FBT_VPT)   // vertices per tile
#define FBT_ECAP (FBT_THREADS * FBT_EPT)   // out-edges per tile
#define FBT_SLOTCAP 7168                   // incidences per tile
#define FBT_C 16                           // tiles = cluster size
#define FBT_TX 4
#define FBT_TY 4
#define FBT_BINS 64
#define FBT_APT 16                         // k_tile_assign: vertices per thread (16 tiles x 1024 vertices / 1024 threads)

struct TilePlan {
  int32_t* vtile = nullptr;  // [S*maxV] tile of every vertex
  int32_t* vloc = nullptr;   // [S*maxV] index inside its tile
  int32_t* tlist = nullptr;  // [S*maxV] vertex ids in tile order
  int32_t* toff = nullptr;   // [S*(FBT_C+1)]
  int32_t* lrow = nullptr;   // [S*maxV] slot base of every vertex inside its tile (written by the solver's prologue)
  int* err = nullptr;        // mapped host flag: capacity exceeded
  int* derr = nullptr;       // the same flag in device memory (read by every CTA after the prologue)
  std::vector<char> dirty;   // per stream: positions / topology changed since the last k_tile_assign
  int available = -1;        // -1 not probed, 0 no (cluster of 16 refused), 1 yes
  size_t smem = 0;
};

#define FBT_HCAP 1024                      // halo records of a tile: one per out-edge whose target lives in another tile
#define FBT_PCAP 1024                      // push list of a tile: one entry per such edge of the OTHER tiles
static inline size_t fbt_smem_bytes() {
  return sizeof(float4) * (FBT_VCAP + FBT_SLOTCAP + FBT_HCAP) + sizeof(uint2) * (FBT_PCAP + FBT_ECAP) + sizeof(int) * (2 * (FBT_VCAP + 1) + 2 * FBT_VCAP) + 64;
}

// ------------------------------------------------------------------------------------ k_tile_assign
// Exclusive prefix of 64 shared-memory counters by one warp: lane l returns the sum before counter 2 l.
static_assert(FBT_BINS == 64, "fbt_scan64: two bins per lane");
__device__ __forceinline__ int fbt_scan64(const int* cnt, int lane) {
  const int n = cnt[2 * lane] + cnt[2 * lane + 1];
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  return incl - n;
}
__global__ void __launch_bounds__(1024)
k_tile_assign(int s, int maxV, const int32_t* __restrict__ nV, const float2* __restrict__ vpos, int32_t* vtile,
              int32_t* vloc, int32_t* tlist, int32_t* toff) {
  __shared__ int s_box[4];
  __shared__ int s_col[FBT_BINS], s_row[FBT_TX][FBT_BINS];
  __shared__ int s_c2s[FBT_BINS];               // bin column -> strip
  __shared__ int s_r2p[FBT_TX][FBT_BINS];       // (strip, bin row) -> part
  __shared__ int s_wcnt[32][FBT_C];             // vertices of tile t seen by warp w (then: exclusive prefix over warps)
  __shared__ int s_off[FBT_C + 1];
  DSG_CLK_DECL
  DSG_CLK
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int V = nV[s];
  const size_t vb = (size_t)s * maxV;
  vpos += vb; vtile += vb; vloc += vb; tlist += vb; toff += (size_t)s * (FBT_C + 1);
  if (tid == 0) { s_box[0] = s_box[1] = 0x7fffffff; s_box[2] = s_box[3] = -0x7fffffff; }
  for (int k = tid; k < FBT_BINS; k += blockDim.x) s_col[k] = 0;
  for (int k = tid; k < FBT_TX * FBT_BINS; k += blockDim.x) (&s_row[0][0])[k] = 0;
  for (int k = tid; k < 32 * FBT_C; k += blockDim.x) (&s_wcnt[0][0])[k] = 0;
  __syncthreads();
  // The kernel is one CTA and every pass used to be a chain of dependent loads (load, atomic, next
  // element): it was bound by memory latency times V / 1024.  Now a thread's vertices (v = tid + 1024 k,
  // at most FBT_APT) are loaded once, all loads in flight together, and stay in registers.
  if (V > FBT_APT * 1024) {   // more than 16 full tiles: let the solver see an overfull tile and decline
    if (tid <= FBT_C) toff[tid] = tid ? V : 0;
    return;
  }
  int px[FBT_APT], py[FBT_APT];
  DSG_CLK  // init
#pragma unroll
  for (int k = 0; k < FBT_APT; ++k) {
    const int v = tid + k * 1024;
    float2 p = make_float2(0.f, 0.f);
    if (v < V) p = vpos[v];
    px[k] = (int)floorf(p.x * 16.0f);
    py[k] = (int)floorf(p.y * 16.0f);
  }
  {  // bounding box in 1/16 px: per-thread, then one redux per warp, then 32 atomics
    int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
#pragma unroll
    for (int k = 0; k < FBT_APT; ++k)
      if (tid + k * 1024 < V) { bx0 = min(bx0, px[k]); by0 = min(by0, py[k]); bx1 = max(bx1, px[k]); by1 = max(by1, py[k]); }
    bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
    bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
    if (lane == 0) { atomicMin(&s_box[0], bx0); atomicMin(&s_box[1], by0); atomicMax(&s_box[2], bx1); atomicMax(&s_box[3], by1); }
  }
  __syncthreads();
  const int x0 = s_box[0], y0 = s_box[1];
  DSG_CLK  // load + box
  // histogram bins: any monotone map onto [0, FBT_BINS) will do (the split only balances the tiles,
  // the solver's arithmetic does not depend on it), so one float multiply instead of a 64-bit division
  const float sx = (float)FBT_BINS / (float)(s_box[2] - x0 + 1), sy = (float)FBT_BINS / (float)(s_box[3] - y0 + 1);
#pragma unroll
  for (int k = 0; k < FBT_APT; ++k) {
    if (tid + k * 1024 >= V) continue;
    px[k] = min(FBT_BINS - 1, (int)((float)(px[k] - x0) * sx));   // from here on: the bins
    py[k] = min(FBT_BINS - 1, (int)((float)(py[k] - y0) * sy));
    atomicAdd(&s_col[px[k]], 1);
  }
  __syncthreads();
  DSG_CLK  // column histogram
  // strips of (nearly) equal count: column c belongs to the last strip k whose quota k V / 4 the
  // columns before c have filled (a warp scan; the serial loop with its divisions cost 6 us)
  if (wid == 0) {
    const int ex = fbt_scan64(s_col, lane);
    const int a0 = ex, a1 = ex + s_col[2 * lane];
    int k0 = 0, k1 = 0;
    for (int k = 1; k < FBT_TX; ++k) { k0 += a0 >= k * V / FBT_TX ? 1 : 0; k1 += a1 >= k * V / FBT_TX ? 1 : 0; }
    s_c2s[2 * lane] = k0;
    s_c2s[2 * lane + 1] = k1;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < FBT_APT; ++k) {
    if (tid + k * 1024 >= V) continue;
    px[k] = s_c2s[px[k]];                                          // from here on: the strip
    atomicAdd(&s_row[px[k]][py[k]], 1);
  }
  __syncthreads();
  DSG_CLK  // strips + row histogram
  if (wid < FBT_TX) {  // each strip cut into parts of (nearly) equal count: one warp per strip, same rule
    const int ex = fbt_scan64(s_row[wid], lane);
    const int n0 = s_row[wid][2 * lane], n1 = s_row[wid][2 * lane + 1];
    const int tot = __shfl_sync(0xffffffffu, ex + n0 + n1, 31);
    const int a0 = ex, a1 = ex + n0;
    int k0 = 0, k1 = 0;
    for (int k = 1; k < FBT_TY; ++k) { k0 += a0 >= k * tot / FBT_TY ? 1 : 0; k1 += a1 >= k * tot / FBT_TY ? 1 : 0; }
    s_r2p[wid][2 * lane] = k0;
    s_r2p[wid][2 * lane + 1] = k1;
  }
  __syncthreads();
  // rank inside the tile (any order: the arithmetic does not depend on it): counters per (warp, tile)
  DSG_CLK  // parts
  // keep the atomics' contention inside the warp, a prefix over the warps makes the ranks global
#pragma unroll
  for (int k = 0; k < FBT_APT; ++k) {
    if (tid + k * 1024 >= V) continue;
    const int t = px[k] * FBT_TY + s_r2p[px[k]][py[k]];
    px[k] = t;                                                     // the tile
    py[k] = atomicAdd(&s_wcnt[wid][t], 1);                         // rank among this warp's vertices of the tile
  }
  __syncthreads();
  if (tid < FBT_C) {
    int acc = 0;
    for (int w = 0; w < 32; ++w) { const int n = s_wcnt[w][tid]; s_wcnt[w][tid] = acc; acc += n; }
    s_off[tid + 1] = acc;   // tile sizes, turned into offsets below
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    s_off[0] = 0;
    for (int t = 0; t < FBT_C; ++t) { acc += s_off[t + 1]; s_off[t + 1] = acc; }
  }
  __syncthreads();
  if (tid <= FBT_C) toff[tid] = s_off[tid];
  DSG_CLK  // ranks + prefix
#pragma unroll
  for (int k = 0; k < FBT_APT; ++k) {
    const int v = tid + k * 1024;
    if (v >= V) continue;
    const int t = px[k], l = py[k] + s_wcnt[wid][t];
    vtile[v] = t;
    vloc[v] = l;
    tlist[s_off[t] + l] = v;
  }
  DSG_CLK  // write
  DSG_CLK_PRINT("k_tile_assign")
}

// ------------------------------------------------------------------------------------ k_nltgv2_tile
__device__ __forceinline__ void fbt_st_cluster_v2(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint32_t fbt_atom_add_cluster(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}
// Exclusive scan over the block (one value per thread), uniform total.
__device__ __forceinline__ int fbt_block_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = lane < nw ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  *total = s_warp[nw - 1];
  return (wid ? s_warp[wid - 1] : 0) + incl - v;
}

struct TileArgs {
  GraphView g;
  const int32_t* vtile;
  const int32_t* vloc;
  const int32_t* tlist;
  const int32_t* toff;
  int32_t* lrow;
  int* err;
  int* derr;
};


// Own edge number le of the tile -> (local source vertex, offset among its out-edges).
__device__ __forceinline__ void fbt_edge_of(const int* s_erow, int nOwn, int le, int* lv, int* off) {
  int lo = 0, hi = nOwn;  // largest lv with s_erow[lv] <= le
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_erow[mid] <= le) lo = mid; else hi = mid;
  }
  *lv = lo;
  *off = le - s_erow[lo];
}

// Synchronisation inside the iteration is point to point (the cluster-wide barrier costs ~1 us per use
// with release / acquire semantics: two of them were 80 % of an iteration):
//   A  "the contributions for my vertices are complete": remote edge threads deliver them with
//      st.async ... mbarrier::complete_tx::bytes on MY barrier; I expect 16 B per remote incidence.
//   B  "the extragradient points I read from other tiles are refreshed": the OWNER pushes them.  In the
//      prologue every out-edge with a target in another tile reserves a halo record in its own tile and
//      registers (target vertex, record address) in the owner's push list through DSMEM; after its primal
//      half-step the owner sends each listed point with st.async ... complete_tx on the reader's barrier
//      (16 B per record expected).  The dual half-step then reads only LOCAL shared memory.  (The first
//      version announced the refresh with a remote mbarrier arrive and the readers pulled the points with
//      ld.shared::cluster: one more one-way latency plus a round trip in every iteration.)
// Neither can run ahead by more than one phase: a tile's next dual half-step needs B from exactly the
// tiles whose A it feeds.
__global__ void __launch_bounds__(FBT_THREADS, 1)
k_nltgv2_tile(TileArgs a, int iters, float sigma, float tau, float tl, float theta, float xmin, float xmax) {
  extern __shared__ __align__(16) uint8_t fbt_smem[];
  float4* s_bar = reinterpret_cast<float4*>(fbt_smem);           // [VCAP] extragradient points of the own vertices
  float4* s_slot = s_bar + FBT_VCAP;                             // [SLOTCAP] K^T q contributions, CSR order per vertex
  float4* s_halo = s_slot + FBT_SLOTCAP;                         // [HCAP] points of other tiles' vertices, pushed by their owners
  uint2* s_plist = reinterpret_cast<uint2*>(s_halo + FBT_HCAP);  // [PCAP] {own vertex | reader tile << 16, reader's record address}
  uint2* s_perm = s_plist + FBT_PCAP;                            // [ECAP] edge slot -> {edge id, source vertex | offset << 10}
  int* s_lrow = reinterpret_cast<int*>(s_perm + FBT_ECAP);       // [VCAP+1] slot base
  int* s_erow = s_lrow + FBT_VCAP + 1;                           // [VCAP+1] own-edge base
  int* s_first = s_erow + FBT_VCAP + 1;                          // [VCAP] first out-edge (global id)
  int* s_nin = s_first + FBT_VCAP;                               // [VCAP] in-degree
  __shared__ int s_warp[32];
  __shared__ __align__(8) uint64_t s_mbar[2];                    // A, B
  __shared__ int s_nrin;                                         // incidences filled by other tiles
  __shared__ unsigned s_nhalo, s_npush;                          // halo records reserved here / push entries registered here
  __shared__ unsigned s_nrem, s_cloc, s_crem;                    // out-edges with a target in another tile; slot counters of the two classes
  const GraphView& g = a.g;
  const int tid = threadIdx.x;
  const int r = (int)fbc_cluster_ctarank();
  const int s = (g.only >= 0) ? g.only : (int)blockIdx.x / FBT_C;
  const int nV = g.nV[s];
  if (nV == 0) return;  // uniform over the cluster
  const size_t vb = (size_t)s * g.maxV, eb = (size_t)s * g.maxE;
  const int32_t* row = g.row + (size_t)s * (g.maxV + 1);
  const int32_t* inc = g.inc + 2 * eb;
  const int32_t* vtile = a.vtile + vb;
  const int32_t* vloc = a.vloc + vb;
  const int32_t* toff = a.toff + (size_t)s * (FBT_C + 1);
  const int base = toff[r], nOwn = toff[r + 1] - base;
  const int32_t* tl_ = a.tlist + vb + base;
  int32_t* g_lrow = a.lrow + vb;
  if (tid == 0) { s_nrin = 0; s_nhalo = 0u; s_npush = 0u; s_nrem = 0u; s_cloc = 0u; s_crem = 0u; }
  __syncthreads();

  // ---- prologue 1: own vertices (state into registers, CSR bookkeeping into shared memory)
  float vx[FBT_VPT], vw1[FBT_VPT], vw2[FBT_VPT], vz[FBT_VPT], vth[FBT_VPT];
  int vid[FBT_VPT], vdeg[FBT_VPT], vrow[FBT_VPT];
  int carry_s = 0, carry_e = 0;
  const bool fits_v = nOwn <= FBT_VCAP;
#pragma unroll
  for (int k = 0; k < FBT_VPT; ++k) {
    const int lv = tid + k * FBT_THREADS;
    vid[k] = -1; vdeg[k] = 0; vrow[k] = 0;
    vx[k] = vw1[k] = vw2[k] = vz[k] = vth[k] = 0.f;
    int od = 0, nin = 0, first = 0;
    if (fits_v && lv < nOwn) {
      const int v = tl_[lv];
      vid[k] = v;
      const int r0 = row[v], r1 = row[v + 1];
      vdeg[k] = r1 - r0;
      int nrem = 0;
      nin = g.vnin[vb + v];  // in-edges come first in the row (ascending edge id)
      for (int j = 0; j < nin; ++j) {  // independent loads: the tiles of the in-edges' sources
        const int ts = vtile[g.eij[eb + (inc[r0 + j] >> 1)].x];
        if (ts != r) ++nrem;
      }
      if (nrem) atomicAdd(&s_nrin, nrem);
      od = vdeg[k] - nin;
      first = od ? (inc[r0 + nin] >> 1) : 0;
      vx[k] = g.x[vb + v]; vw1[k] = g.w1[vb + v]; vw2[k] = g.w2[vb + v];
      vz[k] = g.z[vb + v];
      vth[k] = tl * g.wt[vb + v];
      s_bar[lv] = g.vbar[vb + v];
      s_first[lv] = first;
      s_nin[lv] = nin;
    }
    int tot_s, tot_e;
    // a vertex's slot block holds an odd number of 16-byte records: the gathers of a warp, whose
    // lanes read their blocks at a stride of the (mostly equal) block size, then spread over all banks
    const int ps = carry_s + fbt_block_scan(vdeg[k] ? (vdeg[k] | 1) : 0, s_warp, &tot_s);
    const int pe = carry_e + fbt_block_scan(od, s_warp, &tot_e);
    if (fits_v && lv < nOwn) {
      s_lrow[lv] = ps;
      s_erow[lv] = pe;
      vrow[k] = ps;
      g_lrow[vid[k]] = ps;
    }
    carry_s += tot_s;
    carry_e += tot_e;
  }
  const int nSlot = carry_s, nEdge = carry_e;
  if (tid == 0 && fits_v) { s_lrow[nOwn] = nSlot; s_erow[nOwn] = nEdge; }
  if (tid == 0 && (!fits_v || nSlot > FBT_SLOTCAP || nEdge > FBT_ECAP)) {
    atomicOr(a.derr, 1);
    *reinterpret_cast<volatile int*>(a.err) = 1;
  }
  __threadfence();
  fbc_cluster_sync();  // every tile's slot bases (global) and capacity verdict are visible
  if (__ldcg(a.derr) != 0) return;  // uniform: nothing is solved once a tile did not fit (sticky, reported by the host)

  // ---- prologue 2: own edges = the contiguous out-edge ranges of the own vertices
  float q1[FBT_EPT], q2[FBT_EPT], q3[FBT_EPT], ea[FBT_EPT], ebt[FBT_EPT], dx[FBT_EPT], dy[FBT_EPT];
  uint32_t a_bj[FBT_EPT], a_sj[FBT_EPT], pk[FBT_EPT];  // pk: lv | own slot << 10 | target tile << 23 | remote << 27
  const uint32_t bar_u32 = fbc_smem_u32(s_bar), slot_u32 = fbc_smem_u32(s_slot), halo_u32 = fbc_smem_u32(s_halo);
  const uint32_t plist_u32 = fbc_smem_u32(s_plist), npush_u32 = fbc_smem_u32(&s_npush);
  const uint32_t mbA = fbc_smem_u32(&s_mbar[0]), mbB = fbc_smem_u32(&s_mbar[1]);
  const int kmax = (nEdge + FBT_THREADS - 1) / FBT_THREADS;  // uniform over the CTA
  // Edge slots (thread tid, round k) are handed out by class: a thread's round 0 holds edges whose target
  // is in this tile, the edges into OTHER tiles follow from slot min(512, #local) on, the remaining local
  // edges last.  In the dual half-step the first round then needs no halo (it runs while the owners'
  // pushes are still in flight), and the contributions to other tiles are sent before the last round is
  // computed (they fly meanwhile).  Which local edge lands in which slot is decided by atomics; every
  // edge's arithmetic and every slot address is independent of it.
#pragma unroll 1
  for (int le = tid; le < nEdge; le += FBT_THREADS) {
    int lv, off;
    fbt_edge_of(s_erow, nOwn, le, &lv, &off);
    const int e = s_first[lv] + off;
    const bool rem = vtile[g.eij[eb + e].y] != r;
    s_perm[le] = make_uint2((uint32_t)e | (rem ? 0x80000000u : 0u), (uint32_t)lv | ((uint32_t)off << 10));  // natural order, class in bit 31
    const unsigned m = __ballot_sync(__activemask(), rem);
    if (rem && (m & ((1u << (tid & 31)) - 1u)) == 0u) atomicAdd(&s_nrem, (unsigned)__popc(m));
  }
  __syncthreads();
  const int nRem = (int)s_nrem, nLoc = nEdge - nRem;
  const int remBase = nLoc < FBT_THREADS ? nLoc : FBT_THREADS;  // first slot of the remote class
  const int kB = remBase / FBT_THREADS;                         // the round in which the halo is first read
  uint2 mine[FBT_EPT];
#pragma unroll
  for (int k = 0; k < FBT_EPT; ++k) {
    const int le = tid + k * FBT_THREADS;
    mine[k] = le < nEdge ? s_perm[le] : make_uint2(0u, 0u);
  }
  __syncthreads();  // the natural-order records are in registers: the table is rewritten in slot order
#pragma unroll
  for (int k = 0; k < FBT_EPT; ++k) {
    const int le = tid + k * FBT_THREADS;
    if (le < nEdge) {
      int p;
      if (mine[k].x & 0x80000000u) {
        p = remBase + (int)atomicAdd(&s_crem, 1u);
      } else {
        const int i = (int)atomicAdd(&s_cloc, 1u);
        p = i < remBase ? i : i + nRem;
      }
      s_perm[p] = make_uint2(mine[k].x & 0x7fffffffu, mine[k].y);
    }
  }
  __syncthreads();
  uint32_t evalid = 0u;
#pragma unroll
  for (int k = 0; k < FBT_EPT; ++k) {
    q1[k] = q2[k] = q3[k] = ea[k] = ebt[k] = dx[k] = dy[k] = 0.f;
    a_bj[k] = a_sj[k] = pk[k] = 0u;
    const int le = tid + k * FBT_THREADS;
    if (le < nEdge) {
      const uint2 pe = s_perm[le];
      const int e = (int)pe.x, lv = (int)(pe.y & 1023u), off = (int)(pe.y >> 10);
      const int2 ij = g.eij[eb + e];
      const float4 c = g.ec[eb + e];
      const float4 q = g.q4[eb + e];
      evalid |= 1u << k;
      ea[k] = c.x; ebt[k] = c.y; dx[k] = c.z; dy[k] = c.w;
      q1[k] = q.x; q2[k] = q.y; q3[k] = q.z;
      const int j = ij.y, tj = vtile[j], lj = vloc[j];
      const int pos = g.epos[eb + e];  // position of e among j's incidences
      // a target in this tile is addressed in the CTA's own window (plain ld/st.shared: full shared-memory
      // bandwidth); only targets in other tiles go through the cluster window (DSMEM, ~20 B/clk per SM)
      a_bj[k] = bar_u32 + 16u * (uint32_t)lj;
      a_sj[k] = slot_u32 + 16u * (uint32_t)(__ldcg(g_lrow + j) + pos);
      if (tj != r) {
        a_sj[k] = fbc_mapa(a_sj[k], (uint32_t)tj);
        // the target's point is read from a halo record of THIS tile; its owner is told where to push it
        const uint32_t h = atomicAdd(&s_nhalo, 1u);
        a_bj[k] = halo_u32 + 16u * (h < FBT_HCAP ? h : 0u);
        if (h < FBT_HCAP) {
          const uint32_t idx = fbt_atom_add_cluster(fbc_mapa(npush_u32, (uint32_t)tj), 1u);
          if (idx < FBT_PCAP) {
            fbt_st_cluster_v2(fbc_mapa(plist_u32 + 8u * idx, (uint32_t)tj), (uint32_t)lj | ((uint32_t)r << 16), a_bj[k]);
          } else {  // the owner's push list is full: capacity verdict, nothing is solved
            atomicOr(a.derr, 1);
            *reinterpret_cast<volatile int*>(a.err) = 1;
          }
        }
      }
      pk[k] = (uint32_t)lv | ((uint32_t)(s_lrow[lv] + s_nin[lv] + off) << 10) | ((uint32_t)tj << 23) | (tj != r ? 1u << 27 : 0u);
    }
  }
  __syncthreads();
  const int nRin = s_nrin;
  const uint32_t nHalo = s_nhalo;
  const bool hasA = nRin > 0, hasB = nHalo > 0u;
  if (tid == 0) {
    fbc_mbar_init(mbA, 1);
    fbc_mbar_init(mbB, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (nHalo > FBT_HCAP) {  // capacity verdict of this tile's halo (the owners check their push lists below)
      atomicOr(a.derr, 1);
      *reinterpret_cast<volatile int*>(a.err) = 1;
    }
    if (hasB && nHalo <= FBT_HCAP) fbc_mbar_expect(mbB, 16u * nHalo);  // phase 0 of B: the initial points
  }
  __threadfence();
  fbc_cluster_sync();  // every tile's s_bar is filled, its barriers exist and its push list is complete
  if (__ldcg(a.derr) != 0) return;  // uniform over the cluster: a halo or a push list did not fit
  const uint32_t nPush = s_npush;
  // initial halo: the owners push the points iteration 0 reads
  for (uint32_t i = tid; i < nPush; i += FBT_THREADS) {
    const uint2 e = s_plist[i];
    fbc_st_async(fbc_mapa(e.y, e.x >> 16), s_bar[e.x & 0xffffu], fbc_mapa(mbB, e.x >> 16));
  }

  for (int it = 0; it < iters; ++it) {
    if (hasA && tid == 0) fbc_mbar_expect(mbA, 16u * (uint32_t)nRin);  // phase `it` of A
    // ---- dual half-step (nltgv2.cuh:k_dual_edges) + K^T q into the CSR slots of both endpoints
#pragma unroll
    for (int k = 0; k < FBT_EPT; ++k) {
      if (hasB && k == kB) fbc_mbar_wait(mbB, (uint32_t)(it & 1));  // the halo of this iteration has landed (first round that reads it)
      if (k < kmax && (evalid >> k) & 1u) {
        const float4 bik = fbc_lds(bar_u32 + ((pk[k] & 1023u) << 4));
        const float4 bjk = fbc_lds(a_bj[k]);  // own tile: s_bar; other tile: the halo record its owner refreshed
        float t = bik.x - bjk.x;
        t = fmaf(-dx[k], bik.y, t);
        t = fmaf(-dy[k], bik.z, t);
        const float k1 = ea[k] * t;
        const float k2 = ebt[k] * (bik.y - bjk.y);
        const float k3 = ebt[k] * (bik.z - bjk.z);
        q1[k] = fb_clamp1(fmaf(sigma, k1, q1[k]));
        q2[k] = fb_clamp1(fmaf(sigma, k2, q2[k]));
        q3[k] = fb_clamp1(fmaf(sigma, k3, q3[k]));
        const float a1 = ea[k] * q1[k];
        s_slot[(pk[k] >> 10) & 8191u] = make_float4(a1, fmaf(ebt[k], q2[k], -(dx[k] * a1)), fmaf(ebt[k], q3[k], -(dy[k] * a1)), 0.f);
        const float4 ct = make_float4(-a1, -(ebt[k] * q2[k]), -(ebt[k] * q3[k]), 0.f);
        if (pk[k] >> 27) fbc_st_async(a_sj[k], ct, fbc_mapa(mbA, (pk[k] >> 23) & 15u));
        else fbc_sts(a_sj[k], ct);
      }
    }
    __syncthreads();                                   // the contributions produced in this tile
    // every thread is past this iteration's halo wait: the next phase of B may be armed
    if (hasB && tid == 0 && it + 1 < iters) fbc_mbar_expect(mbB, 16u * nHalo);
    if (hasA) fbc_mbar_wait(mbA, (uint32_t)(it & 1));  // ... and those delivered by the other tiles
    // ---- primal half-step (nltgv2.cuh:k_primal_vertices): CSR-order sum, prox, box, extragradient
#pragma unroll
    for (int k = 0; k < FBT_VPT; ++k)
      if (vid[k] >= 0) {
        float gx = 0.f, g1 = 0.f, g2 = 0.f;
        const float4* sl = s_slot + vrow[k];
        for (int j = 0; j < vdeg[k]; ++j) {
          const float4 c = sl[j];
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
        const float4 nb = make_float4(fmaf(theta, xn - xo, xn), fmaf(theta, w1n - w1o, w1n), fmaf(theta, w2n - w2o, w2n), 0.f);
        s_bar[tid + k * FBT_THREADS] = nb;
        if (it + 1 == iters) g.vbar[vb + vid[k]] = nb;
      }
    __syncthreads();  // the tile's points are refreshed, its slots are consumed
    if (it + 1 < iters) {  // hand the refreshed points to the tiles that read them
      for (uint32_t i = tid; i < nPush; i += FBT_THREADS) {
        const uint2 e = s_plist[i];
        fbc_st_async(fbc_mapa(e.y, e.x >> 16), s_bar[e.x & 0xffffu], fbc_mapa(mbB, e.x >> 16));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < FBT_EPT; ++k)
    if ((evalid >> k) & 1u) g.q4[eb + s_perm[tid + k * FBT_THREADS].x] = make_float4(q1[k], q2[k], q3[k], 0.f);
#pragma unroll
  for (int k = 0; k < FBT_VPT; ++k)
    if (vid[k] >= 0) {
      const size_t v = vb + vid[k];
      g.x[v] = vx[k]; g.w1[v] = vw1[k]; g.w2[v] = vw2[k];
    }
  fbc_cluster_sync();  // no tile retires while a peer could still address its shared memory
}
