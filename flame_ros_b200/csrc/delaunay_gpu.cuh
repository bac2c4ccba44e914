// delaunay_gpu.cuh -- the `sync_graph` + `triangulate` stages of flame::Flame::update on the device
// (timing keys /root/reference/src/utils.cc:151-154; gates regularization/nltgv2/idepth_var_max and
// friends, /root/reference/src/flame_nodelet.cc:249-263).
//
// Why on the device: on the C2 stream every frame adds ~30 and removes ~25 of ~6k vertices and changes
// ~2 % of the edges (scripts/topology_churn.py), so a topology can never be kept from one frame to the
// next; the host triangulator cost 2.7 ms per frame plus two stream synchronisations and the upload
// of the rebuilt graph.  Here the whole chain runs as six small launches with no host involvement:
//   k_ds_stash    previous graph state (x, w, xbar, q, edge lists) aside for the carry-over
//   k_ds_prepare  vertex selection (projected live features below the variance gate, in ascending
//                 feature index), lattice coordinates, bounding box, cell grid, counting sort by
//                 cell, duplicate marking -- one CTA, everything in shared memory
//   k_ds_stars    one warp per vertex: its Delaunay star (delaunay_star.h), exact predicates
//   k_ds_scan     prefix sums of out-degree / started triangles / degree -> edge, triangle and CSR
//                 offsets, the device-side counts (nV, nE, nT), symmetry check
//   k_ds_emit     canonical edges (i<j, sorted by (i,j)) with alpha = 1/|delta|, beta = 1, delta;
//                 canonical triangles (smallest vertex first, sorted)
//   k_ds_csr      CSR incidence in ascending edge id, data term, and the carry-over of (x, w, q)
//                 from the previous graph by feature identity (new vertices start at the dense
//                 prediction or their data term, new edges at q = 0)
// The mesh is bit-identical to the host triangulator's canonical output (tests/test_delaunay_gpu.py).
#pragma once

#include "common.cuh"
#include "delaunay_star.h"
#include "nltgv2.cuh"

#define DSG_MAXCELLS 4096
#define DSG_THREADS 1024

// meta record of one stream (device ints)
enum { DSG_GX = 0, DSG_GY, DSG_SHIFT, DSG_BX0, DSG_BY0, DSG_BX1, DSG_BY1, DSG_NV, DSG_NE, DSG_NT, DSG_ERR,
       DSG_SUMDEG, DSG_OLD_NV, DSG_OLD_NE, DSG_HAVE, DSG_WORK, DSG_META };

struct DelGpu {
  DsPt* vxy = nullptr;          // [S*maxV] lattice coordinates by vertex
  DsPt* sxy = nullptr;          // [S*maxV] cell-sorted
  int32_t* sid = nullptr;       // [S*maxV]
  int32_t* vorder = nullptr;    // [S*maxV] processing order of k_ds_stars: vertices of border cells first
  int32_t* cell_start = nullptr;  // [S*(DSG_MAXCELLS+1)]
  int32_t* star = nullptr;      // [S*maxV*DS_MAXD]
  int32_t* deg = nullptr;       // [S*maxV] degree | closed << 8
  int32_t* od = nullptr;        // [S*maxV] out-degree
  int32_t* tc = nullptr;        // [S*maxV] triangles started
  int32_t* eoff = nullptr;      // [S*(maxV+1)]
  int32_t* toff = nullptr;      // [S*(maxV+1)]
  int32_t* meta = nullptr;      // [S*DSG_META]
  int32_t* f2v = nullptr;       // [2][S*maxF] feature -> vertex of the current / previous graph (by parity)
  // previous graph, for the carry-over
  float* o_x = nullptr; float* o_w1 = nullptr; float* o_w2 = nullptr;
  float4* o_vbar = nullptr; float4* o_q4 = nullptr;
  int2* o_eij = nullptr; int32_t* o_eoff = nullptr;
};

// ------------------------------------------------------------------------------------ block scan
// Exclusive scan of one int per thread over the block; returns the exclusive prefix, *total the sum.
__device__ __forceinline__ int dsg_block_scan(int v, int* s_warp, int* total) {
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

// The same for a 64-bit value (three packed 21-bit counters).
__device__ __forceinline__ unsigned long long dsg_block_scan64(unsigned long long v, unsigned long long* s_warp, unsigned long long* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    unsigned long long w = lane < nw ? s_warp[lane] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  *total = s_warp[nw - 1];
  return (wid ? s_warp[wid - 1] : 0ull) + incl - v;
}

// ------------------------------------------------------------------------------------ k_ds_stash
struct DsgGraph {
  float* x; float* w1; float* w2; float4* vbar; float4* q4; int2* eij;
};
__global__ void __launch_bounds__(256)
k_ds_stash(DsgGraph g, DelGpu d, int s, int maxV, int maxE) {
  int32_t* meta = d.meta + (size_t)s * DSG_META;
  // counts of the PREVIOUS graph, left by k_ds_prepare (which has already stored the new vertex count)
  const int V = meta[DSG_OLD_NV], E = meta[DSG_OLD_NE];
  const size_t vb = (size_t)s * maxV, eb = (size_t)s * maxE, ob = (size_t)s * (maxV + 1);
  const int n = blockDim.x * gridDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int v = t0; v < V; v += n) {
    d.o_x[vb + v] = g.x[vb + v];
    d.o_w1[vb + v] = g.w1[vb + v];
    d.o_w2[vb + v] = g.w2[vb + v];
    d.o_vbar[vb + v] = g.vbar[vb + v];
  }
  for (int v = t0; v <= V; v += n) d.o_eoff[ob + v] = d.eoff[ob + v];
  for (int e = t0; e < E; e += n) {
    d.o_q4[eb + e] = g.q4[eb + e];
    d.o_eij[eb + e] = g.eij[eb + e];
  }
}

// ------------------------------------------------------------------------------------ k_ds_prepare
struct DsgSelect {
  const float2* u_cur;      // [maxF] projected position in the current frame
  const float* var_cur;     // [maxF]
  const int32_t* valid;     // [maxF]
  float var_max;
  int maxF, maxV, W, H;
  // height band (regularization/nltgv2/{min,max}_height): world z of the feature's 3-D point
  int use_height;
  const float* mu_cur;      // [maxF] projected idepth
  const float* K;           // [9] intrinsics (device)
  const float* pose_cur;    // [7] current camera-in-world pose (device: read per frame, so a captured frame stays valid)
  float hmin, hmax;
};
// World z of the 3-D point of a feature seen at pixel u with inverse depth mu in the current camera
// (fp32, fixed expression order: the oracle evaluates the same expressions).
__host__ __device__ __forceinline__ float dsg_world_height(const float* K, const float* q, float ux, float uy, float mu) {
  const float qx = q[0], qy = q[1], qz = q[2], qw = q[3];
  const float n = qx * qx + qy * qy + qz * qz + qw * qw;
  const float s2 = 2.0f / n;
  const float r20 = qx * qz * s2 - qw * qy * s2;
  const float r21 = qy * qz * s2 + qw * qx * s2;
  const float r22 = 1.0f - (qx * qx * s2 + qy * qy * s2);
  const float zc = 1.0f / mu;
  const float xc = ((ux - K[2]) / K[0]) * zc, yc = ((uy - K[5]) / K[4]) * zc;
  return fmaf(r20, xc, fmaf(r21, yc, fmaf(r22, zc, q[6])));
}
// One CTA.  vfeat / vpos / f2v_new are the stream's slices.
//
// Every phase is a pass over <= 64k items by 1024 threads.  Written naively (load, use, next item)
// a pass costs items/1024 memory latencies and the kernel is bound by nothing else (46 us at 5.6k
// vertices); so each pass takes its items DSG_B at a time: all loads of a batch are issued before
// the first use, and the items are strided by the block size so that a warp's loads coalesce.
#define DSG_B 8
#define DSG_WORDS 2048  // ballot words: 65536 items
__device__ __forceinline__ int dsg_select_flag(const DsgSelect& q, int f, int valid, float var) {
  int flag = (valid && var < q.var_max) ? 1 : 0;
  if (flag && q.use_height) {
    const float2 u = q.u_cur[f];
    const float h = dsg_world_height(q.K, q.pose_cur, u.x, u.y, q.mu_cur[f]);
    if (!(h >= q.hmin && h <= q.hmax)) flag = 0;
  }
  return flag;
}
__global__ void __launch_bounds__(DSG_THREADS)
k_ds_prepare(DsgSelect q, DelGpu d, int s, int32_t* vfeat, float2* vpos, int32_t* f2v_new, int32_t* nV_out) {
  __shared__ int s_warp[32];
  __shared__ int s_cnt[DSG_MAXCELLS];
  __shared__ unsigned s_word[DSG_WORDS];
  __shared__ int s_wpre[DSG_WORDS];
  __shared__ int s_box[4];
  __shared__ int s_bad;
  DSG_CLK_DECL
  DSG_CLK
  const int tid = threadIdx.x, lane = tid & 31;
  int32_t* meta = d.meta + (size_t)s * DSG_META;
  const size_t vb = (size_t)s * q.maxV;
  DsPt* vxy = d.vxy + vb;
  DsPt* sxy = d.sxy + vb;
  int32_t* sid = d.sid + vb;
  int32_t* cell_start = d.cell_start + (size_t)s * (DSG_MAXCELLS + 1);
  if (tid == 0) {
    s_box[0] = s_box[1] = 0x7fffffff;
    s_box[2] = s_box[3] = -0x7fffffff;
    s_bad = 0;
    // counts of the graph this one replaces, for k_ds_stash / k_ds_csr (NV is overwritten below)
    meta[DSG_OLD_NV] = meta[DSG_HAVE] ? meta[DSG_NV] : 0;
    meta[DSG_OLD_NE] = meta[DSG_HAVE] ? meta[DSG_NE] : 0;
  }
  // ---- selection in ascending feature index: a warp's 32 flags become one ballot word, a block scan
  // over the words' popcounts ranks them (rounds of 65536 features)
  int carry = 0;
  {
    int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
    for (int base = 0; base < q.maxF; base += 32 * DSG_WORDS) {
      __syncthreads();
      const int nf = min(q.maxF - base, 32 * DSG_WORDS);
      for (int b0 = 0; b0 < nf; b0 += DSG_B * DSG_THREADS) {
        int va[DSG_B];
        float vr[DSG_B];
#pragma unroll
        for (int k = 0; k < DSG_B; ++k) {
          const int f = b0 + k * DSG_THREADS + tid;
          va[k] = 0; vr[k] = 0.f;
          if (f < nf) { va[k] = q.valid[base + f]; vr[k] = q.var_cur[base + f]; }
        }
#pragma unroll
        for (int k = 0; k < DSG_B; ++k) {
          const int f = b0 + k * DSG_THREADS + tid;
          const int flag = f < nf ? dsg_select_flag(q, base + f, va[k], vr[k]) : 0;
          const unsigned word = __ballot_sync(0xffffffffu, flag);
          if (lane == 0 && (f & ~31) < nf) s_word[f >> 5] = word;
        }
      }
      __syncthreads();
      const int nw = (nf + 31) >> 5;
      DSG_CLK  // selection: flags
      {  // exclusive prefix of the words' popcounts: two consecutive words per thread
        const int w0 = 2 * tid, c0 = w0 < nw ? __popc(s_word[w0]) : 0, c1 = w0 + 1 < nw ? __popc(s_word[w0 + 1]) : 0;
        int tot;
        const int ex = dsg_block_scan(c0 + c1, s_warp, &tot);
        s_wpre[w0] = carry + ex;
        s_wpre[w0 + 1] = carry + ex + c0;
        __syncthreads();
        carry += tot;
      }
      DSG_CLK  // selection: word scan
      for (int b0 = 0; b0 < nf; b0 += DSG_B * DSG_THREADS) {
        float2 u[DSG_B];
        int rank[DSG_B];
#pragma unroll
        for (int k = 0; k < DSG_B; ++k) {
          const int f = b0 + k * DSG_THREADS + tid;
          rank[k] = -1;
          u[k] = make_float2(0.f, 0.f);
          if (f < nf) {
            const unsigned word = s_word[f >> 5];
            if ((word >> lane) & 1u) {
              const int r = s_wpre[f >> 5] + __popc(word & ((1u << lane) - 1u));
              if (r < q.maxV) { rank[k] = r; u[k] = q.u_cur[base + f]; }
            }
          }
        }
#pragma unroll
        for (int k = 0; k < DSG_B; ++k) {
          const int f = b0 + k * DSG_THREADS + tid;
          if (f >= nf) continue;
          const int v = rank[k];
          if (v >= 0) {
            int bad = 0;
            DsPt l;
            l.x = ds_lattice(u[k].x, &bad);
            l.y = ds_lattice(u[k].y, &bad);
            if (bad) s_bad = 1;
            vfeat[v] = base + f;
            vpos[v] = u[k];
            vxy[v] = l;
            bx0 = min(bx0, l.x); by0 = min(by0, l.y);
            bx1 = max(bx1, l.x); by1 = max(by1, l.y);
          }
          f2v_new[base + f] = v;
        }
      }
    }
    bx0 = __reduce_min_sync(0xffffffffu, bx0); by0 = __reduce_min_sync(0xffffffffu, by0);
    bx1 = __reduce_max_sync(0xffffffffu, bx1); by1 = __reduce_max_sync(0xffffffffu, by1);
    if (lane == 0) {
      atomicMin(&s_box[0], bx0); atomicMin(&s_box[1], by0);
      atomicMax(&s_box[2], bx1); atomicMax(&s_box[3], by1);
    }
    __syncthreads();
  }
  const int V = carry < q.maxV ? carry : q.maxV;
  DSG_CLK  // selection
  // ---- cell grid: ~2 mean spacings per cell, at most DS_MAXROWS rows and DSG_MAXCELLS cells
  int shift = 9, gx = 1, gy = 1;
  if (V > 0) {
    const float spacing = sqrtf((float)q.W * (float)q.H / (float)V);
    while ((float)(1 << (shift - 6)) < 2.0f * spacing && shift < 20) ++shift;
    const int mx = s_box[2] > 0 ? s_box[2] : 0, my = s_box[3] > 0 ? s_box[3] : 0;
    for (;; ++shift) {
      gx = (mx >> shift) + 1;
      gy = (my >> shift) + 1;
      if (gy <= DS_MAXROWS && gx * gy <= DSG_MAXCELLS) break;
    }
  }
  const int cells = gx * gy;
  for (int c = tid; c < cells; c += DSG_THREADS) s_cnt[c] = 0;
  __syncthreads();
  DsIn in;
  in.gx = gx; in.gy = gy; in.shift = shift;
  // slot inside the cell (any order is fine: the stars do not depend on it); parked in `od`, which
  // k_ds_stars overwrites later
  int32_t* slot = d.od + vb;
  for (int b0 = 0; b0 < V; b0 += DSG_B * DSG_THREADS) {
    DsPt l[DSG_B];
#pragma unroll
    for (int k = 0; k < DSG_B; ++k) {
      const int v = b0 + k * DSG_THREADS + tid;
      l[k].x = l[k].y = 0;
      if (v < V) l[k] = vxy[v];
    }
#pragma unroll
    for (int k = 0; k < DSG_B; ++k) {
      const int v = b0 + k * DSG_THREADS + tid;
      if (v >= V) continue;
      const int c = ds_celly(in, l[k].y) * gx + ds_cellx(in, l[k].x);
      slot[v] = atomicAdd(&s_cnt[c], 1);
    }
  }
  __syncthreads();
  // ---- exclusive scan of the cell counts (in place): a run of cells per thread, one block scan
  DSG_CLK  // cell counts
  {
    const int CI = (cells + DSG_THREADS - 1) / DSG_THREADS;
    const int c0 = tid * CI, c1 = min(cells, c0 + CI);
    int cnt = 0;
    for (int c = c0; c < c1; ++c) cnt += s_cnt[c];
    int tot;
    int ex = dsg_block_scan(cnt, s_warp, &tot);
    for (int c = c0; c < c1; ++c) {
      const int n = s_cnt[c];
      s_cnt[c] = ex;
      cell_start[c] = ex;
      ex += n;
    }
    __syncthreads();
  }
  if (tid == 0) cell_start[cells] = V;
  DSG_CLK  // cell scan
  // ---- scatter (slot -> sorted position); on the way, the ballot words of "vertex lies in a border cell"
  for (int b0 = 0; b0 < V; b0 += DSG_B * DSG_THREADS) {
    DsPt l[DSG_B];
    int sl[DSG_B];
#pragma unroll
    for (int k = 0; k < DSG_B; ++k) {
      const int v = b0 + k * DSG_THREADS + tid;
      l[k].x = l[k].y = 0; sl[k] = 0;
      if (v < V) { l[k] = vxy[v]; sl[k] = slot[v]; }
    }
#pragma unroll
    for (int k = 0; k < DSG_B; ++k) {
      const int v = b0 + k * DSG_THREADS + tid;
      int border = 0;
      if (v < V) {
        const int cx = ds_cellx(in, l[k].x), cy = ds_celly(in, l[k].y);
        const int pos = s_cnt[cy * gx + cx] + sl[k];
        sxy[pos] = l[k];
        sid[pos] = v;
        border = (cx == 0 || cy == 0 || cx == gx - 1 || cy == gy - 1) ? 1 : 0;
      }
      const unsigned word = __ballot_sync(0xffffffffu, border);
      if (lane == 0 && (v & ~31) < V) s_word[v >> 5] = word;
    }
  }
  __syncthreads();
  // ---- duplicates: a point with an identical point of smaller index is left out (sid = ~id).  One
  DSG_CLK  // scatter
  // thread per sorted entry: its cell's entries are consecutive (one miss, then L1)
  for (int k = tid; k < V; k += DSG_THREADS) {
    const DsPt l = sxy[k];
    const int raw0 = sid[k], id = raw0 < 0 ? ~raw0 : raw0;
    const int c = ds_celly(in, l.y) * gx + ds_cellx(in, l.x);
    const int beg = s_cnt[c], end = (c + 1 < cells) ? s_cnt[c + 1] : V;
    bool dup = false;
    for (int m = beg; m < end; ++m) {
      const int raw = sid[m], im = raw < 0 ? ~raw : raw;   // the owner may be marking it right now: same id either way
      const DsPt o = sxy[m];
      if (im < id && o.x == l.x && o.y == l.y) dup = true;
    }
    if (dup) sid[k] = ~id;
  }
  __syncthreads();
  DSG_CLK  // duplicates
  // vertices and their neighbours scan long strips of border cells (10x the median work); started
  // first they overlap with the bulk instead of forming the kernel's tail.  Both classes in ascending
  // vertex index: ranks from the ballot words (border count | interior count << 16 in one scan).
  {
    int32_t* vorder = d.vorder + vb;
    const int nw = (V + 31) >> 5;
    const int w0 = 2 * tid;
    int c0 = 0, c1 = 0;
    if (w0 < nw) { const int nb = __popc(s_word[w0]), nv = min(32, V - 32 * w0); c0 = nb | ((nv - nb) << 16); }
    if (w0 + 1 < nw) { const int nb = __popc(s_word[w0 + 1]), nv = min(32, V - 32 * (w0 + 1)); c1 = nb | ((nv - nb) << 16); }
    int tot;
    const int ex = dsg_block_scan(c0 + c1, s_warp, &tot);  // V < 65536
    s_wpre[w0] = ex;
    s_wpre[w0 + 1] = ex + c0;
    __syncthreads();
    const int nborder = tot & 0xffff;
    for (int v = tid; v < V; v += DSG_THREADS) {
      const unsigned word = s_word[v >> 5];
      const int pre = s_wpre[v >> 5];
      const unsigned below = (1u << lane) - 1u;
      if ((word >> lane) & 1u) vorder[(pre & 0xffff) + __popc(word & below)] = v;
      else vorder[nborder + (pre >> 16) + __popc(~word & below)] = v;
    }
  }
  if (tid == 0) {
    meta[DSG_GX] = gx; meta[DSG_GY] = gy; meta[DSG_SHIFT] = shift;
    meta[DSG_BX0] = s_box[0]; meta[DSG_BY0] = s_box[1]; meta[DSG_BX1] = s_box[2]; meta[DSG_BY1] = s_box[3];
    meta[DSG_NV] = V;
    meta[DSG_WORK] = 0;
    meta[DSG_ERR] = s_bad ? 0x100 : 0;
    *nV_out = V;
  }
  DSG_CLK  // order
  DSG_CLK_PRINT("k_ds_prepare")
}

// ------------------------------------------------------------------------------------ k_ds_stars
#define DSG_WARPS 4  // vertices per CTA: one warp each (the candidate cache takes ~7 KB per warp)
#ifndef DSG_STARS_MINB
#define DSG_STARS_MINB 5  // resident CTAs per SM (96 registers per thread, ~50 B of spills): same single-camera frame time as 4
                           // (128 registers), +6 % with eight cameras sharing the GPU (scripts/gpu_ab_stars_occ.sh)
#endif
__global__ void __launch_bounds__(DSG_WARPS * 32, DSG_STARS_MINB)
k_ds_stars(DelGpu d, int s, int maxV) {
  __shared__ DsScratch s_scr[DSG_WARPS];
  int32_t* meta = d.meta + (size_t)s * DSG_META;
  const int V = meta[DSG_NV];
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t vb = (size_t)s * maxV;
  DsIn in;
  in.n = V;
  in.vxy = d.vxy + vb;
  in.gx = meta[DSG_GX]; in.gy = meta[DSG_GY]; in.shift = meta[DSG_SHIFT];
  in.cell_start = d.cell_start + (size_t)s * (DSG_MAXCELLS + 1);
  in.sxy = d.sxy + vb;
  in.sid = d.sid + vb;
  in.bx0 = meta[DSG_BX0]; in.by0 = meta[DSG_BY0]; in.bx1 = meta[DSG_BX1]; in.by1 = meta[DSG_BY1];
  // Warps fetch vertices from a shared counter (zeroed by k_ds_prepare): a star costs between 0.3x and
  // 10x the median, and with a fixed vertex per warp a CTA's slot stayed occupied until its slowest
  // warp was done (SMs 26 % idle over the launch).
  for (;;) {
  int item = 0;
  if (lane == 0) item = atomicAdd(&meta[DSG_WORK], 1);
  item = __shfl_sync(0xffffffffu, item, 0);
  if (item >= V) break;
  const int p = d.vorder[vb + item];
  int32_t* star = d.star + (vb + p) * DS_MAXD;
  int deg = 0, closed = 0, od = 0, tc = 0;
  // a duplicate point has no star: find its own entry flag through its cell
  bool dup = false;
  {
    const DsPt l = in.vxy[p];
    const int c = ds_celly(in, l.y) * in.gx + ds_cellx(in, l.x);
    for (int k = in.cell_start[c] + lane; k < in.cell_start[c + 1]; k += 32)
      if (in.sid[k] == ~p) dup = true;
    dup = DsW32::any(dup);
  }
#ifdef DSG_CLOCKS
  const long long clk0_ = clock64();
#endif
  if (!dup) {
    const int rc = ds_star<DsW32>(in, p, &s_scr[g], star, &deg, &closed);
    if (rc) {
      if (lane == 0) atomicOr(&meta[DSG_ERR], 1 << rc);
      deg = 0;
    }
    DsW32::sync();
    if (lane == 0) ds_counts(p, star, deg, closed, &od, &tc);
  }
#ifdef DSG_CLOCKS
  if (lane == 0 && clock64() - clk0_ > 40000) printf("star item %d vertex %d: %lld cycles, degree %d, closed %d\n", item, p, clock64() - clk0_, deg, closed);
#endif
  if (lane == 0) {
    // one record per vertex: degree | closed << 8 | out-degree << 16 | started triangles << 24 (all <= 32):
    // the scan that follows reads one array instead of three
    d.deg[vb + p] = deg | (closed << 8) | (od << 16) | (tc << 24);
  }
  }
}

// Persistent grid of the star kernel: four CTAs per SM (its register limit), never more than the vertices need.
static inline int dsg_stars_grid(int device, int maxV) {
  static int sms[64] = {0};
  int& n = sms[device & 63];
  if (n == 0 && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); n = 148; }
  const int need = (maxV + DSG_WARPS - 1) / DSG_WARPS;
  return need < DSG_STARS_MINB * n ? need : DSG_STARS_MINB * n;
}

// (out-degree, started triangles, degree) of a vertex record as three 21-bit counters for the packed scan
__device__ __forceinline__ unsigned long long dsg_scan_pack(int rec) {
  return (unsigned long long)((rec >> 16) & 0xff) | ((unsigned long long)((rec >> 24) & 0xff) << 21) | ((unsigned long long)(rec & 0xff) << 42);
}

// ------------------------------------------------------------------------------------ k_ds_scan
// One CTA: offsets of edges / triangles / CSR rows, counts, consistency.
__global__ void __launch_bounds__(DSG_THREADS)
k_ds_scan(DelGpu d, int s, int maxV, int maxE, int maxT, int32_t* row, int32_t* nE_out, int32_t* nT_out) {
  __shared__ unsigned long long s_warp[32];
  int32_t* meta = d.meta + (size_t)s * DSG_META;
  const int V = meta[DSG_NV];
  const size_t vb = (size_t)s * maxV, ob = (size_t)s * (maxV + 1);
  // every thread takes a run of consecutive vertices; (out-degree, triangles, degree) share one
  // 64-bit scan, 21 bits each (sums <= 6 V < 2^21 for V < 2^18)
  const int VI = (V + DSG_THREADS - 1) / DSG_THREADS;
  const int v0 = min(V, (int)threadIdx.x * VI), v1 = min(V, v0 + VI);
  unsigned long long acc = 0ull;
#pragma unroll 4
  for (int v = v0; v < v1; ++v)
    acc += dsg_scan_pack(d.deg[vb + v]);
  unsigned long long tot;
  unsigned long long ex = dsg_block_scan64(acc, s_warp, &tot);
#pragma unroll 4
  for (int v = v0; v < v1; ++v) {
    d.eoff[ob + v] = (int)(ex & 0x1fffffull);
    d.toff[ob + v] = (int)((ex >> 21) & 0x1fffffull);
    row[v] = (int)(ex >> 42);
    ex += dsg_scan_pack(d.deg[vb + v]);
  }
  const int ce = (int)(tot & 0x1fffffull), ct = (int)((tot >> 21) & 0x1fffffull), cr = (int)(tot >> 42);
  if (threadIdx.x == 0) {
    d.eoff[ob + V] = ce;
    d.toff[ob + V] = ct;
    row[V] = cr;
    int err = meta[DSG_ERR];
    if (cr != 2 * ce) err |= 0x200;              // asymmetric stars
    if (ce > maxE || ct > maxT) err |= 0x400;    // capacity
    const bool ok = err == 0 && ct > 0;
    meta[DSG_ERR] = err;
    meta[DSG_NE] = ok ? ce : 0;
    meta[DSG_NT] = ok ? ct : 0;
    meta[DSG_SUMDEG] = cr;
    meta[DSG_HAVE] = ok ? 1 : 0;
    *nE_out = ok ? ce : 0;
    *nT_out = ok ? ct : 0;
  }
}

// ------------------------------------------------------------------------------------ k_ds_emit
// One warp per vertex, lane k = star entry k (DS_MAXD == 32): the canonical rank of an out-edge /
// triangle is the number of smaller ones in the star (shuffles), and all gathers of a vertex are in
// flight together.  Same order as ds_emit: out-neighbours ascending, triangles (v, a, b) ascending a.
static_assert(DS_MAXD == 32, "k_ds_emit maps star entries to lanes");
__global__ void __launch_bounds__(128)
k_ds_emit(DelGpu d, int s, int maxV, int maxE, int maxT, const float2* __restrict__ vpos, int2* __restrict__ eij,
          float4* __restrict__ ec, int32_t* __restrict__ tri) {
  const int32_t* meta = d.meta + (size_t)s * DSG_META;
  const int V = meta[DSG_NV];
  const int v = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (v >= V || meta[DSG_NT] == 0) return;
  const size_t vb = (size_t)s * maxV, ob = (size_t)s * (maxV + 1);
  const int dg = d.deg[vb + v], deg = dg & 0xff, closed = (dg >> 8) & 1;
  if (deg == 0) return;
  const int e0 = d.eoff[ob + v], t0 = d.toff[ob + v];
  const float2 pv = vpos[v];
  const int w = lane < deg ? d.star[(vb + v) * DS_MAXD + lane] : -1;
  const int b = __shfl_sync(0xffffffffu, w, lane + 1 == deg ? 0 : ((lane + 1) & 31));  // next neighbour counter-clockwise
  const int pairs = closed ? deg : deg - 1;
  const bool is_out = w > v, is_tri = lane < pairs && w > v && b > v;
  int epos = 0, tpos = 0;
  for (int m = 0; m < deg; ++m) {
    const int u = __shfl_sync(0xffffffffu, w, m), ub = __shfl_sync(0xffffffffu, b, m);
    epos += (u > v && u < w) ? 1 : 0;
    tpos += (m < pairs && u > v && ub > v && u < w) ? 1 : 0;
  }
  if (is_out) {
    const float2 pw = vpos[w];
    // dx = pos_i - pos_j in fp32, alpha = 1/|delta|, beta = 1 (same expressions as the host path)
    const float dx = pv.x - pw.x, dy = pv.y - pw.y;
    eij[e0 + epos] = make_int2(v, w);
    ec[e0 + epos] = make_float4(1.0f / sqrtf(dx * dx + dy * dy), 1.0f, dx, dy);
  }
  if (is_tri) {
    int32_t* t = tri + 3 * (size_t)(t0 + tpos);
    t[0] = v; t[1] = w; t[2] = b;
  }
}

// ------------------------------------------------------------------------------------ k_ds_csr
struct DsgCarry {
  const float* mu_cur;      // [maxF] projected idepth of every feature
  const float* var_cur;     // [maxF]
  const int32_t* f2v_old;   // [maxF] feature -> vertex of the previous graph
  const float* idmap;       // previous dense map (prediction) or NULL
  int W, H, adaptive, use_prediction;
};
// Thread t serves vertex t (data term, state carry-over, in-degree) and edge t (dual carry-over, its
// two places in the CSR incidence: ascending edge id = in-edges by source, then out-edges).  Edge-
// parallel because the per-vertex version was one long chain of dependent loads per incident edge.
__global__ void __launch_bounds__(128)
k_ds_csr(DelGpu d, DsgCarry q, int s, int maxV, int maxE, const int32_t* __restrict__ vfeat,
         const float2* __restrict__ vpos, const int2* __restrict__ eij, const int32_t* __restrict__ row,
         int32_t* __restrict__ inc, float* z, float* wt, float* x, float* w1, float* w2, float4* vbar, float4* q4,
         int32_t* __restrict__ epos, int32_t* __restrict__ vnin) {
  int32_t* meta = d.meta + (size_t)s * DSG_META;
  const int V = meta[DSG_NV];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t vb = (size_t)s * maxV, eb = (size_t)s * maxE, ob = (size_t)s * (maxV + 1);
  const int oV = meta[DSG_OLD_NV], oE = meta[DSG_OLD_NE];
  const bool mesh = meta[DSG_NT] != 0;
  if (t < V) {
    // ---- data term + vertex state
    const int v = t;
    const int f = vfeat[v];
    const float zz = q.mu_cur[f];
    z[v] = zz;
    wt[v] = q.adaptive ? (1.0f / q.var_cur[f]) : 1.0f;
    const int ov = oV > 0 ? q.f2v_old[f] : -1;
    if (ov >= 0 && ov < oV) {
      x[v] = d.o_x[vb + ov];
      w1[v] = d.o_w1[vb + ov];
      w2[v] = d.o_w2[vb + ov];
      vbar[v] = d.o_vbar[vb + ov];
    } else {
      float x0 = zz;
      if (q.use_prediction && q.idmap) {
        const int px = (int)rintf(vpos[v].x), py = (int)rintf(vpos[v].y);
        if (px >= 0 && py >= 0 && px < q.W && py < q.H) {
          const float p = q.idmap[py * q.W + px];
          if (p == p && p > 0.0f) x0 = p;
        }
      }
      x[v] = x0;
      w1[v] = 0.0f;
      w2[v] = 0.0f;
      vbar[v] = make_float4(x0, 0.0f, 0.0f, 0.0f);
    }
    // in-degree = neighbours with a smaller index = degree - out-degree
    vnin[v] = mesh ? (d.deg[vb + v] & 0xff) - (d.eoff[ob + v + 1] - d.eoff[ob + v]) : 0;
  }
  if (!mesh || t >= meta[DSG_NE]) return;
  const int e = t;
  const int2 vw = eij[e];
  const int v = vw.x, w = vw.y;
  // ---- carry q over from the previous graph (both endpoints persisted, edge existed)
  float4 qq = make_float4(0.f, 0.f, 0.f, 0.f);
  if (oV > 0) {
    const int ov = q.f2v_old[vfeat[v]], ow = q.f2v_old[vfeat[w]];
    if (ov >= 0 && ov < oV && ow >= 0 && ow < oV) {
      const int b = d.o_eoff[ob + ov], en = d.o_eoff[ob + ov + 1];
      for (int k = b; k < en && k < oE; ++k)
        if (d.o_eij[eb + k].y == ow) {
          qq = d.o_q4[eb + k];
          break;
        }
    }
  }
  q4[e] = qq;
  // ---- out-edge of v: after v's in-edges, in edge order
  const int e0 = d.eoff[ob + v], odv = d.eoff[ob + v + 1] - e0;
  const int niv = (d.deg[vb + v] & 0xff) - odv;
  inc[row[v] + niv + (e - e0)] = e << 1;
  // ---- in-edge of w: its rank among w's smaller neighbours (ascending source)
  const int degw = d.deg[vb + w] & 0xff;
  const int32_t* __restrict__ star = d.star + (vb + w) * DS_MAXD;
  int k = 0;
  bool listed = false;
  for (int m = 0; m < degw; ++m) {
    const int u = star[m];
    k += u < v ? 1 : 0;
    listed = listed || u == v;
  }
  if (!listed) {
    atomicOr(&meta[DSG_ERR], 0x800);  // w does not list v
    return;
  }
  epos[e] = k;  // position of the edge in its target's row: the tile solver's slot address
  inc[row[w] + k] = (e << 1) | 1;
}

// ------------------------------------------------------------------------------------ rescale_data
// regularization/nltgv2/rescale_data ("Rescale data to have mean 1"): scale = mean(z), accumulated in
// double (exact for a few thousand fp32 values of similar magnitude, hence independent of the order).
__global__ void __launch_bounds__(DSG_THREADS)
k_rescale_mean(const int32_t* __restrict__ nV, int s, const float* __restrict__ z, float* __restrict__ scale) {
  __shared__ double s_sum[32];
  const int V = nV[s];
  double acc = 0.0;
  for (int v = threadIdx.x; v < V; v += blockDim.x) acc += (double)z[v];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_sum[k];
    float sc = V > 0 ? (float)(t / (double)V) : 1.0f;
    if (!(sc > 0.0f)) sc = 1.0f;
    *scale = sc;
  }
}
// dir = 0: keep z aside, divide z, x, w, xbar by the scale; dir = 1: multiply back, restore z.
__global__ void __launch_bounds__(256)
k_rescale_apply(const int32_t* __restrict__ nV, int s, int dir, const float* __restrict__ scale, float* z,
                float* z_keep, float* x, float* w1, float* w2, float4* vbar) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV[s]) return;
  const float sc = *scale;
  float4 b = vbar[v];
  if (dir == 0) {
    z_keep[v] = z[v];
    z[v] = z[v] / sc;
    x[v] = x[v] / sc; w1[v] = w1[v] / sc; w2[v] = w2[v] / sc;
    b.x = b.x / sc; b.y = b.y / sc; b.z = b.z / sc;
  } else {
    z[v] = z_keep[v];
    x[v] = x[v] * sc; w1[v] = w1[v] * sc; w2[v] = w2[v] * sc;
    b.x = b.x * sc; b.y = b.y * sc; b.z = b.z * sc;
  }
  vbar[v] = b;
}
