// delaunay_star.h -- per-vertex Delaunay stars: the core of the GPU `triangulate` stage.
//
// Stands in for the `triangulate` stage of flame::Flame::update (timing key
// /root/reference/src/utils.cc:154; the external flame core wraps Shewchuk's Triangle).  The host
// triangulator (delaunay.h) inserts points one by one into a shared structure -- inherently serial,
// 2.7 ms per frame at 5.9k vertices and every frame changes ~2 % of the edges, so nothing can be
// cached.  Here every vertex computes ITS OWN star (its Delaunay neighbours in counter-clockwise
// order) independently, one warp per vertex:
//   1. nearest neighbour n0 (always a Delaunay neighbour: its diametral circle is empty);
//   2. counter-clockwise sweep: given the Delaunay edge (p, cur), the next neighbour is the point c
//      strictly left of p->cur whose circle (p, cur, c) contains no other left point (a max under
//      the in-circle order: one pass over the candidates, lanes in parallel, shuffle reduction);
//   3. an open star (p on the convex hull) is completed by the mirrored clockwise sweep from n0.
// Candidates come from a uniform cell grid.  A pass over a block of cells is conclusive only when
// the block covers (circle or half-plane) INTERSECTED with the bounding box of all points; else the
// block grows to that region and the pass is repeated.  Hull edges and hull slivers therefore cost
// a strip of boundary cells, not the whole point set.
// Exactness: lattice coordinates (1/64 px), orientation in 64-bit integers, in-circle by a double
// filter with a 128-bit integer fallback -- the same predicates as delaunay.h, so both produce THE
// Delaunay triangulation.  Co-circular point sets (integer-pixel detections form exact rectangles)
// are triangulated canonically by both: the polygon of points on one empty circle becomes a fan
// from its smallest vertex index.  Seen from p with the edge (p, cur) and the tied set T on the
// sweep side: p smallest -> the member of T angularly closest to cur; cur smallest -> the farthest;
// otherwise the smallest index itself (which then is in T).
//
// The same source is compiled for the device (32 lanes) and, in tests/cpp/star_sim.cc, for the host
// with a 1-lane "warp" so the logic is checked against the host triangulator without a GPU.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define DS_FN __host__ __device__ __forceinline__
#define DS_MEM __host__ __device__ __forceinline__ static
#define DS_COLD __host__ __device__ __noinline__
#else
#define DS_FN static inline
#define DS_MEM static inline
#define DS_COLD static
#endif

#define DS_MAXD 32      // largest star (error above)
#define DS_MAXROWS 64   // grid rows (cells along y) at most
#define DS_COORD_LIM (1 << 20)

typedef __int128 ds_i128;

struct DsPt {
  int32_t x, y;
};

struct DsIn {
  int n;                      // vertices (including duplicates)
  const DsPt* vxy;            // [n] lattice coordinates by vertex id
  int gx, gy, shift;          // cell grid: cell = clamp(coord >> shift)
  const int32_t* cell_start;  // [gx*gy + 1] into the cell-sorted arrays
  const DsPt* sxy;            // [n] cell-sorted coordinates
  const int32_t* sid;         // [n] vertex id of the sorted entry, ~id for a duplicate point
  int bx0, by0, bx1, by1;     // bounding box of all points (lattice)
};

enum { DS_OK = 0, DS_E_DEGREE = 1, DS_E_TIE = 2, DS_E_LOOP = 3 };

// ------------------------------------------------------------------------------------ predicates
DS_FN int64_t ds_orient(DsPt a, DsPt b, DsPt c) {
  return (int64_t)(b.x - a.x) * (int64_t)(c.y - a.y) - (int64_t)(b.y - a.y) * (int64_t)(c.x - a.x);
}

// > 0 iff p lies strictly inside the circumcircle of the counter-clockwise triangle (a, b, c);
// 0 iff on it.  |coordinates| < 2^20: differences < 2^21, squared lengths and 2x2 minors < 2^43 are
// exact in double; the three-term sum is decided by a static error bound, else exactly in 128 bits.
// Stages 2 and 3 of the in-circle test, out of line: one copy in the kernel instead of one per call
// site (the fp32 stage decides nearly every candidate; the star kernel stalled on instruction fetch).
DS_COLD int ds_incircle_exact(int32_t iadx, int32_t iady, int32_t ibdx, int32_t ibdy, int32_t icdx, int32_t icdy) {
  {
    // stage 2, fp64: squared lengths and 2x2 minors are exact, only the three-term sum rounds
    const double adx = iadx, ady = iady, bdx = ibdx, bdy = ibdy, cdx = icdx, cdy = icdy;
    const double al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;  // exact
    const double ma = bdx * cdy - bdy * cdx, mb = cdx * ady - cdy * adx, mc = adx * bdy - ady * bdx;  // exact
    const double ta = al * ma, tb = bl * mb, tc = cl * mc;
    const double det = ta + tb + tc;
    const double bound = 8.9e-16 * (fabs(ta) + fabs(tb) + fabs(tc));
    if (det > bound) return 1;
    if (det < -bound) return -1;
  }
  // stage 3, exact: 128-bit integers (co-circular detections land here)
  const int64_t adx = iadx, ady = iady, bdx = ibdx, bdy = ibdy, cdx = icdx, cdy = icdy;
  const int64_t al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
  const int64_t ma = bdx * cdy - bdy * cdx, mb = cdx * ady - cdy * adx, mc = adx * bdy - ady * bdx;
  const ds_i128 d = (ds_i128)al * (ds_i128)ma + (ds_i128)bl * (ds_i128)mb + (ds_i128)cl * (ds_i128)mc;
  return d > 0 ? 1 : (d < 0 ? -1 : 0);
}
DS_FN int ds_incircle(DsPt a, DsPt b, DsPt c, DsPt p) {
  const int32_t iadx = a.x - p.x, iady = a.y - p.y, ibdx = b.x - p.x, ibdy = b.y - p.y;
  const int32_t icdx = c.x - p.x, icdy = c.y - p.y;
  {
    // stage 1, fp32: the differences are exact (< 2^24); every product and sum rounds.  The bound is
    // Shewchuk's "permanent" form: 10 eps (sum of the magnitudes of all terms), eps = 2^-24, rounded up
    // to 1e-6.  Nearly every candidate is far from the circle and is decided here at fp32 latency.
    const float adx = (float)iadx, ady = (float)iady, bdx = (float)ibdx, bdy = (float)ibdy, cdx = (float)icdx, cdy = (float)icdy;
    const float bc1 = bdx * cdy, bc2 = bdy * cdx, ca1 = cdx * ady, ca2 = cdy * adx, ab1 = adx * bdy, ab2 = ady * bdx;
    const float al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    const float det = al * (bc1 - bc2) + bl * (ca1 - ca2) + cl * (ab1 - ab2);
    const float perm = al * (fabsf(bc1) + fabsf(bc2)) + bl * (fabsf(ca1) + fabsf(ca2)) + cl * (fabsf(ab1) + fabsf(ab2));
    const float bound = 1e-6f * perm;
    if (det > bound) return 1;
    if (det < -bound) return -1;
  }
  return ds_incircle_exact(iadx, iady, ibdx, ibdy, icdx, icdy);
}

// dir = +1: counter-clockwise sweep (candidates strictly left of p->cur), -1: clockwise (right).
DS_FN bool ds_side(DsPt p, DsPt cur, DsPt c, int dir) {
  const int64_t o = ds_orient(p, cur, c);
  return dir > 0 ? o > 0 : o < 0;
}
// > 0: c strictly inside the circle (p, cur, b); 0: on it.  b and c are on the sweep side.
DS_FN int ds_inside(DsPt p, DsPt cur, DsPt b, DsPt c, int dir) {
  const DsPt u = dir > 0 ? cur : b, v = dir > 0 ? b : cur;  // (p, u, v) counter-clockwise; one in-circle call site
  return ds_incircle(p, u, v, c);
}

// ------------------------------------------------------------------------------------ lanes
#ifdef __CUDA_ARCH__
#define DS_F2U(f) __float_as_uint(f)
#else
static inline unsigned ds_f2u_host(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
#define DS_F2U(f) ds_f2u_host(f)
#endif
struct DsSeq {  // host simulation: one lane
  static const int LANES = 1;
  DS_MEM int lane() { return 0; }
  DS_MEM int shfl_xor(int v, int) { return v; }
  DS_MEM long long shfl_xor(long long v, int) { return v; }
  DS_MEM int bcast(int v, int) { return v; }
  DS_MEM bool any(bool p) { return p; }
  DS_MEM unsigned min_u32(unsigned v) { return v; }
  DS_MEM int first(bool) { return 0; }
  DS_MEM void sync() {}
};
#ifdef __CUDACC__
// One warp per vertex.  (Eight lanes per vertex, four vertices per warp, was measured and is 2x
// slower: the four sweeps diverge at every data-dependent branch and serialise.)
struct DsW32 {
  static const int LANES = 32;
  __device__ __forceinline__ static int lane() { return threadIdx.x & 31; }
  __device__ __forceinline__ static int shfl_xor(int v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
  __device__ __forceinline__ static long long shfl_xor(long long v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
  __device__ __forceinline__ static int bcast(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  __device__ __forceinline__ static bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
  __device__ __forceinline__ static unsigned min_u32(unsigned v) { return __reduce_min_sync(0xffffffffu, v); }  // redux.sync
  // lowest lane whose predicate holds (the caller guarantees there is one)
  __device__ __forceinline__ static int first(bool p) { return __ffs(__ballot_sync(0xffffffffu, p)) - 1; }
  __device__ __forceinline__ static void sync() { __syncwarp(); }
};
#endif

// ------------------------------------------------------------------------------------ cell blocks
#ifndef DS_CACHE
#define DS_CACHE 512  // candidates of a block held in the group's scratch (larger blocks are re-read, narrowed row by row)
#endif
#ifdef DS_STATS
extern long long ds_stat_visits, ds_stat_passes, ds_stat_iters32;
extern int ds_dbg_p;
#define DS_STAT_PASS(blk, S) { if (ds_dbg_p >= 0) printf("  pass block [%d..%d]x[%d..%d] rows %d cache %d\n", (blk).x0, (blk).x1, (blk).y0, (blk).y1, (blk).nrows, (blk).ncache); ++ds_stat_passes; int t_ = 0; if ((blk).ncache >= 0) t_ = ((blk).ncache + 31) / 32; else for (int r_ = 0; r_ < (blk).nrows; ++r_) t_ += ((S)->rowcnt[r_] + 31) / 32; ds_stat_iters32 += t_; }
#define DS_STAT_VISIT ++ds_stat_visits;
#else
#define DS_STAT_PASS(blk, S)
#define DS_STAT_VISIT
#endif
struct DsBlock {
  int x0, y0, x1, y1;  // inclusive cell rectangle
  int nrows;           // rows of the row table (1 when the rectangle spans the full grid width)
  int ncache;          // candidates in the scratch cache, -1 = too many (iterate the rows)
};
// Scratch of one lane group (shared memory on the device).
struct DsScratch {
  int rowbeg[DS_MAXROWS], rowcnt[DS_MAXROWS];
  int ccw[DS_MAXD], cw[DS_MAXD];
  int cid[DS_CACHE];
  DsPt cxy[DS_CACHE];
};
// for (every candidate of the block: id ID, coordinates PT) BODY -- lanes stride (SEQ = 0) or every
// lane visits all (SEQ = 1).  `continue` inside BODY skips to the next candidate.
#define DS_FOR_CAND(W, in, S, blk, SEQ, ID, PT, ...)  \
  DS_STAT_PASS(blk, S)                                                                                                                        \
  if ((blk).ncache >= 0) {                                                                        \
    for (int k_ = (SEQ) ? 0 : W::lane(); k_ < (blk).ncache; k_ += (SEQ) ? 1 : W::LANES) {         \
      const int ID = (S)->cid[k_];                                                                \
      const DsPt PT = (S)->cxy[k_];                                                               \
      DS_STAT_VISIT                                                                               \
      __VA_ARGS__                                                                                 \
    }                                                                                             \
  } else {                                                                                        \
    for (int r_ = 0; r_ < (blk).nrows; ++r_) {                                                    \
      const int beg_ = (S)->rowbeg[r_], end_ = beg_ + (S)->rowcnt[r_];                            \
      for (int k_ = beg_ + ((SEQ) ? 0 : W::lane()); k_ < end_; k_ += (SEQ) ? 1 : W::LANES) {      \
        const int ID = (in).sid[k_];                                                              \
        const DsPt PT = (in).sxy[k_];                                                             \
        DS_STAT_VISIT                                                                             \
        __VA_ARGS__                                                                               \
      }                                                                                           \
    }                                                                                             \
  }

DS_FN int ds_cellx(const DsIn& in, int64_t x) {
  const int64_t c = x >> in.shift;
  return c < 0 ? 0 : (c >= in.gx ? in.gx - 1 : (int)c);
}
DS_FN int ds_celly(const DsIn& in, int64_t y) {
  const int64_t c = y >> in.shift;
  return c < 0 ? 0 : (c >= in.gy ? in.gy - 1 : (int)c);
}

// What a block that is too large for the cache is scanned FOR: the sweep side of the chord p->cur
// (A (x - px) < B (y - py)) and, once a best candidate exists, the disc through (p, cur, best).  Such
// blocks belong to hull vertices and slivers, whose regions are thin diagonal strips or caps inside a
// large bounding rectangle: every cell row is narrowed to the part the region can reach.  Only
// uncached blocks are narrowed -- they are dropped at the next sweep step, so no later step inherits
// a block that was not scanned in full.
struct DsClip {
  int mode;            // 0: none, 1: half-plane, 2: half-plane and disc
  float px, py, A, B;  // half-plane
  float qx, qy, rr;    // disc centre (absolute lattice coordinates), radius incl. margin
};
DS_FN void ds_clip_halfplane(DsClip* c, DsPt p, DsPt cur, int dir) {
  c->mode = 1;
  c->px = (float)p.x; c->py = (float)p.y;
  c->A = (float)(cur.y - p.y) * (float)dir;
  c->B = (float)(cur.x - p.x) * (float)dir;
  c->qx = c->qy = c->rr = 0.0f;
}
// Row table of a block: contiguous ranges of the cell-sorted arrays (cells are sorted row-major);
// small blocks are copied into the scratch cache.
template <class W>
DS_FN void ds_block_rows(const DsIn& in, DsBlock& b, DsScratch* S, const DsClip* clip = nullptr) {
  W::sync();
  if (b.x0 == 0 && b.x1 == in.gx - 1) {
    b.nrows = 1;
    if (W::lane() == 0) {
      S->rowbeg[0] = in.cell_start[b.y0 * in.gx];
      S->rowcnt[0] = in.cell_start[(b.y1 + 1) * in.gx] - S->rowbeg[0];
    }
  } else {
    b.nrows = b.y1 - b.y0 + 1;
    for (int r = W::lane(); r < b.nrows; r += W::LANES) {
      const int beg = in.cell_start[(b.y0 + r) * in.gx + b.x0];
      S->rowbeg[r] = beg;
      S->rowcnt[r] = in.cell_start[(b.y0 + r) * in.gx + b.x1 + 1] - beg;
    }
  }
  W::sync();
  int total = 0;
  for (int r = 0; r < b.nrows; ++r) total += S->rowcnt[r];
  b.ncache = -1;
  if (total <= DS_CACHE) {
    int base = 0;
    for (int r = 0; r < b.nrows; ++r) {
      const int beg = S->rowbeg[r], cnt = S->rowcnt[r];
      for (int k = W::lane(); k < cnt; k += W::LANES) {
        S->cid[base + k] = in.sid[beg + k];
        S->cxy[base + k] = in.sxy[beg + k];
      }
      base += cnt;
    }
    b.ncache = total;
    W::sync();
  } else if (clip && clip->mode) {
    // every cell row narrowed to what the region can reach in it (conservative: margins cover the fp32 roundings)
    W::sync();
    b.nrows = b.y1 - b.y0 + 1;
    const float lim = 4194304.0f;  // 2^22: beyond every lattice coordinate
    for (int r = W::lane(); r < b.nrows; r += W::LANES) {
      const int Y = b.y0 + r;
      // lattice extent of the row's cells; the first / last row hold everything below / above
      const float ya = Y == 0 ? (float)in.by0 : (float)((int64_t)Y << in.shift);
      const float yb = Y == in.gy - 1 ? (float)in.by1 : (float)((((int64_t)Y + 1) << in.shift) - 1);
      float xlo = -lim, xhi = lim;
      bool empty = false;
      if (clip->A != 0.0f) {
        const float t0 = clip->B * (ya - clip->py) / clip->A, t1 = clip->B * (yb - clip->py) / clip->A;
        if (clip->A > 0.0f) {
          const float t = fmaxf(t0, t1);
          xhi = fminf(xhi, clip->px + t + 2.0f + 1e-6f * fabsf(t));
        } else {
          const float t = fminf(t0, t1);
          xlo = fmaxf(xlo, clip->px + t - 2.0f - 1e-6f * fabsf(t));
        }
      }
      if (clip->mode == 2) {
        const float dy = clip->qy < ya ? ya - clip->qy : (clip->qy > yb ? clip->qy - yb : 0.0f);
        if (dy > clip->rr + 2.0f) {
          empty = true;
        } else {
          const float w = sqrtf(fmaxf(clip->rr * clip->rr - dy * dy, 0.0f) + 1e-6f * clip->rr * clip->rr) + 2.0f + 1e-6f * fabsf(clip->qx);
          xlo = fmaxf(xlo, clip->qx - w);
          xhi = fminf(xhi, clip->qx + w);
        }
      }
      int beg = 0, cnt = 0;
      if (!empty && xlo <= xhi) {
        const int ca = ds_cellx(in, (int64_t)floorf(fmaxf(xlo, -lim))), cb = ds_cellx(in, (int64_t)ceilf(fminf(xhi, lim)));
        const int xa = ca > b.x0 ? ca : b.x0, xb = cb < b.x1 ? cb : b.x1;
        if (xa <= xb) {
          beg = in.cell_start[Y * in.gx + xa];
          cnt = in.cell_start[Y * in.gx + xb + 1] - beg;
        }
      }
      S->rowbeg[r] = beg;
      S->rowcnt[r] = cnt;
    }
    W::sync();
  }
}
DS_FN bool ds_block_cover(const DsIn& in, DsBlock& b, int64_t rx0, int64_t ry0, int64_t rx1, int64_t ry1) {
  const int nx0 = ds_cellx(in, rx0), nx1 = ds_cellx(in, rx1), ny0 = ds_celly(in, ry0), ny1 = ds_celly(in, ry1);
  const int sx = b.x1 - b.x0 + 1, sy = b.y1 - b.y0 + 1;
  bool ch = false;
  if (nx0 < b.x0) { b.x0 = nx0 > b.x0 - sx ? nx0 : b.x0 - sx; ch = true; }
  if (ny0 < b.y0) { b.y0 = ny0 > b.y0 - sy ? ny0 : b.y0 - sy; ch = true; }
  if (nx1 > b.x1) { b.x1 = nx1 < b.x1 + sx ? nx1 : b.x1 + sx; ch = true; }
  if (ny1 > b.y1) { b.y1 = ny1 < b.y1 + sy ? ny1 : b.y1 + sy; ch = true; }
  if (b.x0 < 0) b.x0 = 0;
  if (b.y0 < 0) b.y0 = 0;
  if (b.x1 > in.gx - 1) b.x1 = in.gx - 1;
  if (b.y1 > in.gy - 1) b.y1 = in.gy - 1;
  return ch;
}
DS_FN bool ds_block_all(const DsIn& in, const DsBlock& b) {
  return b.x0 == 0 && b.y0 == 0 && b.x1 == in.gx - 1 && b.y1 == in.gy - 1;
}

// Bounding box of the part of the closed disc through (p, cur, best) that lies on the sweep side of
// the chord p->cur, clipped to the bounding box of all points: the circular segment spanned by the
// chord's endpoints and those of the circle's four axis-extreme points that are on the sweep side.
// (Only sweep-side points can contradict `best`; for a sliver near the hull the segment is a thin
// cap while the whole disc would cover half the image.)  Conservative: rounded outward.
struct DsBox { int64_t r[4]; int ok; };
DS_FN void ds_circle_region_f64_impl(const DsIn& in, DsPt p, DsPt cur, DsPt best, int dir, int64_t* r) {
  // (p, a, b) counter-clockwise
  const DsPt a = dir > 0 ? cur : best, b = dir > 0 ? best : cur;
  const double ax = (double)(a.x - p.x), ay = (double)(a.y - p.y), bx = (double)(b.x - p.x), by = (double)(b.y - p.y);
  const double d = 2.0 * (ax * by - ay * bx);  // > 0, exact
  const double a2 = ax * ax + ay * ay, b2 = bx * bx + by * by;
  const double ux = (by * a2 - ay * b2) / d, uy = (ax * b2 - bx * a2) / d;  // centre - p
  const double rad = sqrt(ux * ux + uy * uy);
  const double m = 2.0 + rad * 1e-9 + (fabs(ux) + fabs(uy)) * 1e-9;
  const double cx = (double)(cur.x - p.x), cy = (double)(cur.y - p.y);   // chord
  const double tol = 1e-6 * (fabs(cx) + fabs(cy)) * (rad + 1.0) + 4.0;
  double x0 = fmin(0.0, cx), x1 = fmax(0.0, cx), y0 = fmin(0.0, cy), y1 = fmax(0.0, cy);
  const double ex[4] = {ux - rad, ux + rad, ux, ux}, ey[4] = {uy, uy, uy - rad, uy + rad};
  for (int k = 0; k < 4; ++k) {
    const double side = (cx * ey[k] - cy * ex[k]) * (double)dir;
    if (side > -tol) {
      x0 = fmin(x0, ex[k]); x1 = fmax(x1, ex[k]);
      y0 = fmin(y0, ey[k]); y1 = fmax(y1, ey[k]);
    }
  }
  x0 = fmax((double)p.x + x0 - m, (double)in.bx0); y0 = fmax((double)p.y + y0 - m, (double)in.by0);
  x1 = fmin((double)p.x + x1 + m, (double)in.bx1); y1 = fmin((double)p.y + y1 + m, (double)in.by1);
  r[0] = (int64_t)floor(x0); r[1] = (int64_t)floor(y0); r[2] = (int64_t)ceil(x1); r[3] = (int64_t)ceil(y1);
}

// Out of line (returned by value): rare, and inlined its double-precision code was hoisted in front of
// every sweep step by loop-invariant code motion.
DS_COLD DsBox ds_circle_region_f64(const DsIn& in, DsPt p, DsPt cur, DsPt best, int dir) {
  DsBox b;
  ds_circle_region_f64_impl(in, p, cur, best, dir, b.r);
  b.ok = 1;
  return b;
}
// The same region at fp32 cost.  Only the circumcentre's numerators and denominator need the exact
// (double) products -- they cancel for slivers; everything after the division only has to be
// CONSERVATIVE, so it runs in fp32 with margins that cover every rounding (relative 2^-22 on the
// centre and radius, against 2e-6 here).  Circles too large for that (radius > 1e12 lattice units)
// take the double path.  The region never decides a predicate: a looser box only means more
// candidates are looked at, so host and device may round differently here without consequence.
// Returns true when the WHOLE disc already lies inside the block's cells (the common case for an
// interior vertex: nothing to refine, nothing to grow); otherwise r receives the segment's box.
DS_FN bool ds_circle_region(const DsIn& in, DsPt p, DsPt cur, DsPt best, int dir, const DsBlock& blk, int64_t* r, DsClip* clip) {
  ds_clip_halfplane(clip, p, cur, dir);
  const DsPt a = dir > 0 ? cur : best, b = dir > 0 ? best : cur;  // (p, a, b) counter-clockwise
  const double ax = (double)(a.x - p.x), ay = (double)(a.y - p.y), bx = (double)(b.x - p.x), by = (double)(b.y - p.y);
  const double d = 2.0 * (ax * by - ay * bx);  // > 0, exact
  const double a2 = ax * ax + ay * ay, b2 = bx * bx + by * by;
  const float fd = (float)d;
  const float ux = (float)(by * a2 - ay * b2) / fd, uy = (float)(ax * b2 - bx * a2) / fd;  // centre - p
  const float rad = sqrtf(ux * ux + uy * uy);
  if (!(rad < 1e12f)) {
    const DsBox b = ds_circle_region_f64(in, p, cur, best, dir);
    r[0] = b.r[0]; r[1] = b.r[1]; r[2] = b.r[2]; r[3] = b.r[3];
    return false;
  }
  const float ext = rad + fabsf(ux) + fabsf(uy);
  const float m = 2.0f + ext * 2e-6f;
  {
    // the disc's own box against the block's extent in lattice units (a border cell extends to infinity:
    // ds_cellx / ds_celly clamp)
    const float big = 3.0e38f, dm = rad + m;
    // (cell index << shift stays below 2^22: 32-bit arithmetic, exact in fp32)
    const float lx = blk.x0 == 0 ? -big : (float)(blk.x0 << in.shift), hx = blk.x1 == in.gx - 1 ? big : (float)(((blk.x1 + 1) << in.shift) - 1);
    const float ly = blk.y0 == 0 ? -big : (float)(blk.y0 << in.shift), hy = blk.y1 == in.gy - 1 ? big : (float)(((blk.y1 + 1) << in.shift) - 1);
    const float qx = (float)p.x + ux, qy = (float)p.y + uy;
    if (qx - dm >= lx && qx + dm <= hx && qy - dm >= ly && qy + dm <= hy) return true;
  }
  clip->mode = 2;
  clip->qx = (float)p.x + ux; clip->qy = (float)p.y + uy; clip->rr = rad + m;
  const float cx = (float)(cur.x - p.x), cy = (float)(cur.y - p.y);  // chord, exact
  const float tol = 4e-6f * (fabsf(cx) + fabsf(cy)) * ext + 4.0f;
  const float fdir = (float)dir;
  float x0 = cx < 0.0f ? cx : 0.0f, x1 = cx > 0.0f ? cx : 0.0f, y0 = cy < 0.0f ? cy : 0.0f, y1 = cy > 0.0f ? cy : 0.0f;
  // the circle's axis-extreme points that lie on the sweep side of the chord
  const bool w = (cx * uy - cy * (ux - rad)) * fdir > -tol, e = (cx * uy - cy * (ux + rad)) * fdir > -tol;
  const bool n = (cx * (uy - rad) - cy * ux) * fdir > -tol, t = (cx * (uy + rad) - cy * ux) * fdir > -tol;
  if (w) x0 = fminf(x0, ux - rad);
  if (e) x1 = fmaxf(x1, ux + rad);
  if (n) y0 = fminf(y0, uy - rad);
  if (t) y1 = fmaxf(y1, uy + rad);
  if (w || e) { y0 = fminf(y0, uy); y1 = fmaxf(y1, uy); }
  if (n || t) { x0 = fminf(x0, ux); x1 = fmaxf(x1, ux); }
  x0 = fmaxf((float)p.x + x0 - m, (float)in.bx0); y0 = fmaxf((float)p.y + y0 - m, (float)in.by0);
  x1 = fminf((float)p.x + x1 + m, (float)in.bx1); y1 = fminf((float)p.y + y1 + m, (float)in.by1);
  r[0] = (int64_t)floorf(x0); r[1] = (int64_t)floorf(y0); r[2] = (int64_t)ceilf(x1); r[3] = (int64_t)ceilf(y1);
  return false;
}

// Bounding box of (half-plane on the sweep side of p->cur) clipped to the bounding box of all
// points; false when the intersection is empty.
DS_FN bool ds_halfplane_region_impl(const DsIn& in, DsPt p, DsPt cur, int dir, int64_t* r) {
  const DsPt c[4] = {{in.bx0, in.by0}, {in.bx1, in.by0}, {in.bx1, in.by1}, {in.bx0, in.by1}};
  int64_t o[4];
  for (int k = 0; k < 4; ++k) o[k] = ds_orient(p, cur, c[k]) * dir;
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  bool any = false;
  for (int k = 0; k < 4; ++k) {
    if (o[k] > 0) {
      any = true;
      x0 = fmin(x0, (double)c[k].x); x1 = fmax(x1, (double)c[k].x);
      y0 = fmin(y0, (double)c[k].y); y1 = fmax(y1, (double)c[k].y);
    }
    const int k2 = (k + 1) & 3;
    if ((o[k] > 0) != (o[k2] > 0)) {
      const double t = (double)o[k] / ((double)o[k] - (double)o[k2]);
      const double ix = (double)c[k].x + t * (double)(c[k2].x - c[k].x);
      const double iy = (double)c[k].y + t * (double)(c[k2].y - c[k].y);
      any = true;
      x0 = fmin(x0, ix - 2.0); x1 = fmax(x1, ix + 2.0);
      y0 = fmin(y0, iy - 2.0); y1 = fmax(y1, iy + 2.0);
    }
  }
  if (!any) return false;
  x0 = fmax(x0, (double)in.bx0); y0 = fmax(y0, (double)in.by0);
  x1 = fmin(x1, (double)in.bx1); y1 = fmin(y1, (double)in.by1);
  r[0] = (int64_t)floor(x0); r[1] = (int64_t)floor(y0); r[2] = (int64_t)ceil(x1); r[3] = (int64_t)ceil(y1);
  return true;
}

// Out of line for the same reason: it depends only on (p, cur, dir), so inlined it was computed (four
// double divisions) at the top of EVERY sweep step -- 22 % of the star kernel's instructions -- although
// only a step that finds no candidate (a hull end) needs it.
DS_COLD DsBox ds_halfplane_region(const DsIn& in, DsPt p, DsPt cur, int dir) {
  DsBox b;
  b.ok = ds_halfplane_region_impl(in, p, cur, dir, b.r) ? 1 : 0;
  return b;
}

// ------------------------------------------------------------------------------------ sweep step
// Next neighbour of p after cur in direction dir; -1 when there is none (hull end).  All lanes
// return the same value; *nxy receives its coordinates.  *err is set on an invariant violation.
template <class W>
DS_FN int ds_next(const DsIn& in, int p, DsPt pp, int curid, DsPt cur, int dir, DsBlock& blk, DsScratch* S,
                  DsPt* nxy, int* err) {
  if (blk.ncache < 0) {
    // the previous step needed a block too large for the cache (a sliver near the hull): this step
    // starts from the 3x3 block again instead of dragging the large one along
    const int cxp = ds_cellx(in, pp.x), cyp = ds_celly(in, pp.y);
    blk.x0 = cxp > 0 ? cxp - 1 : 0; blk.x1 = cxp < in.gx - 1 ? cxp + 1 : in.gx - 1;
    blk.y0 = cyp > 0 ? cyp - 1 : 0; blk.y1 = cyp < in.gy - 1 ? cyp + 1 : in.gy - 1;
    ds_block_rows<W>(in, blk, S);
  }
  for (int guard = 0; guard < 4 * DS_MAXROWS; ++guard) {
    // One pass finds the best candidate AND whether other candidates lie on its circle: on the sweep
    // side the discs through the chord (p, cur) are nested, so a candidate on the FINAL circle is
    // either compared with a best that already has that circle (in-circle == 0: tie recorded) or
    // becomes best itself and meets the others later.  A strictly better candidate shrinks the
    // circle and clears the flag.
    int bid = -1;
    bool tie = false;
    DsPt bxy = {0, 0};
    DS_FOR_CAND(W, in, S, blk, 0, id, c, {
      if (id < 0 || id == p || id == curid) continue;
      if (!ds_side(pp, cur, c, dir)) continue;
      if (bid < 0) {
        bid = id;
        bxy = c;
        continue;
      }
      const int r = ds_inside(pp, cur, bxy, c, dir);
      if (r > 0) {
        bid = id;
        bxy = c;
        tie = false;
      } else if (r == 0) {
        tie = true;
      }
    })
    if (W::LANES > 1) {
      // Across lanes: the lane bests' inscribed angles over the chord (p, cur) PROPOSE a winner (fp32
      // cotangent, one redux.sync), one exact in-circle per lane DISPOSES: the proposal stands iff no
      // lane's best lies strictly inside its circle (the lanes' own candidates are then outside too,
      // the sweep-side discs being nested); bests ON the circle are the ties.  A refuted proposal
      // (fp32 misordering two nearly co-circular candidates) falls back to the pairwise tournament.
      unsigned ukey = 0xffffffffu;
      if (bid >= 0) {
        const float ux = (float)(pp.x - bxy.x), uy = (float)(pp.y - bxy.y), vx = (float)(cur.x - bxy.x), vy = (float)(cur.y - bxy.y);
        const float key = (ux * vx + uy * vy) / fabsf(ux * vy - uy * vx);
        const unsigned kb = DS_F2U(key);
        ukey = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);  // order-preserving
        if (ukey == 0xffffffffu) ukey = 0xfffffffeu;
      }
      const unsigned umin = W::min_u32(ukey);
      if (umin != 0xffffffffu) {
        const int src = W::first(ukey == umin);
        const int wid = W::bcast(bid, src);
        DsPt wxy;
        wxy.x = W::bcast(bxy.x, src);
        wxy.y = W::bcast(bxy.y, src);
        int r = -1;
        if (bid >= 0 && bid != wid) r = ds_inside(pp, cur, wxy, bxy, dir);
        if (!W::any(r > 0)) {
          tie = W::any(bid >= 0 && (bid == wid ? tie : r == 0));
          bid = wid;
          bxy = wxy;
        } else {
          for (int o = W::LANES >> 1; o > 0; o >>= 1) {
            const int oid = W::shfl_xor(bid, o);
            const int otie = W::shfl_xor(tie ? 1 : 0, o);
            DsPt oxy;
            oxy.x = W::shfl_xor(bxy.x, o);
            oxy.y = W::shfl_xor(bxy.y, o);
            if (oid < 0 || oid == bid) continue;
            if (bid < 0) {
              bid = oid; bxy = oxy; tie = otie != 0;
              continue;
            }
            const int r2 = ds_inside(pp, cur, bxy, oxy, dir);
            if (r2 > 0) { bid = oid; bxy = oxy; tie = otie != 0; }
            else if (r2 == 0) tie = true;
          }
          bid = W::bcast(bid, 0);  // tied lanes may disagree on the representative: lane 0 decides
          bxy.x = W::bcast(bxy.x, 0);
          bxy.y = W::bcast(bxy.y, 0);
          tie = W::bcast(tie ? 1 : 0, 0) != 0;
        }
      }
    }
    int64_t reg[4];
    if (bid < 0) {
      if (ds_block_all(in, blk)) return -1;
      const DsBox hb = ds_halfplane_region(in, pp, cur, dir);
      if (!hb.ok) return -1;
      reg[0] = hb.r[0]; reg[1] = hb.r[1]; reg[2] = hb.r[2]; reg[3] = hb.r[3];
      if (!ds_block_cover(in, blk, reg[0], reg[1], reg[2], reg[3])) return -1;
      DsClip clip;
      ds_clip_halfplane(&clip, pp, cur, dir);
      ds_block_rows<W>(in, blk, S, &clip);
      continue;
    }
    if (!ds_block_all(in, blk)) {
      DsClip clip;
      if (!ds_circle_region(in, pp, cur, bxy, dir, blk, reg, &clip) && ds_block_cover(in, blk, reg[0], reg[1], reg[2], reg[3])) {
        ds_block_rows<W>(in, blk, S, &clip);
        continue;
      }
    }
    // the circle (p, cur, best) is empty; other points ON it make a co-circular polygon
    if (tie) {
      // canonical fan from the smallest index (all lanes run the same sequential pass)
      int minid = p < curid ? p : curid;
      int cl = bid, fa = bid, mt = -1;
      DsPt clxy = bxy, faxy = bxy, mtxy = bxy;
      if (bid < minid) { minid = bid; mt = bid; }
      DS_FOR_CAND(W, in, S, blk, 1, id, c, {
        if (id < 0 || id == p || id == curid || id == bid) continue;
        if (!ds_side(pp, cur, c, dir) || ds_inside(pp, cur, bxy, c, dir) != 0) continue;
        if (id < minid) { minid = id; mt = id; mtxy = c; }
        if (ds_orient(pp, c, clxy) * dir > 0) { cl = id; clxy = c; }   // c comes before the closest so far
        if (ds_orient(pp, faxy, c) * dir > 0) { fa = id; faxy = c; }   // c comes after the farthest so far
      })
      if (minid == p) { bid = cl; bxy = clxy; }
      else if (minid == curid) { bid = fa; bxy = faxy; }
      else if (mt >= 0) { bid = mt; bxy = mtxy; }
      else { *err = DS_E_TIE; return -1; }
    }
    *nxy = bxy;
    return bid;
  }
  *err = DS_E_LOOP;
  return -1;
}

// ------------------------------------------------------------------------------------ star of p
// S: the lane group's scratch.  Result (lane 0 writes): star[0..deg) counter-clockwise; *closed = 1
// for an interior vertex.
template <class W>
DS_FN int ds_star(const DsIn& in, int p, DsScratch* S, int* star, int* deg_out, int* closed_out) {
  const DsPt pp = in.vxy[p];
  const int cxp = ds_cellx(in, pp.x), cyp = ds_celly(in, pp.y);
  DsBlock blk;
  blk.x0 = cxp > 0 ? cxp - 1 : 0; blk.x1 = cxp < in.gx - 1 ? cxp + 1 : in.gx - 1;
  blk.y0 = cyp > 0 ? cyp - 1 : 0; blk.y1 = cyp < in.gy - 1 ? cyp + 1 : in.gy - 1;
  ds_block_rows<W>(in, blk, S);
  *deg_out = 0;
  *closed_out = 0;
  // ---- nearest neighbour
  int n0 = -1;
  DsPt n0xy = {0, 0};
  for (int guard = 0; guard < 4 * DS_MAXROWS; ++guard) {
    long long bd = 0x7fffffffffffffffll;
    int bid = -1;
    DsPt bxy = {0, 0};
    DS_FOR_CAND(W, in, S, blk, 0, id, c, {
      if (id < 0 || id == p) continue;
      const long long dx = c.x - pp.x, dy = c.y - pp.y, d2 = dx * dx + dy * dy;
      if (d2 < bd || (d2 == bd && id < bid)) { bd = d2; bid = id; bxy = c; }
    })
    if (W::LANES > 1) {
      // min over lanes of (d2, id): three redux.sync (high word, low word among those, id among those)
      const unsigned hi = (unsigned)((unsigned long long)bd >> 32), lo = (unsigned)bd;
      const unsigned mh = W::min_u32(hi);
      const unsigned ml = W::min_u32(hi == mh ? lo : 0xffffffffu);
      const bool mine = bid >= 0 && hi == mh && lo == ml;
      const unsigned mid = W::min_u32(mine ? (unsigned)bid : 0xffffffffu);
      if (mid == 0xffffffffu) {
        bid = -1;
      } else {
        const int src = W::first(mine && (unsigned)bid == mid);
        bid = (int)mid;
        bxy.x = W::bcast(bxy.x, src);
        bxy.y = W::bcast(bxy.y, src);
        bd = ((long long)mh << 32) | (long long)ml;
      }
    }
    if (bid < 0) {
      if (ds_block_all(in, blk)) return DS_OK;  // the only point
      const int hx = blk.x1 - blk.x0 + 1, hy = blk.y1 - blk.y0 + 1;
      blk.x0 = blk.x0 - hx < 0 ? 0 : blk.x0 - hx; blk.x1 = blk.x1 + hx > in.gx - 1 ? in.gx - 1 : blk.x1 + hx;
      blk.y0 = blk.y0 - hy < 0 ? 0 : blk.y0 - hy; blk.y1 = blk.y1 + hy > in.gy - 1 ? in.gy - 1 : blk.y1 + hy;
      ds_block_rows<W>(in, blk, S);
      continue;
    }
    const int64_t rr = (int64_t)ceil(sqrt((double)bd)) + 2;
    int64_t x0 = pp.x - rr, x1 = pp.x + rr, y0 = pp.y - rr, y1 = pp.y + rr;
    if (x0 < in.bx0) x0 = in.bx0;
    if (y0 < in.by0) y0 = in.by0;
    if (x1 > in.bx1) x1 = in.bx1;
    if (y1 > in.by1) y1 = in.by1;
    if (!ds_block_all(in, blk) && ds_block_cover(in, blk, x0, y0, x1, y1)) {
      ds_block_rows<W>(in, blk, S);
      continue;
    }
    n0 = bid;
    n0xy = bxy;
    break;
  }
  if (n0 < 0) return DS_E_LOOP;
  // ---- counter-clockwise sweep from n0
  int err = DS_OK, nccw = 1, ncw = 0, closed = 0;
  if (W::lane() == 0) S->ccw[0] = n0;
  int curid = n0;
  DsPt cur = n0xy;
  // one loop for both sweeps (a single inlined copy of ds_next): counter-clockwise until the star
  // closes at n0 or the hull ends it, then -- hull vertex -- clockwise from n0 again
  int dir = +1;
  for (;;) {
    DsPt nxy;
    const int nx = ds_next<W>(in, p, pp, curid, cur, dir, blk, S, &nxy, &err);
    if (err) return err;
    if (dir > 0) {
      if (nx < 0) {
        dir = -1;
        curid = n0;
        cur = n0xy;
        continue;
      }
      if (nx == n0) { closed = 1; break; }
      if (nccw >= DS_MAXD) return DS_E_DEGREE;
      if (W::lane() == 0) S->ccw[nccw] = nx;
      ++nccw;
    } else {
      if (nx < 0) break;
      if (nccw + ncw >= DS_MAXD) return DS_E_DEGREE;
      if (W::lane() == 0) S->cw[ncw] = nx;
      ++ncw;
    }
    curid = nx;
    cur = nxy;
  }
  W::sync();
  if (W::lane() == 0) {
    for (int k = 0; k < ncw; ++k) star[k] = S->cw[ncw - 1 - k];
    for (int k = 0; k < nccw; ++k) star[ncw + k] = S->ccw[k];
  }
  W::sync();
  *deg_out = nccw + ncw;
  *closed_out = closed;
  return DS_OK;
}

// ------------------------------------------------------------------------------------ stars -> mesh
// Canonical mesh from the stars (shared by the device kernels and the host simulation):
//   edges (i, j), i < j, sorted by (i, j): vertex v contributes its neighbours larger than v;
//   triangles (v, a, b) counter-clockwise with v the smallest index, sorted by (v, a): vertex v
//   contributes every consecutive star pair (a, b) with a > v and b > v.
DS_FN void ds_counts(int v, const int* star, int deg, int closed, int* od, int* tc) {
  int o = 0, t = 0;
  for (int k = 0; k < deg; ++k) o += star[k] > v ? 1 : 0;
  const int pairs = closed ? deg : deg - 1;
  for (int k = 0; k < pairs; ++k) {
    const int a = star[k], b = star[k + 1 == deg ? 0 : k + 1];
    t += (a > v && b > v) ? 1 : 0;
  }
  *od = o;
  *tc = t;
}
// outs[0..od): ascending; tris[0..3*tc): (v, a, b) ascending a.  Returns od; *tc_out = tc.
DS_FN int ds_emit(int v, const int* star, int deg, int closed, int* outs, int* tris, int* tc_out) {
  int o = 0, t = 0;
  for (int k = 0; k < deg; ++k) {
    const int w = star[k];
    if (w <= v) continue;
    int pos = o++;
    while (pos > 0 && outs[pos - 1] > w) { outs[pos] = outs[pos - 1]; --pos; }
    outs[pos] = w;
  }
  const int pairs = closed ? deg : deg - 1;
  for (int k = 0; k < pairs; ++k) {
    const int a = star[k], b = star[k + 1 == deg ? 0 : k + 1];
    if (!(a > v && b > v)) continue;
    int pos = t++;
    while (pos > 0 && tris[3 * (pos - 1) + 1] > a) {
      tris[3 * pos + 1] = tris[3 * (pos - 1) + 1];
      tris[3 * pos + 2] = tris[3 * (pos - 1) + 2];
      --pos;
    }
    tris[3 * pos + 1] = a;
    tris[3 * pos + 2] = b;
  }
  for (int k = 0; k < t; ++k) tris[3 * k] = v;
  *tc_out = t;
  return o;
}

// Lattice coordinate of a pixel position (same rounding as delaunay.h: llroundf(v * 64)).
DS_FN int32_t ds_lattice(float v, int* bad) {
  const long long l = llroundf(v * 64.0f);
  if (l <= -(long long)DS_COORD_LIM || l >= (long long)DS_COORD_LIM) { *bad = 1; return 0; }
  return (int32_t)l;
}
