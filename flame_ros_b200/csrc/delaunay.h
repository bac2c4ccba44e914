// delaunay.h -- incremental (Bowyer-Watson) Delaunay triangulation of the graph vertices.
//
// Stands in for the `triangulate` stage of flame::Flame::update (timing key
// /root/reference/src/utils.cc:154; the external flame core wraps Shewchuk's Triangle, which is not
// in this image).  Host C++: the point set is a few thousand vertices and the work is
// pointer-chasing, so it runs on the CPU overlapped with GPU work; the solver consumes its edge list.
//
// Robustness: vertex positions are snapped to a 1/64 px lattice and all predicates (orientation,
// in-circle) are evaluated exactly in 64/128-bit integers, so co-circular grid detections and
// collinear points cannot corrupt the structure.  The convex hull is closed with "ghost" triangles
// (a hull edge + a vertex at infinity), which makes point location and cavity carving uniform.
// Output: triangles (counter-clockwise in the y-down pixel frame's mathematical sense of the
// stored coordinates), canonical edges (i<j, sorted by (i,j)).  Duplicate points are merged onto
// the first occurrence and reported through `dup_of`.
#pragma once

#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace fbdel {

typedef __int128 i128;

struct Triangulator {
  static const int GHOST = -1;
  std::vector<int32_t> px, py;  // lattice coordinates (|v| < 2^30: differences and orientation products fit 64 bits)
  // one 32-byte record per triangle: tt[8t + k] = vertex k (ccw; ghost triangles hold GHOST at slot 2),
  // tt[8t + 4 + k] = triangle across the edge (v[k], v[k+1]); vertices and neighbours of a triangle
  // share half a cache line
  std::vector<int> tt;
  std::vector<char> dead;
  std::vector<int> free_list;
  std::vector<int> dup_of;      // -1, or index of the earlier identical point
  int last = 0;

  static int64_t orient(int64_t ax, int64_t ay, int64_t bx, int64_t by, int64_t cx, int64_t cy) {
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);  // > 0: c is to the left of a->b (ccw)
  }
  int64_t orientv(int a, int b, int c) const { return orient(px[a], py[a], px[b], py[b], px[c], py[c]); }

  // > 0 iff p lies strictly inside the circumcircle of ccw triangle (a,b,c).  Lattice coordinates
  // are < 2^24, so differences and the 2x2 minors are exact in double; only the final three-term
  // sum rounds.  A static error bound decides the sign; inside the bound the determinant is
  // re-evaluated exactly in 128-bit integers (co-circular grid detections land there).
  int incircle_sign(int a, int b, int c, int p) const {
    const int64_t adx = px[a] - px[p], ady = py[a] - py[p];
    const int64_t bdx = px[b] - px[p], bdy = py[b] - py[p];
    const int64_t cdx = px[c] - px[p], cdy = py[c] - py[p];
    const int64_t al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    const int64_t ma = bdx * cdy - bdy * cdx, mb = cdx * ady - cdy * adx, mc = adx * bdy - ady * bdx;
    if (exact_double) {
      const double ta_ = (double)al * (double)ma, tb_ = (double)bl * (double)mb, tc_ = (double)cl * (double)mc;
      const double det = ta_ + tb_ + tc_;
      const double bound = 8.9e-16 * (fabs(ta_) + fabs(tb_) + fabs(tc_));  // 8 ulp of the magnitude sum
      if (det > bound) return 1;
      if (det < -bound) return -1;
    }
    const i128 det = (i128)al * (i128)ma + (i128)bl * (i128)mb + (i128)cl * (i128)mc;
    return det > 0 ? 1 : (det < 0 ? -1 : 0);
  }
  bool exact_double = false;  // set by run(): all |coordinates| < 2^24 => al, ma, .. exact in double

  // Does the (possibly ghost) triangle t conflict with point p (p inside its circumdisk)?
  bool conflicts(int t, int p) const {
    const int a = tt[8 * t], b = tt[8 * t + 1], c = tt[8 * t + 2];
    if (c != GHOST) return incircle_sign(a, b, c, p) > 0;
    // ghost (a,b,inf): the "disk" is the open half-plane to the left of a->b (outside the hull, which
    // lies to the right of the reversed hull edge) plus the open segment ab
    const int64_t o = orientv(a, b, p);
    if (o > 0) return true;
    if (o < 0) return false;
    const int64_t dx = px[b] - px[a], dy = py[b] - py[a];
    const int64_t t0 = (px[p] - px[a]) * dx + (py[p] - py[a]) * dy;
    return t0 > 0 && t0 < dx * dx + dy * dy;
  }

  int new_tri(int a, int b, int c) {
    int t;
    if (!free_list.empty()) {
      t = free_list.back();
      free_list.pop_back();
      dead[t] = 0;
    } else {
      t = (int)dead.size();
      dead.push_back(0);
      tt.resize(tt.size() + 8);
    }
    tt[8 * t] = a; tt[8 * t + 1] = b; tt[8 * t + 2] = c;
    tt[8 * t + 4] = tt[8 * t + 4 + 1] = tt[8 * t + 4 + 2] = -1;
    return t;
  }

  // Walk from `last` toward p; returns a triangle that conflicts with p.
  int locate(int p) const {
    int t = last;
    if (dead[t]) {
      for (t = 0; t < (int)dead.size(); ++t)
        if (!dead[t]) break;
    }
    for (int guard = 0; guard < 4 * (int)dead.size() + 64; ++guard) {
      if (tt[8 * t + 2] == GHOST) {
        if (conflicts(t, p)) return t;
        // slide along the hull: move to the neighbouring ghost whose edge faces p
        const int a = tt[8 * t], b = tt[8 * t + 1];
        const int64_t dx = px[b] - px[a], dy = py[b] - py[a];
        const int64_t t0 = (px[p] - px[a]) * dx + (py[p] - py[a]) * dy;
        if (orientv(a, b, p) < 0) {
          t = tt[8 * t + 4];  // p is on the hull's side: step into the real triangle
        } else {
          t = (t0 <= 0) ? tt[8 * t + 4 + 2] : tt[8 * t + 4 + 1];  // collinear, beyond a or beyond b
        }
        continue;
      }
      bool moved = false;
      for (int k = 0; k < 3; ++k) {
        const int a = tt[8 * t + k], b = tt[8 * t + (k + 1) % 3];
        if (orientv(a, b, p) < 0) {
          t = tt[8 * t + 4 + k];
          moved = true;
          break;
        }
      }
      if (!moved) return t;  // p inside or on the boundary of t: t's circumdisk contains p
    }
    // walking failed (should not happen): exhaustive search
    for (int u = 0; u < (int)dead.size(); ++u)
      if (!dead[u] && conflicts(u, p)) return u;
    return -1;
  }

  // scratch of insert(): plain arrays with explicit counts (push_back bookkeeping was ~25 % of an
  // insertion); `mark` holds the stamp of the insertion that last put a triangle into a cavity
  std::vector<int> cavity, stack_, bnd_a, bnd_b, bnd_out, bnd_new;
  std::vector<int> mark, vslot;
  int stamp = 0;

  void grow_scratch(size_t need) {
    if (cavity.size() < need) {
      const size_t cap = std::max<size_t>(2 * need, 64);
      cavity.resize(cap); stack_.resize(cap);
      bnd_a.resize(3 * cap); bnd_b.resize(3 * cap); bnd_out.resize(3 * cap); bnd_new.resize(3 * cap);
    }
  }

  bool insert(int p) {
    int t0 = locate(p);
    if (t0 < 0) return false;
    for (int k = 0; k < 3; ++k) {
      const int v = tt[8 * t0 + k];
      if (v != GHOST && px[v] == px[p] && py[v] == py[p]) return false;  // duplicate vertex
    }
    if (!conflicts(t0, p)) {
      // p sits exactly on a vertex/edge whose triangle does not strictly contain it in its disk:
      // look for any neighbour that conflicts (happens for points on shared edges / cocircular)
      bool found = false;
      for (int k = 0; k < 3 && !found; ++k) {
        const int n = tt[8 * t0 + 4 + k];
        if (n >= 0 && conflicts(n, p)) {
          t0 = n;
          found = true;
        }
      }
      if (!found) {
        for (int u = 0; u < (int)dead.size() && !found; ++u)
          if (!dead[u] && conflicts(u, p)) {
            t0 = u;
            found = true;
          }
      }
      if (!found) return false;  // duplicate of an existing vertex
    }
    if (mark.size() < dead.size() + 16) mark.resize(2 * dead.size() + 64, 0);
    ++stamp;
    grow_scratch(16);
    int ncav = 0, nstack = 0;
    stack_[nstack++] = t0;
    mark[t0] = stamp;
    while (nstack > 0) {
      const int t = stack_[--nstack];
      if ((size_t)ncav + 4 > cavity.size()) grow_scratch(cavity.size() + 4);
      cavity[ncav++] = t;
      for (int k = 0; k < 3; ++k) {
        const int n = tt[8 * t + 4 + k];
        if (n >= 0 && mark[n] != stamp && conflicts(n, p)) {
          mark[n] = stamp;
          stack_[nstack++] = n;
        }
      }
    }
    // boundary edges (a,b) of the cavity with the triangle outside
    int nb = 0;
    for (int c = 0; c < ncav; ++c) {
      const int t = cavity[c];
      for (int k = 0; k < 3; ++k) {
        const int n = tt[8 * t + 4 + k];
        if (n < 0 || mark[n] != stamp) {
          bnd_a[nb] = tt[8 * t + k];
          bnd_b[nb] = tt[8 * t + (k == 2 ? 0 : k + 1)];
          bnd_out[nb] = n;
          ++nb;
        }
      }
    }
    for (int c = 0; c < ncav; ++c) {
      dead[cavity[c]] = 1;
      free_list.push_back(cavity[c]);
    }
    // fan of new triangles (a, b, p); ghost boundary edges keep GHOST in slot 2
    for (int i = 0; i < nb; ++i) {
      const int a = bnd_a[i], b = bnd_b[i];
      int t, slot;
      if (a == GHOST) { t = new_tri(b, p, GHOST); slot = 2; }       // edge (inf, b): new ghost (b, p, inf), edge (inf,b) is slot 2
      else if (b == GHOST) { t = new_tri(p, a, GHOST); slot = 1; }  // edge (a, inf): new ghost (p, a, inf), edge (a,inf) is slot 1
      else { t = new_tri(a, b, p); slot = 0; }                      // (a,b,p): edge (a,b) is slot 0
      bnd_new[i] = t;
      // link across the boundary edge
      const int n = bnd_out[i];
      tt[8 * t + 4 + slot] = n;
      if (n >= 0) {
        const int* nv = &tt[8 * n];
        const int k = (nv[0] == b && nv[1] == a) ? 0 : ((nv[1] == b && nv[2] == a) ? 1 : 2);
        tt[8 * n + 4 + k] = t;
      }
    }
    if (mark.size() < dead.size() + 16) mark.resize(2 * dead.size() + 64, 0);
    // link the fan triangles to each other: the boundary edges form a cycle around p, so triangle i's
    // edge (b_i, p) is shared with the triangle whose boundary edge starts at b_i (its edge (p, a))
    if (vslot.size() < px.size() + 1) vslot.assign(px.size() + 1, -1);
    const int gidx = (int)px.size();  // slot of the ghost vertex
    for (int i = 0; i < nb; ++i) vslot[bnd_a[i] == GHOST ? gidx : bnd_a[i]] = i;
    for (int i = 0; i < nb; ++i) {
      const int a = bnd_a[i], b = bnd_b[i];
      const int j = vslot[b == GHOST ? gidx : b];
      const int bp = (a == GHOST) ? 0 : (b == GHOST ? 2 : 1);
      const int aj = bnd_a[j], bj = bnd_b[j];
      const int pa = (aj == GHOST) ? 1 : (bj == GHOST ? 0 : 2);
      tt[8 * bnd_new[i] + 4 + bp] = bnd_new[j];
      tt[8 * bnd_new[j] + 4 + pa] = bnd_new[i];
    }
    last = bnd_new[0];
    return true;
  }

  // Canonical form for co-circular point sets (integer-pixel detections form exact rectangles): the
  // points on one empty circle span a convex polygon that every Delaunay triangulation fills somehow;
  // flipping a tied diagonal (a,b) to (c,d) whenever min(c,d) < min(a,b) terminates (the sum of the
  // diagonals' smaller endpoints falls) in the fan from the polygon's smallest vertex index.  The
  // per-vertex stars of delaunay_star.h produce the same fan, so both triangulators agree exactly.
  void canonicalize() {
    std::vector<int> work;
    for (int t = 0; t < (int)dead.size(); ++t)
      if (!dead[t] && tt[8 * t + 2] != GHOST) work.push_back(t);
    while (!work.empty()) {
      const int t = work.back();
      work.pop_back();
      if (dead[t] || tt[8 * t + 2] == GHOST) continue;
      for (int k = 0; k < 3; ++k) {
        const int n = tt[8 * t + 4 + k];
        if (n < 0 || tt[8 * n + 2] == GHOST) continue;
        const int a = tt[8 * t + k], b = tt[8 * t + (k + 1) % 3], c = tt[8 * t + (k + 2) % 3];
        int kn = 0;  // edge (b,a) in n
        while (kn < 3 && !(tt[8 * n + kn] == b && tt[8 * n + (kn + 1) % 3] == a)) ++kn;
        if (kn == 3) continue;
        const int d = tt[8 * n + (kn + 2) % 3];
        if (std::min(c, d) >= std::min(a, b) || incircle_sign(a, b, c, d) != 0) continue;
        // t = (a,b,c), n = (b,a,d)  ->  t = (a,d,c), n = (d,b,c)
        const int X = tt[8 * t + 4 + (k + 1) % 3], Y = tt[8 * t + 4 + (k + 2) % 3];
        const int Z = tt[8 * n + 4 + (kn + 1) % 3], U = tt[8 * n + 4 + (kn + 2) % 3];
        tt[8 * t] = a; tt[8 * t + 1] = d; tt[8 * t + 2] = c;
        tt[8 * t + 4] = Z; tt[8 * t + 5] = n; tt[8 * t + 6] = Y;
        tt[8 * n] = d; tt[8 * n + 1] = b; tt[8 * n + 2] = c;
        tt[8 * n + 4] = U; tt[8 * n + 5] = X; tt[8 * n + 6] = t;
        if (Z >= 0) for (int m = 0; m < 3; ++m) if (tt[8 * Z + 4 + m] == n) { tt[8 * Z + 4 + m] = t; break; }
        if (X >= 0) for (int m = 0; m < 3; ++m) if (tt[8 * X + 4 + m] == t) { tt[8 * X + 4 + m] = n; break; }
        work.push_back(t);
        work.push_back(n);
        break;
      }
    }
  }

  // pts: n points (x,y) in pixels.  Returns false when all points are collinear / fewer than 3.
  bool run(int n, const float* pts, std::vector<int>& tris, std::vector<int>& edges) {
    tris.clear();
    edges.clear();
    px.resize(n);
    py.resize(n);
    dup_of.assign(n, -1);
    for (int i = 0; i < n; ++i) {
      const long long lx = llroundf(pts[2 * i] * 64.0f), ly = llroundf(pts[2 * i + 1] * 64.0f);
      if (lx <= -(1ll << 30) || lx >= (1ll << 30) || ly <= -(1ll << 30) || ly >= (1ll << 30)) return false;  // not pixel coordinates
      px[i] = (int32_t)lx;
      py[i] = (int32_t)ly;
    }
    tt.clear(); dead.clear(); free_list.clear(); mark.clear(); vslot.clear();
    stamp = 0;
    if (n < 3) return false;
    tt.reserve(8 * (2 * (size_t)n + 16)); dead.reserve(2 * (size_t)n + 16);
    // Insertion order: biased randomised rounds (BRIO).  A fixed pseudo-random permutation is cut into
    // rounds of doubling size; inside a round the points are visited along a snake over a grid with
    // about two points per cell.  Points of a late round fall into the interior of an already
    // well-shaped triangulation (cavities of ~4 triangles, short walks from the previous point); a
    // plain raster sweep instead inserts every point on the current hull (cavities of 6+ triangles
    // full of ghosts, 4x more triangles created than survive).  Deterministic: no clock, no rand().
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    {
      uint64_t st = 0x9e3779b97f4a7c15ull;
      for (int i = n - 1; i > 0; --i) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        const int j = (int)((st >> 33) % (uint64_t)(i + 1));
        std::swap(order[i], order[j]);
      }
    }
    int64_t xmin = px[0], xmax = px[0], ymin = py[0], ymax = py[0];
    for (int i = 1; i < n; ++i) {
      xmin = std::min<int64_t>(xmin, px[i]); xmax = std::max<int64_t>(xmax, px[i]);
      ymin = std::min<int64_t>(ymin, py[i]); ymax = std::max<int64_t>(ymax, py[i]);
    }
    exact_double = (xmax - xmin) < (1 << 24) && (ymax - ymin) < (1 << 24);
    {
      std::vector<uint64_t> key(n);
      int hi = n;
      std::vector<int> starts;  // round boundaries, last round first
      while (hi > 0) {
        const int lo = hi > 64 ? hi / 2 : 0;
        starts.push_back(lo);
        hi = lo;
      }
      hi = n;
      for (size_t r = 0; r < starts.size(); ++r) {
        const int lo = starts[r], cntr = hi - lo;
        int g = 1;
        while (g * g * 2 < cntr) ++g;
        const int64_t cw = std::max<int64_t>(1, (xmax - xmin) / g + 1), ch = std::max<int64_t>(1, (ymax - ymin) / g + 1);
        for (int k = lo; k < hi; ++k) {
          const int i = order[k];
          const int64_t cy = (py[i] - ymin) / ch;
          int64_t cx = (px[i] - xmin) / cw;
          if (cy & 1) cx = g - cx;
          key[k] = ((uint64_t)(cy * (g + 2) + cx) << 32) | (uint32_t)i;
        }
        std::sort(key.begin() + lo, key.begin() + hi);
        for (int k = lo; k < hi; ++k) order[k] = (int)(key[k] & 0xffffffffu);
        hi = lo;
      }
    }
    // identical lattice points: the smallest index keeps the vertex (canonical, like delaunay_star.h),
    // the others are left unreferenced.  Open-addressing table keyed by the coordinates.
    {
      size_t cap = 16;
      while (cap < 2 * (size_t)n) cap <<= 1;
      std::vector<int> tab(cap, -1);
      bool any = false;
      for (int i = 0; i < n; ++i) {
        size_t h = ((uint64_t)(uint32_t)px[i] * 0x9e3779b97f4a7c15ull ^ (uint64_t)(uint32_t)py[i] * 0xc2b2ae3d27d4eb4full) >> 20;
        for (h &= cap - 1;; h = (h + 1) & (cap - 1)) {
          const int j = tab[h];
          if (j < 0) { tab[h] = i; break; }
          if (px[j] == px[i] && py[j] == py[i]) { dup_of[i] = j; any = true; break; }  // j < i
        }
      }
      if (any) order.erase(std::remove_if(order.begin(), order.end(), [&](int i) { return dup_of[i] >= 0; }), order.end());
    }
    // seed triangle: first point, the next distinct point, the next non-collinear point
    int i0 = order[0], i1 = -1, i2 = -1;
    size_t k1 = 1;
    for (; k1 < order.size(); ++k1)
      if (px[order[k1]] != px[i0] || py[order[k1]] != py[i0]) {
        i1 = order[k1];
        break;
      }
    if (i1 < 0) return false;
    size_t k2 = k1 + 1;
    for (; k2 < order.size(); ++k2)
      if (orientv(i0, i1, order[k2]) != 0) {
        i2 = order[k2];
        break;
      }
    if (i2 < 0) return false;
    if (orientv(i0, i1, i2) < 0) std::swap(i1, i2);
    const int t = new_tri(i0, i1, i2);
    const int g0 = new_tri(i1, i0, GHOST), g1 = new_tri(i2, i1, GHOST), g2 = new_tri(i0, i2, GHOST);
    tt[8 * t + 4] = g0; tt[8 * t + 4 + 1] = g1; tt[8 * t + 4 + 2] = g2;
    tt[8 * g0 + 4] = t; tt[8 * g1 + 4] = t; tt[8 * g2 + 4] = t;
    // ghost (a,b,inf): slot 1 = edge (b,inf) -> ghost starting at ... , slot 2 = edge (inf,a)
    // g0=(i1,i0): (b=i0,inf) continues to the ghost whose a == i0: g2=(i0,i2); (inf,a=i1): ghost whose b == i1: g1
    tt[8 * g0 + 4 + 1] = g2; tt[8 * g0 + 4 + 2] = g1;
    tt[8 * g1 + 4 + 1] = g0; tt[8 * g1 + 4 + 2] = g2;
    tt[8 * g2 + 4 + 1] = g1; tt[8 * g2 + 4 + 2] = g0;
    last = t;
    std::vector<char> done(n, 0);
    done[i0] = done[i1] = done[i2] = 1;
    for (int idx : order) {
      if (done[idx]) continue;
      done[idx] = 1;
      if (!insert(idx)) dup_of[idx] = -2;  // resolved below
    }
    canonicalize();
    // duplicates: map to the first vertex with identical lattice coordinates
    {
      bool any = false;
      for (int i = 0; i < n; ++i) any |= dup_of[i] == -2;
      if (any) {
        std::vector<int> byc(n);
        for (int i = 0; i < n; ++i) byc[i] = i;
        std::sort(byc.begin(), byc.end(), [&](int a, int b) {
          if (px[a] != px[b]) return px[a] < px[b];
          if (py[a] != py[b]) return py[a] < py[b];
          return a < b;
        });
        for (int k = 0; k < n;) {
          int e = k;
          int keep = -1;
          while (e < n && px[byc[e]] == px[byc[k]] && py[byc[e]] == py[byc[k]]) {
            if (dup_of[byc[e]] != -2 && keep < 0) keep = byc[e];
            ++e;
          }
          for (int m = k; m < e; ++m)
            if (dup_of[byc[m]] == -2) dup_of[byc[m]] = keep;
          k = e;
        }
      }
    }
    // triangles, and every edge once: from the lower-numbered of its two triangles (hull edges from
    // their only real triangle); then a counting sort by the smaller endpoint and a short insertion
    // sort inside each bucket (degree ~6) give the canonical (i<j, sorted by (i,j)) list in O(E)
    std::vector<int> ea, eb, cnt(n + 1, 0);
    ea.reserve(3 * (size_t)n);
    eb.reserve(3 * (size_t)n);
    for (int u = 0; u < (int)dead.size(); ++u) {
      if (dead[u] || tt[8 * u + 2] == GHOST) continue;
      {  // smallest vertex first (orientation kept); sorted into the canonical order below
        const int v0 = tt[8 * u], v1 = tt[8 * u + 1], v2 = tt[8 * u + 2];
        const int r = (v0 < v1 && v0 < v2) ? 0 : ((v1 < v2) ? 1 : 2);
        tris.push_back(tt[8 * u + r]);
        tris.push_back(tt[8 * u + (r + 1) % 3]);
        tris.push_back(tt[8 * u + (r + 2) % 3]);
      }
      for (int m = 0; m < 3; ++m) {
        const int nb = tt[8 * u + 4 + m];
        if (nb >= 0 && tt[8 * nb + 2] != GHOST && nb < u) continue;
        int a = tt[8 * u + m], b = tt[8 * u + (m == 2 ? 0 : m + 1)];
        if (a > b) std::swap(a, b);
        ea.push_back(a);
        eb.push_back(b);
        cnt[a + 1]++;
      }
    }
    for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    edges.assign(2 * ea.size(), 0);
    {
      std::vector<int> fill(cnt.begin(), cnt.begin() + n);
      for (size_t k = 0; k < ea.size(); ++k) {
        const int a = ea[k], b = eb[k];
        int pos = fill[a]++;
        while (pos > cnt[a] && edges[2 * (pos - 1) + 1] > b) {  // insertion sort within the bucket
          edges[2 * pos + 1] = edges[2 * (pos - 1) + 1];
          --pos;
        }
        edges[2 * pos] = a;
        edges[2 * pos + 1] = b;
      }
      for (int a = 0; a < n; ++a)
        for (int k = cnt[a]; k < cnt[a + 1]; ++k) edges[2 * k] = a;
    }
    // canonical triangle order: (v0 = smallest vertex, v1, v2) counter-clockwise, sorted by (v0, v1)
    // -- a counting sort by v0 and an insertion sort inside each bucket (a vertex starts ~2 triangles)
    {
      const size_t T = tris.size() / 3;
      std::vector<int> tc(n + 1, 0), out(tris.size());
      for (size_t t = 0; t < T; ++t) tc[tris[3 * t] + 1]++;
      for (int i = 0; i < n; ++i) tc[i + 1] += tc[i];
      std::vector<int> fill(tc.begin(), tc.begin() + n);
      for (size_t t = 0; t < T; ++t) {
        const int v0 = tris[3 * t], v1 = tris[3 * t + 1], v2 = tris[3 * t + 2];
        int pos = fill[v0]++;
        while (pos > tc[v0] && out[3 * (pos - 1) + 1] > v1) {
          out[3 * pos + 1] = out[3 * (pos - 1) + 1];
          out[3 * pos + 2] = out[3 * (pos - 1) + 2];
          --pos;
        }
        out[3 * pos] = v0; out[3 * pos + 1] = v1; out[3 * pos + 2] = v2;
      }
      for (int a = 0; a < n; ++a)
        for (int k = tc[a]; k < tc[a + 1]; ++k) out[3 * k] = a;
      tris.swap(out);
    }
    return true;
  }
};

}  // namespace fbdel
