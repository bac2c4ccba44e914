// star_sim.cc -- host simulation of the GPU triangulation (flame_ros_b200/csrc/delaunay_star.h with
// a one-lane "warp"): same predicates, same sweep, same canonical emission as the device kernels of
// delaunay_gpu.cuh.  Test infrastructure: lets the CPU test-suite check the per-vertex star
// algorithm against the host triangulator (fb_delaunay) and Qhull without a GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifdef DS_STATS
long long ds_stat_visits = 0, ds_stat_passes = 0, ds_stat_iters32 = 0;
int ds_dbg_p = -1;
#endif
#include "../../flame_ros_b200/csrc/delaunay_star.h"

extern "C" int star_sim_delaunay(int n, const float* pts, int cell_px, int32_t* tris, int32_t* n_tris,
                                 int32_t* edges, int32_t* n_edges, int32_t* max_deg) {
  *n_tris = *n_edges = 0;
  if (max_deg) *max_deg = 0;
  if (n < 1) return 0;
  std::vector<DsPt> vxy(n);
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    vxy[i].x = ds_lattice(pts[2 * i], &bad);
    vxy[i].y = ds_lattice(pts[2 * i + 1], &bad);
  }
  if (bad) return -1;
  DsIn in;
  in.n = n;
  in.vxy = vxy.data();
  in.bx0 = in.bx1 = vxy[0].x;
  in.by0 = in.by1 = vxy[0].y;
  for (int i = 1; i < n; ++i) {
    in.bx0 = std::min(in.bx0, vxy[i].x); in.bx1 = std::max(in.bx1, vxy[i].x);
    in.by0 = std::min(in.by0, vxy[i].y); in.by1 = std::max(in.by1, vxy[i].y);
  }
  int shift = 6;
  while ((1 << (shift - 6)) < cell_px) ++shift;
  // the grid covers [0, max]: points with negative coordinates fall into the border cells
  while ((std::max(in.by1, 0) >> shift) + 1 > DS_MAXROWS || (std::max(in.bx1, 0) >> shift) + 1 > 4096) ++shift;
  in.shift = shift;
  in.gx = (std::max(in.bx1, 0) >> shift) + 1;
  in.gy = (std::max(in.by1, 0) >> shift) + 1;
  const int cells = in.gx * in.gy;
  std::vector<int32_t> cell_start(cells + 1, 0), sid(n), cellof(n);
  std::vector<DsPt> sxy(n);
  in.cell_start = cell_start.data();
  for (int i = 0; i < n; ++i) {
    cellof[i] = ds_celly(in, vxy[i].y) * in.gx + ds_cellx(in, vxy[i].x);
    cell_start[cellof[i] + 1]++;
  }
  for (int c = 0; c < cells; ++c) cell_start[c + 1] += cell_start[c];
  {
    std::vector<int> fill(cell_start.begin(), cell_start.begin() + cells);
    for (int i = n - 1; i >= 0; --i) {  // any order inside a cell is allowed (exercise a non-trivial one)
      const int k = fill[cellof[i]]++;
      sxy[k] = vxy[i];
      sid[k] = i;
    }
  }
  for (int c = 0; c < cells; ++c)  // duplicates: every point with an identical one of smaller index
    for (int k = cell_start[c]; k < cell_start[c + 1]; ++k)
      for (int m = cell_start[c]; m < cell_start[c + 1]; ++m)
        if (m != k && sxy[m].x == sxy[k].x && sxy[m].y == sxy[k].y) {
          const int im = sid[m] < 0 ? ~sid[m] : sid[m], ik = sid[k] < 0 ? ~sid[k] : sid[k];
          if (im < ik && sid[k] >= 0) sid[k] = ~ik;
        }
  in.sxy = sxy.data();
  in.sid = sid.data();
  std::vector<char> dup(n, 0);
  for (int k = 0; k < n; ++k)
    if (sid[k] < 0) dup[~sid[k]] = 1;
  std::vector<int> star((size_t)n * DS_MAXD), deg(n, 0), closed(n, 0), od(n, 0), tc(n, 0);
  DsScratch scratch;
  for (int p = 0; p < n; ++p) {
    if (dup[p]) continue;
#ifdef DS_STATS
    const long long v0 = ds_stat_visits, p0 = ds_stat_passes, i0 = ds_stat_iters32;
    ds_dbg_p = (getenv("DS_DBG_P") && atoi(getenv("DS_DBG_P")) == p) ? p : -1;
    if (ds_dbg_p >= 0) printf("vertex %d at (%d,%d) cell (%d,%d) grid %dx%d shift %d bbox %d %d %d %d\n", p, vxy[p].x, vxy[p].y, ds_cellx(in, vxy[p].x), ds_celly(in, vxy[p].y), in.gx, in.gy, in.shift, in.bx0, in.by0, in.bx1, in.by1);
#endif
    const int rc = ds_star<DsSeq>(in, p, &scratch, &star[(size_t)p * DS_MAXD], &deg[p], &closed[p]);
#ifdef DS_STATS
    if (getenv("DS_STATS_DUMP"))
      printf("%d %lld %lld %lld %d %d\n", p, ds_stat_visits - v0, ds_stat_passes - p0, ds_stat_iters32 - i0, deg[p], closed[p]);
#endif
    if (rc) return rc;
    ds_counts(p, &star[(size_t)p * DS_MAXD], deg[p], closed[p], &od[p], &tc[p]);
    if (max_deg) *max_deg = std::max(*max_deg, deg[p]);
  }
  int E = 0, T = 0;
  long long sumdeg = 0;
  for (int p = 0; p < n; ++p) {
    int outs[DS_MAXD], tr[3 * DS_MAXD], t = 0;
    const int o = ds_emit(p, &star[(size_t)p * DS_MAXD], deg[p], closed[p], outs, tr, &t);
    for (int k = 0; k < o; ++k) { edges[2 * E] = p; edges[2 * E + 1] = outs[k]; ++E; }
    for (int k = 0; k < 3 * t; ++k) tris[3 * T + k] = tr[k];
    T += t;
    sumdeg += deg[p];
  }
  *n_tris = T;
  *n_edges = E;
  if (sumdeg != 2ll * E) return 100;  // asymmetric stars
  return 0;
}
