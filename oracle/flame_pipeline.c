/*
 * flame_pipeline.c -- CPU ORACLE of the whole per-frame pipeline behind flame::Flame::update
 * (call sites /root/reference/src/flame_nodelet.cc:634, src/flame_offline_tum.cc:578; stage names
 * /root/reference/src/utils.cc:143-156).  TEST INFRASTRUCTURE ONLY, PARITY UNPINNED (flame_oracle.h).
 *
 * Two parts:
 *   1. fo_delaunay: the oracle's OWN Delaunay triangulator -- a sorted sweep (points in (x, y)
 *      order, each new point joins the edges of the current hull it sees, Lawson flips restore the
 *      empty-circle property) with exact integer predicates on a 1/64 px lattice.  Written
 *      independently of the product's two triangulators (incremental Bowyer-Watson with ghost
 *      triangles on the host, per-vertex stars on the GPU); all three must produce the same
 *      canonical mesh: co-circular point sets fan out from their smallest index, identical points
 *      keep the smallest index, triangles (v0 smallest, counter-clockwise) sorted by (v0, v1),
 *      edges (i < j) sorted by (i, j).
 *   2. fo_pipeline_*: frame creation -> epipolar update of the feature pool -> projection -> graph
 *      sync (vertex selection, triangulation, carry-over of x / w / q by feature identity) -> NLTGV2-L1
 *      iterations -> dense interpolation -> on poseframes: ring insertion + grid detection; every
 *      numeric stage is the oracle function of flame_oracle.c.  It is the CPU arm of bench.py's
 *      `e2e_update` leg and the checker of tests/test_gpu_update.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "flame_oracle.h"

typedef __int128 i128;

/* ======================================================================== */
/* Delaunay: sorted sweep + Lawson flips                                     */
/* ======================================================================== */

typedef struct {
  int n;
  const int32_t *px, *py;
  int* tv;   /* [3T] vertices, counter-clockwise */
  int* tn;   /* [3T] neighbour across edge (v[k], v[k+1]); -1 = hull */
  int nt, cap;
  int *hnext, *hprev, *htri; /* hull: vertex -> next / previous hull vertex, triangle on edge (v, next) */
  int* stack;
  int nstack, stack_cap;
} fod;

static int64_t fod_orient(const fod* d, int a, int b, int c) {
  return (int64_t)(d->px[b] - d->px[a]) * (int64_t)(d->py[c] - d->py[a]) -
         (int64_t)(d->py[b] - d->py[a]) * (int64_t)(d->px[c] - d->px[a]);
}

/* > 0: p strictly inside the circle through the counter-clockwise triangle (a, b, c); 0: on it */
static int fod_incircle(const fod* d, int a, int b, int c, int p) {
  const int64_t adx = d->px[a] - d->px[p], ady = d->py[a] - d->py[p];
  const int64_t bdx = d->px[b] - d->px[p], bdy = d->py[b] - d->py[p];
  const int64_t cdx = d->px[c] - d->px[p], cdy = d->py[c] - d->py[p];
  const int64_t al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
  const int64_t ma = bdx * cdy - bdy * cdx, mb = cdx * ady - cdy * adx, mc = adx * bdy - ady * bdx;
  /* |coordinates| < 2^20: every factor is exact in double; only the three-term sum rounds */
  const double ta = (double)al * (double)ma, tb = (double)bl * (double)mb, tc = (double)cl * (double)mc;
  const double det = ta + tb + tc, bound = 8.9e-16 * (fabs(ta) + fabs(tb) + fabs(tc));
  if (det > bound) return 1;
  if (det < -bound) return -1;
  const i128 ex = (i128)al * (i128)ma + (i128)bl * (i128)mb + (i128)cl * (i128)mc;
  return ex > 0 ? 1 : (ex < 0 ? -1 : 0);
}

static int fod_new_tri(fod* d, int a, int b, int c) {
  const int t = d->nt++;
  d->tv[3 * t] = a; d->tv[3 * t + 1] = b; d->tv[3 * t + 2] = c;
  d->tn[3 * t] = d->tn[3 * t + 1] = d->tn[3 * t + 2] = -1;
  return t;
}
static int fod_edge_slot(const fod* d, int t, int a, int b) {
  for (int k = 0; k < 3; ++k)
    if (d->tv[3 * t + k] == a && d->tv[3 * t + (k + 1) % 3] == b) return k;
  return -1;
}
static void fod_push(fod* d, int t) {
  if (d->nstack == d->stack_cap) {
    d->stack_cap *= 2;
    d->stack = (int*)realloc(d->stack, sizeof(int) * (size_t)d->stack_cap);
  }
  d->stack[d->nstack++] = t;
}

/* Flip the edge shared by t (slot k) and its neighbour: t=(a,b,c), n=(b,a,e) -> t=(a,e,c), n=(e,b,c). */
static void fod_flip(fod* d, int t, int k) {
  const int n = d->tn[3 * t + k];
  const int a = d->tv[3 * t + k], b = d->tv[3 * t + (k + 1) % 3], c = d->tv[3 * t + (k + 2) % 3];
  const int kn = fod_edge_slot(d, n, b, a);
  const int e = d->tv[3 * n + (kn + 2) % 3];
  const int X = d->tn[3 * t + (k + 1) % 3], Y = d->tn[3 * t + (k + 2) % 3];
  const int Z = d->tn[3 * n + (kn + 1) % 3], U = d->tn[3 * n + (kn + 2) % 3];
  d->tv[3 * t] = a; d->tv[3 * t + 1] = e; d->tv[3 * t + 2] = c;
  d->tn[3 * t] = Z; d->tn[3 * t + 1] = n; d->tn[3 * t + 2] = Y;
  d->tv[3 * n] = e; d->tv[3 * n + 1] = b; d->tv[3 * n + 2] = c;
  d->tn[3 * n] = U; d->tn[3 * n + 1] = X; d->tn[3 * n + 2] = t;
  if (Z >= 0) d->tn[3 * Z + fod_edge_slot(d, Z, e, a)] = t;
  else d->htri[a] = t; /* hull edge (a, e) now belongs to t */
  if (X >= 0) d->tn[3 * X + fod_edge_slot(d, X, c, b)] = n;
  else d->htri[b] = n; /* hull edge (b, c) now belongs to n */
}

/* Lawson: pop triangles, flip any edge whose opposite vertex is strictly inside; with `ties` also
 * the canonical rule for co-circular quadrilaterals (diagonal to the smaller of the two minima). */
static void fod_legalize(fod* d, int ties) {
  while (d->nstack > 0) {
    const int t = d->stack[--d->nstack];
    for (int k = 0; k < 3; ++k) {
      const int n = d->tn[3 * t + k];
      if (n < 0) continue;
      const int a = d->tv[3 * t + k], b = d->tv[3 * t + (k + 1) % 3], c = d->tv[3 * t + (k + 2) % 3];
      const int kn = fod_edge_slot(d, n, b, a);
      const int e = d->tv[3 * n + (kn + 2) % 3];
      const int s = fod_incircle(d, a, b, c, e);
      int do_flip = s > 0;
      if (!do_flip && ties && s == 0) {
        const int mab = a < b ? a : b, mce = c < e ? c : e;
        do_flip = mce < mab;
      }
      if (do_flip) {
        fod_flip(d, t, k);
        fod_push(d, t);
        fod_push(d, n);
        break;
      }
    }
  }
}

/* After inserting the sweep point p every pending triangle holds p in slot 2, so only its edge in
 * slot 0 (opposite p) can be illegal; a flip leaves two such triangles. */
static void fod_legalize_apex(fod* d) {
  while (d->nstack > 0) {
    const int t = d->stack[--d->nstack];
    const int n = d->tn[3 * t];
    if (n < 0) continue;
    const int a = d->tv[3 * t], b = d->tv[3 * t + 1], c = d->tv[3 * t + 2];
    const int kn = fod_edge_slot(d, n, b, a);
    const int e = d->tv[3 * n + (kn + 2) % 3];
    if (fod_incircle(d, a, b, c, e) > 0) {
      fod_flip(d, t, 0); /* t = (a, e, p), n = (e, b, p) */
      fod_push(d, t);
      fod_push(d, n);
    }
  }
}

/* Counting sort of `cnt` records of `w` ints by their first int (< n), insertion sort by the second
 * inside each bucket (buckets hold a handful of records). */
static void fod_sort_records(int32_t* rec, int cnt, int w, int n) {
  int* start = (int*)calloc((size_t)n + 1, sizeof(int));
  int32_t* out = (int32_t*)malloc(sizeof(int32_t) * (size_t)w * (size_t)(cnt > 0 ? cnt : 1));
  for (int k = 0; k < cnt; ++k) start[rec[w * k] + 1]++;
  for (int i = 0; i < n; ++i) start[i + 1] += start[i];
  int* fill = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  memcpy(fill, start, sizeof(int) * (size_t)n);
  for (int k = 0; k < cnt; ++k) {
    const int v0 = rec[w * k];
    int pos = fill[v0]++;
    while (pos > start[v0] && out[w * (pos - 1) + 1] > rec[w * k + 1]) {
      memcpy(out + w * pos, out + w * (pos - 1), sizeof(int32_t) * (size_t)w);
      --pos;
    }
    memcpy(out + w * pos, rec + w * k, sizeof(int32_t) * (size_t)w);
  }
  memcpy(rec, out, sizeof(int32_t) * (size_t)w * (size_t)cnt);
  free(start); free(out); free(fill);
}

static const int32_t *g_sx, *g_sy;
static int fod_cmp_xy(const void* pa, const void* pb) {
  const int a = *(const int*)pa, b = *(const int*)pb;
  if (g_sx[a] != g_sx[b]) return g_sx[a] < g_sx[b] ? -1 : 1;
  if (g_sy[a] != g_sy[b]) return g_sy[a] < g_sy[b] ? -1 : 1;
  return a < b ? -1 : (a > b ? 1 : 0);
}
/* pts [2n] pixels.  tris: capacity 3*2n, edges: capacity 2*3n.  Returns 0, or -1 when degenerate
 * (fewer than 3 distinct points, all collinear, coordinates out of range). */
int fo_delaunay(int n, const float* pts, int32_t* tris, int32_t* n_tris, int32_t* edges, int32_t* n_edges) {
  *n_tris = *n_edges = 0;
  if (n < 3) return -1;
  int32_t* px = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  int32_t* py = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  int* order = (int*)malloc(sizeof(int) * (size_t)n);
  int rc = -1;
  fod d;
  memset(&d, 0, sizeof(d));
  for (int i = 0; i < n; ++i) {
    const long long lx = llroundf(pts[2 * i] * 64.0f), ly = llroundf(pts[2 * i + 1] * 64.0f);
    if (lx <= -(1ll << 20) || lx >= (1ll << 20) || ly <= -(1ll << 20) || ly >= (1ll << 20)) goto done;
    px[i] = (int32_t)lx;
    py[i] = (int32_t)ly;
    order[i] = i;
  }
  g_sx = px;
  g_sy = py;
  qsort(order, (size_t)n, sizeof(int), fod_cmp_xy);
  /* identical points: the sort puts the smallest index first; drop the others */
  int m = 0;
  for (int k = 0; k < n; ++k)
    if (m == 0 || px[order[k]] != px[order[m - 1]] || py[order[k]] != py[order[m - 1]]) order[m++] = order[k];
  if (m < 3) goto done;
  d.n = n; d.px = px; d.py = py;
  d.cap = 2 * m + 8;
  d.tv = (int*)malloc(sizeof(int) * 3 * (size_t)d.cap);
  d.tn = (int*)malloc(sizeof(int) * 3 * (size_t)d.cap);
  d.hnext = (int*)malloc(sizeof(int) * (size_t)n);
  d.hprev = (int*)malloc(sizeof(int) * (size_t)n);
  d.htri = (int*)malloc(sizeof(int) * (size_t)n);
  d.stack_cap = 256;
  d.stack = (int*)malloc(sizeof(int) * (size_t)d.stack_cap);
  /* seed: the collinear prefix order[0..k) and the first point off its line */
  int k = 2;
  while (k < m && fod_orient(&d, order[0], order[1], order[k]) == 0) ++k;
  if (k == m) goto done; /* all collinear */
  {
    const int apex = order[k];
    const int ccw = fod_orient(&d, order[0], order[1], apex) > 0;
    /* fan over the collinear chain: triangles (c_i, c_{i+1}, apex) (or reversed) */
    int prev_t = -1;
    for (int i = 0; i + 1 < k; ++i) {
      const int a = order[i], b = order[i + 1];
      const int t = ccw ? fod_new_tri(&d, a, b, apex) : fod_new_tri(&d, b, a, apex);
      if (prev_t >= 0) {
        /* shared edge (a, apex): ccw: prev=(p,a,apex) slot 1 = (a,apex); t slot 2 = (apex,a) */
        if (ccw) { d.tn[3 * prev_t + 1] = t; d.tn[3 * t + 2] = prev_t; }
        else { d.tn[3 * prev_t + 2] = t; d.tn[3 * t + 1] = prev_t; } /* prev=(a,p,apex): slot 2 = (apex,a); t=(b,a,apex): slot 1 = (a,apex) */
      }
      prev_t = t;
    }
    /* hull (counter-clockwise): ccw: c_0 -> c_1 .. -> c_{k-1} -> apex -> c_0 */
    if (ccw) {
      for (int i = 0; i + 1 < k; ++i) { d.hnext[order[i]] = order[i + 1]; d.hprev[order[i + 1]] = order[i]; d.htri[order[i]] = i; }
      d.hnext[order[k - 1]] = apex; d.hprev[apex] = order[k - 1]; d.htri[order[k - 1]] = k - 2;
      d.hnext[apex] = order[0]; d.hprev[order[0]] = apex; d.htri[apex] = 0;
    } else { /* c_{k-1} -> .. -> c_0 -> apex -> c_{k-1} */
      for (int i = k - 1; i > 0; --i) { d.hnext[order[i]] = order[i - 1]; d.hprev[order[i - 1]] = order[i]; d.htri[order[i]] = i - 1; }
      d.hnext[order[0]] = apex; d.hprev[apex] = order[0]; d.htri[order[0]] = 0;
      d.hnext[apex] = order[k - 1]; d.hprev[order[k - 1]] = apex; d.htri[apex] = k - 2;
    }
    for (int t = 0; t < d.nt; ++t) fod_push(&d, t);
    fod_legalize(&d, 0);
    /* remaining points in sweep order (order[k] was consumed as the apex) */
    int last = apex;
    for (int idx = k + 1; idx < m; ++idx) {
      const int p = order[idx];
      /* a hull vertex whose outgoing edge p sees: start at the most recent hull vertex, walk */
      /* `last` is the lexicographic maximum of the hull and p is larger still, so p sees at least one
       * of the two hull edges at `last`; the walk is a fallback that never runs far */
      int v = last;
      int guard = 0;
      if (fod_orient(&d, v, d.hnext[v], p) >= 0) v = d.hprev[last];
      while (fod_orient(&d, v, d.hnext[v], p) >= 0 && guard++ <= n) v = d.hnext[v]; /* edge (v,next) with p strictly to its right */
      if (guard > n) goto done;
      /* extend the visible chain backwards and forwards */
      int first = v;
      while (fod_orient(&d, d.hprev[first], first, p) < 0) first = d.hprev[first];
      int end = d.hnext[v];
      while (fod_orient(&d, end, d.hnext[end], p) < 0) end = d.hnext[end];
      /* visible edges: (first, next(first)), ..., (prev(end), end); new triangles (b, a, p) */
      int a = first, prev_new = -1;
      while (a != end) {
        const int b = d.hnext[a];
        const int told = d.htri[a];
        const int t = fod_new_tri(&d, b, a, p); /* (b,a,p): slot 0 = (b,a), slot 1 = (a,p), slot 2 = (p,b) */
        d.tn[3 * t] = told;
        d.tn[3 * told + fod_edge_slot(&d, told, a, b)] = t;
        if (prev_new >= 0) { /* previous = (a, a_prev, p): its slot 2 = (p, a) meets t's slot 1 = (a, p) */
          d.tn[3 * t + 1] = prev_new;
          d.tn[3 * prev_new + 2] = t;
        }
        fod_push(&d, t);
        prev_new = t;
        a = b;
      }
      /* hull: first -> p -> end; edge (first, p) belongs to the first new triangle (slot 1 = (first, p)),
       * edge (p, end) to the last one (slot 2 = (p, end)) */
      {
        int t_first = d.tn[3 * d.htri[first] + fod_edge_slot(&d, d.htri[first], first, d.hnext[first])];
        d.htri[first] = t_first;
        d.htri[p] = prev_new;
        d.hnext[first] = p; d.hprev[p] = first;
        d.hnext[p] = end; d.hprev[end] = p;
      }
      last = p;
      fod_legalize_apex(&d);
    }
  }
  /* canonical co-circular fans */
  for (int t = 0; t < d.nt; ++t) fod_push(&d, t);
  fod_legalize(&d, 1);
  /* output */
  {
    int T = d.nt, E = 0;
    for (int t = 0; t < T; ++t) {
      const int v0 = d.tv[3 * t], v1 = d.tv[3 * t + 1], v2 = d.tv[3 * t + 2];
      const int r = (v0 < v1 && v0 < v2) ? 0 : ((v1 < v2) ? 1 : 2);
      tris[3 * t] = d.tv[3 * t + r];
      tris[3 * t + 1] = d.tv[3 * t + (r + 1) % 3];
      tris[3 * t + 2] = d.tv[3 * t + (r + 2) % 3];
      for (int kk = 0; kk < 3; ++kk) {
        const int nb = d.tn[3 * t + kk];
        if (nb >= 0 && nb < t) continue;
        int a = d.tv[3 * t + kk], b = d.tv[3 * t + (kk + 1) % 3];
        if (a > b) { const int tmp = a; a = b; b = tmp; }
        edges[2 * E] = a;
        edges[2 * E + 1] = b;
        ++E;
      }
    }
    fod_sort_records(tris, T, 3, n);
    fod_sort_records(edges, E, 2, n);
    *n_tris = T;
    *n_edges = E;
    rc = T > 0 ? 0 : -1;
  }
done:
  free(px); free(py); free(order);
  free(d.tv); free(d.tn); free(d.hnext); free(d.hprev); free(d.htri); free(d.stack);
  return rc;
}

/* ======================================================================== */
/* The per-frame pipeline                                                    */
/* ======================================================================== */

struct fo_pipeline {
  int W, H, n_slots, maxF, maxV, nthreads;
  float K[9];
  fo_update_params up;
  fo_epi_params ep;
  uint8_t* imgs;  /* [n_slots][H][W] */
  float* poses;   /* [n_slots][7] */
  /* feature pool */
  float *u_ref, *mu, *var, *u_cur, *mu_cur, *var_cur, *u_cmp;
  int32_t *ref_slot, *dropouts, *alive, *valid, *status;
  int32_t counters[FO_NUM_COUNTERS];
  int* pf_img_id;
  int pf_next, have_pf;
  /* graph of the last successful frame */
  int V, E, T, have_graph;
  int32_t *vert_feat, *edges, *tris;
  float *pos, *alpha, *beta, *z, *wt;
  float *x, *w1, *w2, *xb, *w1b, *w2b, *q1, *q2, *q3;
  float* idmap;
  /* scratch */
  int32_t *n_vfeat, *n_edges, *n_tris, *f2v;
  float *n_pos, *n_alpha, *n_beta, *n_z, *n_wt, *n_st[9];
  float* mag;
  uint8_t* occ;
  float* det_xy;
  int32_t* det_ok;
  double ms[FO_STAGE_NUM];
};

static double fo_now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

void fo_default_update_params(fo_update_params* p) {
  /* /root/reference/cfg/flame_nodelet.yaml:69-71,82-89 */
  p->detection_win_size = 16;
  p->min_grad_mag = 5.0f;
  p->detection_border = 8;
  p->idepth_init = 0.5f;
  p->idepth_var_init = 0.25f;
  p->idepth_var_max_graph = 0.01f;
  p->adaptive_data_weights = 0;
  p->init_with_prediction = 1;
  p->do_nltgv2 = 1;
  p->iters = 50;
  p->rparams.data_factor = 0.15f; p->rparams.step_x = 0.001f; p->rparams.step_q = 125.0f;
  p->rparams.theta = 0.25f; p->rparams.x_min = 0.0f; p->rparams.x_max = 10.0f;
  /* /root/reference/cfg/flame_nodelet.yaml:67,70,83,90-92 */
  p->rescale_data = 0;
  p->min_height = -1e14f;
  p->max_height = 1e14f;
  p->check_sticky_obstacles = 0;
  p->min_error = 100.0f;
  p->do_letterbox = 0;
}

#define FO_ALLOC(ptr, type, count) ptr = (type*)calloc((size_t)(count) > 0 ? (size_t)(count) : 1, sizeof(type))

fo_pipeline* fo_pipeline_create(int W, int H, const float* K, int n_slots, int max_features, int max_vertices,
                                const fo_update_params* up, const fo_epi_params* ep, int nthreads) {
  if (W < 16 || H < 16 || n_slots < 3 || max_features < 1 || max_vertices < 3 || !K || !up || !ep) return NULL;
  if (up->check_sticky_obstacles != 0 || up->min_error != 100.0f) return NULL; /* not restated: defaults only */
  fo_pipeline* P = (fo_pipeline*)calloc(1, sizeof(fo_pipeline));
  P->W = W; P->H = H; P->n_slots = n_slots; P->maxF = max_features; P->maxV = max_vertices;
  P->nthreads = nthreads < 1 ? 1 : nthreads;
  memcpy(P->K, K, sizeof(float) * 9);
  P->up = *up;
  P->ep = *ep;
  const size_t npx = (size_t)W * H;
  const int F = max_features, V = max_vertices, E = 3 * max_vertices, T = 2 * max_vertices;
  FO_ALLOC(P->imgs, uint8_t, (size_t)n_slots * npx);
  FO_ALLOC(P->poses, float, n_slots * 7);
  for (int k = 0; k < n_slots; ++k) P->poses[7 * k + 3] = 1.0f;
  FO_ALLOC(P->u_ref, float, 2 * F); FO_ALLOC(P->mu, float, F); FO_ALLOC(P->var, float, F);
  FO_ALLOC(P->u_cur, float, 2 * F); FO_ALLOC(P->mu_cur, float, F); FO_ALLOC(P->var_cur, float, F);
  FO_ALLOC(P->u_cmp, float, 2 * F);
  FO_ALLOC(P->ref_slot, int32_t, F); FO_ALLOC(P->dropouts, int32_t, F); FO_ALLOC(P->alive, int32_t, F);
  FO_ALLOC(P->valid, int32_t, F); FO_ALLOC(P->status, int32_t, F);
  FO_ALLOC(P->pf_img_id, int, n_slots - 1);
  for (int k = 0; k < n_slots - 1; ++k) P->pf_img_id[k] = -1;
  FO_ALLOC(P->vert_feat, int32_t, V); FO_ALLOC(P->edges, int32_t, 2 * E); FO_ALLOC(P->tris, int32_t, 3 * T);
  FO_ALLOC(P->pos, float, 2 * V); FO_ALLOC(P->alpha, float, E); FO_ALLOC(P->beta, float, E);
  FO_ALLOC(P->z, float, V); FO_ALLOC(P->wt, float, V);
  FO_ALLOC(P->x, float, V); FO_ALLOC(P->w1, float, V); FO_ALLOC(P->w2, float, V);
  FO_ALLOC(P->xb, float, V); FO_ALLOC(P->w1b, float, V); FO_ALLOC(P->w2b, float, V);
  FO_ALLOC(P->q1, float, E); FO_ALLOC(P->q2, float, E); FO_ALLOC(P->q3, float, E);
  FO_ALLOC(P->idmap, float, npx);
  FO_ALLOC(P->n_vfeat, int32_t, V); FO_ALLOC(P->n_edges, int32_t, 2 * E + 8); FO_ALLOC(P->n_tris, int32_t, 3 * T + 8);
  FO_ALLOC(P->f2v, int32_t, F);
  FO_ALLOC(P->n_pos, float, 2 * V); FO_ALLOC(P->n_alpha, float, E); FO_ALLOC(P->n_beta, float, E);
  FO_ALLOC(P->n_z, float, V); FO_ALLOC(P->n_wt, float, V);
  for (int k = 0; k < 6; ++k) FO_ALLOC(P->n_st[k], float, V);
  for (int k = 6; k < 9; ++k) FO_ALLOC(P->n_st[k], float, E);
  FO_ALLOC(P->mag, float, npx);
  FO_ALLOC(P->occ, uint8_t, (W / 4) * (H / 4));
  FO_ALLOC(P->det_xy, float, 2 * (W / 4) * (H / 4));
  FO_ALLOC(P->det_ok, int32_t, (W / 4) * (H / 4));
  return P;
}

void fo_pipeline_destroy(fo_pipeline* P) {
  if (!P) return;
  free(P->imgs); free(P->poses); free(P->u_ref); free(P->mu); free(P->var); free(P->u_cur); free(P->mu_cur);
  free(P->var_cur); free(P->u_cmp); free(P->ref_slot); free(P->dropouts); free(P->alive); free(P->valid);
  free(P->status); free(P->pf_img_id); free(P->vert_feat); free(P->edges); free(P->tris); free(P->pos);
  free(P->alpha); free(P->beta); free(P->z); free(P->wt); free(P->x); free(P->w1); free(P->w2); free(P->xb);
  free(P->w1b); free(P->w2b); free(P->q1); free(P->q2); free(P->q3); free(P->idmap); free(P->n_vfeat);
  free(P->n_edges); free(P->n_tris); free(P->f2v); free(P->n_pos); free(P->n_alpha); free(P->n_beta);
  free(P->n_z); free(P->n_wt);
  for (int k = 0; k < 9; ++k) free(P->n_st[k]);
  free(P->mag); free(P->occ); free(P->det_xy); free(P->det_ok);
  free(P);
}

/* Ring insertion of the current frame + grid detection in the cells without a live feature. */
static void fo_new_poseframe(fo_pipeline* P, int img_id) {
  const fo_update_params* up = &P->up;
  const int cur = P->n_slots - 1, slot = P->pf_next, W = P->W, H = P->H;
  const size_t npx = (size_t)W * H;
  if (P->pf_img_id[slot] >= 0)
    for (int f = 0; f < P->maxF; ++f)
      if (P->alive[f] && P->ref_slot[f] == slot) P->alive[f] = 0;
  memcpy(P->imgs + (size_t)slot * npx, P->imgs + (size_t)cur * npx, npx);
  memcpy(P->poses + 7 * slot, P->poses + 7 * cur, sizeof(float) * 7);
  P->pf_img_id[slot] = img_id;
  P->pf_next = (slot + 1) % (P->n_slots - 1);
  P->have_pf = 1;
  const int win = up->detection_win_size, cx = W / win, cy = H / win, cells = cx * cy;
  memset(P->occ, 0, (size_t)cells);
  for (int f = 0; f < P->maxF; ++f) {
    if (!P->valid[f]) continue;
    const int i = (int)floorf(P->u_cur[2 * f]) / win, j = (int)floorf(P->u_cur[2 * f + 1]) / win;
    if (i >= 0 && j >= 0 && i < cx && j < cy) P->occ[j * cx + i] = 1;
  }
  fo_gradient_mag(W, H, P->imgs + (size_t)cur * npx, P->mag);
  fo_detect_features_rows(W, H, P->mag, win, up->detection_border, up->min_grad_mag, P->occ,
                          up->do_letterbox ? H / 3 : 0, up->do_letterbox ? (2 * H) / 3 : H, P->det_xy, P->det_ok);
  int f = 0;
  for (int c = 0; c < cells; ++c) {
    if (!P->det_ok[c]) continue;
    while (f < P->maxF && P->alive[f]) ++f;
    if (f >= P->maxF) break;
    const float dx = P->det_xy[2 * c], dy = P->det_xy[2 * c + 1];
    float m = up->idepth_init;
    if (up->init_with_prediction && P->have_graph) {
      const float q = P->idmap[(int)dy * W + (int)dx];
      if (q == q && q > 0.0f) m = q;
    }
    P->u_ref[2 * f] = dx; P->u_ref[2 * f + 1] = dy;
    P->ref_slot[f] = slot;
    P->mu[f] = m;
    P->var[f] = up->idepth_var_init;
    P->dropouts[f] = 0;
    P->alive[f] = 1;
    ++f;
  }
}

int fo_pipeline_update(fo_pipeline* P, int img_id, const float* pose, const uint8_t* gray, int is_poseframe) {
  const fo_update_params* up = &P->up;
  const int cur = P->n_slots - 1, W = P->W, H = P->H, F = P->maxF;
  const size_t npx = (size_t)W * H;
  double t0 = fo_now_ms(), t1;
  const double t_begin = t0;
  memset(P->ms, 0, sizeof(P->ms));
#define FO_LAP(stage) t1 = fo_now_ms(); P->ms[stage] += t1 - t0; t0 = t1
  memcpy(P->imgs + (size_t)cur * npx, gray, npx);
  memcpy(P->poses + 7 * cur, pose, sizeof(float) * 7);
  FO_LAP(FO_STAGE_FRAME);
  if (!P->have_pf) {
    memset(P->valid, 0, sizeof(int32_t) * (size_t)F);
    fo_new_poseframe(P, img_id);
    FO_LAP(FO_STAGE_DETECT);
    P->ms[FO_STAGE_UPDATE] = fo_now_ms() - t_begin;
    return 0;
  }
  fo_idepth_update(W, H, P->n_slots, P->imgs, P->poses, P->K, cur, F, P->ref_slot, P->u_ref, P->mu, P->var,
                   P->dropouts, P->alive, P->status, P->u_cmp, P->counters, &P->ep, P->nthreads);
  FO_LAP(FO_STAGE_IDEPTH);
  fo_project_features(W, H, P->n_slots, P->poses, P->K, cur, F, P->ref_slot, P->u_ref, P->mu, P->var, P->alive,
                      P->u_cur, P->mu_cur, P->var_cur, P->valid);
  for (int f = 0; f < F; ++f)
    if (P->alive[f] && !P->valid[f]) P->alive[f] = 0;
  FO_LAP(FO_STAGE_PROJECT);
  /* ---- graph sync: vertex selection in ascending feature index */
  int V = 0;
  const int use_height = up->min_height > -1e13f || up->max_height < 1e13f;
  float r20 = 0.f, r21 = 0.f, r22 = 1.f;
  {
    /* third row of the camera-to-world rotation of the current pose, fp32, fixed expression order */
    const float* q = P->poses + 7 * cur;
    const float qx = q[0], qy = q[1], qz = q[2], qw = q[3];
    const float n = qx * qx + qy * qy + qz * qz + qw * qw;
    const float s2 = 2.0f / n;
    r20 = qx * qz * s2 - qw * qy * s2;
    r21 = qy * qz * s2 + qw * qx * s2;
    r22 = 1.0f - (qx * qx * s2 + qy * qy * s2);
  }
  for (int f = 0; f < F && V < P->maxV; ++f)
    if (P->valid[f] && P->var_cur[f] < up->idepth_var_max_graph) {
      if (use_height) {
        const float zc = 1.0f / P->mu_cur[f];
        const float xc = ((P->u_cur[2 * f] - P->K[2]) / P->K[0]) * zc;
        const float yc = ((P->u_cur[2 * f + 1] - P->K[5]) / P->K[4]) * zc;
        const float h = fmaf(r20, xc, fmaf(r21, yc, fmaf(r22, zc, P->poses[7 * cur + 6])));
        if (!(h >= up->min_height && h <= up->max_height)) continue;
      }
      P->n_vfeat[V] = f;
      P->n_pos[2 * V] = P->u_cur[2 * f];
      P->n_pos[2 * V + 1] = P->u_cur[2 * f + 1];
      ++V;
    }
  FO_LAP(FO_STAGE_SYNC);
  int updated = 0, nT = 0, nE = 0;
  int have_tri = V >= 3 && fo_delaunay(V, P->n_pos, P->n_tris, &nT, P->n_edges, &nE) == 0;
  FO_LAP(FO_STAGE_TRIANGULATE);
  if (have_tri) {
    float** st = P->n_st; /* x w1 w2 xb w1b w2b | q1 q2 q3 */
    for (int k = 0; k < 6; ++k) memset(st[k], 0, sizeof(float) * (size_t)V);
    for (int k = 6; k < 9; ++k) memset(st[k], 0, sizeof(float) * (size_t)nE);
    for (int f = 0; f < F; ++f) P->f2v[f] = -1;
    if (P->have_graph)
      for (int k = 0; k < P->V; ++k) P->f2v[P->vert_feat[k]] = k;
    for (int k = 0; k < V; ++k) {
      const int f = P->n_vfeat[k];
      P->n_z[k] = P->mu_cur[f];
      P->n_wt[k] = up->adaptive_data_weights ? (1.0f / P->var_cur[f]) : 1.0f;
      const int o = P->f2v[f];
      if (o >= 0) {
        st[0][k] = P->x[o]; st[1][k] = P->w1[o]; st[2][k] = P->w2[o];
        st[3][k] = P->xb[o]; st[4][k] = P->w1b[o]; st[5][k] = P->w2b[o];
      } else {
        float x0 = P->n_z[k];
        if (up->init_with_prediction && P->have_graph) {
          const int px = (int)rintf(P->n_pos[2 * k]), py = (int)rintf(P->n_pos[2 * k + 1]);
          if (px >= 0 && py >= 0 && px < W && py < H) {
            const float p = P->idmap[py * W + px];
            if (p == p && p > 0.0f) x0 = p;
          }
        }
        st[0][k] = st[3][k] = x0;
      }
    }
    /* persisting edges keep their dual: both edge lists are sorted by (feature_i, feature_j) */
    {
      int o = 0;
      for (int e = 0; e < nE; ++e) {
        const int i = P->n_edges[2 * e], j = P->n_edges[2 * e + 1];
        const float dx = P->n_pos[2 * i] - P->n_pos[2 * j], dy = P->n_pos[2 * i + 1] - P->n_pos[2 * j + 1];
        P->n_alpha[e] = 1.0f / sqrtf(dx * dx + dy * dy);
        P->n_beta[e] = 1.0f;
        if (!P->have_graph) continue;
        const int64_t key = ((int64_t)P->n_vfeat[i] << 32) | (uint32_t)P->n_vfeat[j];
        while (o < P->E && (((int64_t)P->vert_feat[P->edges[2 * o]] << 32) | (uint32_t)P->vert_feat[P->edges[2 * o + 1]]) < key) ++o;
        if (o < P->E && (((int64_t)P->vert_feat[P->edges[2 * o]] << 32) | (uint32_t)P->vert_feat[P->edges[2 * o + 1]]) == key) {
          st[6][e] = P->q1[o]; st[7][e] = P->q2[o]; st[8][e] = P->q3[o];
        }
      }
    }
    FO_LAP(FO_STAGE_SYNC);
    if (up->do_nltgv2 && up->iters > 0) {
      float scale = 1.0f;
      if (up->rescale_data) { /* mean of the data term, accumulated in double */
        double sum = 0.0;
        for (int k = 0; k < V; ++k) sum += (double)P->n_z[k];
        scale = (float)(sum / (double)V);
        if (!(scale > 0.0f)) scale = 1.0f;
        for (int k = 0; k < V; ++k) {
          P->n_wt[k] = P->n_wt[k]; /* weights are not rescaled */
          st[0][k] /= scale; st[1][k] /= scale; st[2][k] /= scale;
          st[3][k] /= scale; st[4][k] /= scale; st[5][k] /= scale;
        }
        memcpy(P->mag, P->n_z, sizeof(float) * (size_t)V); /* original data term kept aside (scratch) */
        for (int k = 0; k < V; ++k) P->n_z[k] /= scale;
      }
      fo_nltgv2_solve(V, nE, P->n_pos, P->n_edges, P->n_alpha, P->n_beta, P->n_z, P->n_wt, st[0], st[1], st[2], st[3],
                      st[4], st[5], st[6], st[7], st[8], &up->rparams, up->iters, P->nthreads);
      if (up->rescale_data) {
        for (int k = 0; k < V; ++k)
          for (int a = 0; a < 6; ++a) st[a][k] *= scale;
        memcpy(P->n_z, P->mag, sizeof(float) * (size_t)V);
      }
    }
    FO_LAP(FO_STAGE_SOLVE);
    fo_rasterize_idepth_mt(W, H, V, P->n_pos, st[0], nT, P->n_tris, NULL, P->idmap, P->nthreads);
    FO_LAP(FO_STAGE_INTERP);
    /* commit */
    P->V = V; P->E = nE; P->T = nT; P->have_graph = 1;
    memcpy(P->vert_feat, P->n_vfeat, sizeof(int32_t) * (size_t)V);
    memcpy(P->edges, P->n_edges, sizeof(int32_t) * 2 * (size_t)nE);
    memcpy(P->tris, P->n_tris, sizeof(int32_t) * 3 * (size_t)nT);
    memcpy(P->pos, P->n_pos, sizeof(float) * 2 * (size_t)V);
    memcpy(P->alpha, P->n_alpha, sizeof(float) * (size_t)nE);
    memcpy(P->beta, P->n_beta, sizeof(float) * (size_t)nE);
    memcpy(P->z, P->n_z, sizeof(float) * (size_t)V);
    memcpy(P->wt, P->n_wt, sizeof(float) * (size_t)V);
    float* dstv[6] = {P->x, P->w1, P->w2, P->xb, P->w1b, P->w2b};
    float* dste[3] = {P->q1, P->q2, P->q3};
    for (int k = 0; k < 6; ++k) memcpy(dstv[k], st[k], sizeof(float) * (size_t)V);
    for (int k = 0; k < 3; ++k) memcpy(dste[k], st[6 + k], sizeof(float) * (size_t)nE);
    FO_LAP(FO_STAGE_SYNC);
    updated = 1;
  }
  if (is_poseframe) {
    fo_new_poseframe(P, img_id);
    FO_LAP(FO_STAGE_DETECT);
  }
  P->ms[FO_STAGE_UPDATE] = fo_now_ms() - t_begin;
  return updated;
}

void fo_pipeline_sizes(const fo_pipeline* P, int32_t* V, int32_t* T, int32_t* E) {
  *V = P->have_graph ? P->V : 0;
  *T = P->have_graph ? P->T : 0;
  *E = P->have_graph ? P->E : 0;
}

void fo_pipeline_mesh(const fo_pipeline* P, float* vtx, float* idepth, int32_t* tris, int32_t* edges, int32_t* vert_feat) {
  if (!P->have_graph) return;
  if (vtx) memcpy(vtx, P->pos, sizeof(float) * 2 * (size_t)P->V);
  if (idepth) memcpy(idepth, P->x, sizeof(float) * (size_t)P->V);
  if (tris) memcpy(tris, P->tris, sizeof(int32_t) * 3 * (size_t)P->T);
  if (edges) memcpy(edges, P->edges, sizeof(int32_t) * 2 * (size_t)P->E);
  if (vert_feat) memcpy(vert_feat, P->vert_feat, sizeof(int32_t) * (size_t)P->V);
}

/* getInverseDepthMap (filter == NULL) / getFilteredInverseDepthMap: out [H*W], NaN = no depth. */
void fo_pipeline_idepthmap(const fo_pipeline* P, const fo_tri_filter_params* filter, float* out) {
  const size_t npx = (size_t)P->W * P->H;
  if (!P->have_graph) {
    for (size_t i = 0; i < npx; ++i) out[i] = NAN;
    return;
  }
  if (!filter) {
    memcpy(out, P->idmap, sizeof(float) * npx);
    return;
  }
  uint8_t* valid = (uint8_t*)malloc((size_t)P->T > 0 ? (size_t)P->T : 1);
  fo_triangle_validity(P->W, P->H, P->K, P->V, P->pos, P->x, P->T, P->tris, filter, valid);
  fo_rasterize_idepth_mt(P->W, P->H, P->V, P->pos, P->x, P->T, P->tris, valid, out, P->nthreads);
  free(valid);
}

void fo_pipeline_features(const fo_pipeline* P, float* u_ref, int32_t* ref_slot, float* mu, float* var,
                          int32_t* dropouts, int32_t* alive, int32_t* valid) {
  const size_t F = (size_t)P->maxF;
  if (u_ref) memcpy(u_ref, P->u_ref, sizeof(float) * 2 * F);
  if (ref_slot) memcpy(ref_slot, P->ref_slot, sizeof(int32_t) * F);
  if (mu) memcpy(mu, P->mu, sizeof(float) * F);
  if (var) memcpy(var, P->var, sizeof(float) * F);
  if (dropouts) memcpy(dropouts, P->dropouts, sizeof(int32_t) * F);
  if (alive) memcpy(alive, P->alive, sizeof(int32_t) * F);
  if (valid) memcpy(valid, P->valid, sizeof(int32_t) * F);
}

void fo_pipeline_stage_ms(const fo_pipeline* P, double* ms /*[FO_STAGE_NUM]*/) {
  memcpy(ms, P->ms, sizeof(P->ms));
}
