/*
 * flame_oracle.c -- CPU ORACLE for the FLaME hot path.  TEST INFRASTRUCTURE ONLY.
 * PARITY UNPINNED -- see flame_oracle.h for provenance and the rules on who may
 * link this file.  Build: `make -C oracle` (gcc -O2 -ffp-contract=off -fopenmp).
 *
 * Everything here is a restatement written for this repository; nothing is
 * derived from reference sources (the reference tree does not contain the
 * algorithm).  Parameter names/defaults cite /root/reference/cfg/flame_nodelet.yaml.
 */
#include "flame_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ======================================================================== */
/* NLTGV2-L1 primal-dual (rparams at                  */
/* /root/reference/src/flame_nodelet.cc:256-259)                            */
/* ======================================================================== */

/* fminf / fmaxf as inline code: with -fno-fast-math gcc emits calls into libm for them (40 call sites
 * in the hot loops: 3.4x on the solver).  Semantics kept exactly: a NaN operand yields the other one,
 * and of two zeros the minimum is -0, the maximum +0 (what libm and the CUDA intrinsics return). */
static inline float fo_fminf(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  return (a < b || (a == b && __builtin_signbit(a))) ? a : b;
}
static inline float fo_fmaxf(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  return (a > b || (a == b && !__builtin_signbit(a))) ? a : b;
}
#define fminf fo_fminf
#define fmaxf fo_fmaxf

static inline float clamp1(float t) { return fminf(fmaxf(t, -1.0f), 1.0f); }

/* Build CSR incidence (ascending edge id per vertex). inc = (edge<<1)|role, role 0 = source. */
static void build_csr(int V, int E, const int32_t* ij, int32_t* row, int32_t* inc) {
  memset(row, 0, sizeof(int32_t) * (size_t)(V + 1));
  for (int e = 0; e < E; ++e) {
    row[ij[2 * e] + 1]++;
    row[ij[2 * e + 1] + 1]++;
  }
  for (int v = 0; v < V; ++v) row[v + 1] += row[v];
  int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)(V > 0 ? V : 1));
  memcpy(fill, row, sizeof(int32_t) * (size_t)V);
  for (int e = 0; e < E; ++e) {
    inc[fill[ij[2 * e]]++] = (e << 1);
    inc[fill[ij[2 * e + 1]]++] = (e << 1) | 1;
  }
  free(fill);
}

void fo_nltgv2_solve(int V, int E, const float* pos, const int32_t* ij,
                     const float* alpha, const float* beta, const float* z,
                     const float* wt, float* x, float* w1, float* w2, float* xb,
                     float* w1b, float* w2b, float* q1, float* q2, float* q3,
                     const fo_nltgv2_params* p, int iters, int nthreads) {
  if (V <= 0 || iters <= 0) return;
  int32_t* row = (int32_t*)malloc(sizeof(int32_t) * (size_t)(V + 1));
  int32_t* inc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * E + 1));
  float* dx = (float*)malloc(sizeof(float) * (size_t)(E + 1));
  float* dy = (float*)malloc(sizeof(float) * (size_t)(E + 1));
  build_csr(V, E, ij, row, inc);
  for (int e = 0; e < E; ++e) {
    int i = ij[2 * e], j = ij[2 * e + 1];
    dx[e] = pos[2 * i] - pos[2 * j];
    dy[e] = pos[2 * i + 1] - pos[2 * j + 1];
  }
  const float sigma = p->step_q, tau = p->step_x, theta = p->theta;
  const float tl = p->step_x * p->data_factor;
  const float xmin = p->x_min, xmax = p->x_max;
  (void)nthreads;
#ifdef _OPENMP
  int nt = nthreads > 1 ? nthreads : 1;
#endif
  for (int it = 0; it < iters; ++it) {
    /* dual half-step over edges */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nt) if (nt > 1)
#endif
    for (int e = 0; e < E; ++e) {
      int i = ij[2 * e], j = ij[2 * e + 1];
      float t = xb[i] - xb[j];
      t = fmaf(-dx[e], w1b[i], t);
      t = fmaf(-dy[e], w2b[i], t);
      float k1 = alpha[e] * t;
      float k2 = beta[e] * (w1b[i] - w1b[j]);
      float k3 = beta[e] * (w2b[i] - w2b[j]);
      q1[e] = clamp1(fmaf(sigma, k1, q1[e]));
      q2[e] = clamp1(fmaf(sigma, k2, q2[e]));
      q3[e] = clamp1(fmaf(sigma, k3, q3[e]));
    }
    /* primal half-step + L1 prox + box + extragradient over vertices */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nt) if (nt > 1)
#endif
    for (int v = 0; v < V; ++v) {
      float gx = 0.0f, g1 = 0.0f, g2 = 0.0f;
      for (int s = row[v]; s < row[v + 1]; ++s) {
        int e = inc[s] >> 1;
        float a1 = alpha[e] * q1[e];
        if ((inc[s] & 1) == 0) {
          gx += a1;
          g1 += fmaf(beta[e], q2[e], -(dx[e] * a1));
          g2 += fmaf(beta[e], q3[e], -(dy[e] * a1));
        } else {
          gx -= a1;
          g1 -= beta[e] * q2[e];
          g2 -= beta[e] * q3[e];
        }
      }
      float xo = x[v], w1o = w1[v], w2o = w2[v];
      float xp = fmaf(-tau, gx, xo);
      float w1n = fmaf(-tau, g1, w1o);
      float w2n = fmaf(-tau, g2, w2o);
      float th = tl * wt[v];
      float d = xp - z[v];
      float xn = (d > th) ? (xp - th) : ((d < -th) ? (xp + th) : z[v]);
      xn = fminf(fmaxf(xn, xmin), xmax);
      x[v] = xn;
      w1[v] = w1n;
      w2[v] = w2n;
      xb[v] = fmaf(theta, xn - xo, xn);
      w1b[v] = fmaf(theta, w1n - w1o, w1n);
      w2b[v] = fmaf(theta, w2n - w2o, w2n);
    }
  }
  free(row);
  free(inc);
  free(dx);
  free(dy);
}

void fo_nltgv2_costs(int V, int E, const float* pos, const int32_t* ij,
                     const float* alpha, const float* beta, const float* z,
                     const float* wt, const float* x, const float* w1,
                     const float* w2, float data_factor, double* smoothness,
                     double* data) {
  double s = 0.0, d = 0.0;
  for (int e = 0; e < E; ++e) {
    int i = ij[2 * e], j = ij[2 * e + 1];
    float dx = pos[2 * i] - pos[2 * j];
    float dy = pos[2 * i + 1] - pos[2 * j + 1];
    float t = x[i] - x[j];
    t = fmaf(-dx, w1[i], t);
    t = fmaf(-dy, w2[i], t);
    float k1 = alpha[e] * t;
    float k2 = beta[e] * (w1[i] - w1[j]);
    float k3 = beta[e] * (w2[i] - w2[j]);
    s += (double)(fabsf(k1) + fabsf(k2) + fabsf(k3));
  }
  for (int v = 0; v < V; ++v) d += (double)((data_factor * wt[v]) * fabsf(x[v] - z[v]));
  *smoothness = s;
  *data = d;
}

/* ======================================================================== */
/* Epipolar geometry                                                         */
/* ======================================================================== */

static void quat_to_R(const float* q, float* R) {
  float x = q[0], y = q[1], z = q[2], w = q[3];
  float n = x * x + y * y + z * z + w * w;
  float s = 2.0f / n;
  float xx = x * x * s, yy = y * y * s, zz = z * z * s;
  float xy = x * y * s, xz = x * z * s, yz = y * z * s;
  float wx = w * x * s, wy = w * y * s, wz = w * z * s;
  R[0] = 1.0f - (yy + zz); R[1] = xy - wz;          R[2] = xz + wy;
  R[3] = xy + wz;          R[4] = 1.0f - (xx + zz); R[5] = yz - wx;
  R[6] = xz - wy;          R[7] = yz + wx;          R[8] = 1.0f - (xx + yy);
}

void fo_epi_geometry(const float* K, const float* pose_ref, const float* pose_cmp, float* G) {
  float Rr[9], Rc[9], R[9], t[3], d[3];
  quat_to_R(pose_ref, Rr);
  quat_to_R(pose_cmp, Rc);
  /* R = Rc^T Rr ; t = Rc^T (t_r - t_c) */
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      R[3 * r + c] = Rc[0 + r] * Rr[0 + c] + Rc[3 + r] * Rr[3 + c] + Rc[6 + r] * Rr[6 + c];
  for (int k = 0; k < 3; ++k) d[k] = pose_ref[4 + k] - pose_cmp[4 + k];
  for (int r = 0; r < 3; ++r) t[r] = Rc[0 + r] * d[0] + Rc[3 + r] * d[1] + Rc[6 + r] * d[2];
  const float fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  /* M = K R (K = [fx 0 cx; 0 fy cy; 0 0 1]) */
  float M[9];
  for (int c = 0; c < 3; ++c) {
    M[c] = fx * R[c] + cx * R[6 + c];
    M[3 + c] = fy * R[3 + c] + cy * R[6 + c];
    M[6 + c] = R[6 + c];
  }
  /* A = M Kinv, Kinv = [1/fx 0 -cx/fx; 0 1/fy -cy/fy; 0 0 1] */
  for (int r = 0; r < 3; ++r) {
    float a0 = M[3 * r] / fx, a1 = M[3 * r + 1] / fy;
    G[3 * r] = a0;
    G[3 * r + 1] = a1;
    G[3 * r + 2] = M[3 * r + 2] - (a0 * cx + a1 * cy);
  }
  /* b = K t */
  G[9] = fx * t[0] + cx * t[2];
  G[10] = fy * t[1] + cy * t[2];
  G[11] = t[2];
  /* c = -R^T t (cmp centre in ref coordinates); e = K c */
  float c3[3];
  for (int k = 0; k < 3; ++k) c3[k] = -(R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2]);
  G[12] = fx * c3[0] + cx * c3[2];
  G[13] = fy * c3[1] + cy * c3[2];
  G[14] = c3[2];
}

/* ======================================================================== */
/* Epipolar inverse-depth update                      */
/* ======================================================================== */

#define FO_MAX_WIN 15
#define FO_MAX_SEARCH 256

static inline int inside_img(float x, float y, int W, int H) {
  return x >= 0.0f && y >= 0.0f && x < (float)(W - 1) && y < (float)(H - 1);
}

static inline float bilin(const uint8_t* img, int W, float x, float y) {
  float xf = floorf(x), yf = floorf(y);
  int x0 = (int)xf, y0 = (int)yf;
  float fx = x - xf, fy = y - yf;
  const uint8_t* p = img + (size_t)y0 * (size_t)W + (size_t)x0;
  float i00 = (float)p[0], i10 = (float)p[1], i01 = (float)p[W], i11 = (float)p[W + 1];
  float a = fmaf(fx, i10 - i00, i00);
  float b = fmaf(fx, i11 - i01, i01);
  return fmaf(fy, b - a, a);
}

/* inverse depth of the point on the epipolar line at pixel (x,y), dominant axis */
static inline float idepth_at(float x, float y, int use_x, float P0x, float P0y, float P0z,
                              float bx, float by, float bz) {
  if (use_x) return fmaf(x, P0z, -P0x) / fmaf(-x, bz, bx);
  return fmaf(y, P0z, -P0y) / fmaf(-y, bz, by);
}

static int update_one(int W, int H, const uint8_t* iref, const uint8_t* icmp, const float* G,
                      float ux, float uy, float* mu_io, float* var_io, float* ucmp,
                      const fo_epi_params* p) {
  const int win = p->win_size, h = win / 2;
  const float bx = G[9], by = G[10], bz = G[11];
  const float P0x = fmaf(G[0], ux, fmaf(G[1], uy, G[2]));
  const float P0y = fmaf(G[3], ux, fmaf(G[4], uy, G[5]));
  const float P0z = fmaf(G[6], ux, fmaf(G[7], uy, G[8]));
  const float m = *mu_io, v = *var_io;
  const float sigma = sqrtf(v);
  float xi_lo = fmaxf(fmaf(-p->search_sigma, sigma, m), p->idepth_min);
  float xi_hi = fminf(fmaf(p->search_sigma, sigma, m), p->idepth_max);
  xi_hi = fmaxf(xi_hi, xi_lo);
  const float pz_mu = fmaf(m, bz, P0z);
  const float pz_lo = fmaf(xi_lo, bz, P0z);
  const float pz_hi = fmaf(xi_hi, bz, P0z);
  if (!(pz_mu > 1e-6f) || !(pz_lo > 1e-6f) || !(pz_hi > 1e-6f)) return FO_FAIL_OUT_OF_IMAGE;
  const float umx = fmaf(m, bx, P0x) / pz_mu, umy = fmaf(m, by, P0y) / pz_mu;
  /* direction of increasing idepth along the epipolar line in cmp */
  const float dxv = fmaf(-umx, bz, bx), dyv = fmaf(-umy, bz, by);
  const float dn = sqrtf(fmaf(dxv, dxv, dyv * dyv));
  const float gpar = dn / pz_mu;
  if (!(gpar >= p->min_parallax)) return FO_NO_PARALLAX;
  const float lx = dxv / dn, ly = dyv / dn;
  const float ulx = fmaf(xi_lo, bx, P0x) / pz_lo, uly = fmaf(xi_lo, by, P0y) / pz_lo;
  const float uhx = fmaf(xi_hi, bx, P0x) / pz_hi, uhy = fmaf(xi_hi, by, P0y) / pz_hi;
  float s_lo = fmaf(ulx - umx, lx, (uly - umy) * ly);
  float s_hi = fmaf(uhx - umx, lx, (uhy - umy) * ly);
  const float half = 0.5f * (float)(p->max_search_px - 4);
  s_lo = fminf(fmaxf(s_lo, -half), 0.0f);
  s_hi = fmaxf(fminf(s_hi, half), 0.0f);
  const float s0 = floorf(s_lo) - 1.0f;
  int n_steps = (int)(ceilf(s_hi) - s0) + 2;
  if (n_steps > p->max_search_px) n_steps = p->max_search_px;

  /* reference patch along the epipolar direction in ref */
  float lrx = fmaf(ux, G[14], -G[12]), lry = fmaf(uy, G[14], -G[13]);
  const float lrn = sqrtf(fmaf(lrx, lrx, lry * lry));
  if (!(lrn > 1e-12f)) return FO_NO_PARALLAX;
  lrx = lrx / lrn;
  lry = lry / lrn;
  float ref[FO_MAX_WIN];
  for (int k = 0; k < win; ++k) {
    float kk = (float)(k - h);
    float x = fmaf(kk, lrx, ux), y = fmaf(kk, lry, uy);
    if (!inside_img(x, y, W, H)) return FO_FAIL_OUT_OF_IMAGE;
    ref[k] = bilin(iref, W, x, y);
  }
  float grad2 = 0.0f;
  for (int k = 0; k + 1 < win; ++k) {
    float d = ref[k + 1] - ref[k];
    grad2 = fmaf(d, d, grad2);
  }
  grad2 = grad2 / (float)(win - 1);
  if (grad2 < p->min_grad_mag * p->min_grad_mag) return FO_FAIL_REF_PATCH_GRADIENT;
  /* 2-D image gradient at the reference pixel */
  if (!inside_img(ux - 1.0f, uy - 1.0f, W, H) || !inside_img(ux + 1.0f, uy + 1.0f, W, H))
    return FO_FAIL_OUT_OF_IMAGE;
  const float gx = 0.5f * (bilin(iref, W, ux + 1.0f, uy) - bilin(iref, W, ux - 1.0f, uy));
  const float gy = 0.5f * (bilin(iref, W, ux, uy + 1.0f) - bilin(iref, W, ux, uy - 1.0f));

  /* sample the comparison image once along the line */
  float line[FO_MAX_SEARCH + FO_MAX_WIN];
  uint8_t ok[FO_MAX_SEARCH + FO_MAX_WIN];
  const int n_samp = n_steps + 2 * h;
  for (int mI = 0; mI < n_samp; ++mI) {
    float s = s0 + (float)(mI - h);
    float x = fmaf(s, lx, umx), y = fmaf(s, ly, umy);
    ok[mI] = (uint8_t)inside_img(x, y, W, H);
    line[mI] = ok[mI] ? bilin(icmp, W, x, y) : 0.0f;
  }
  /* sliding SSD, arg-min (ties: smallest n) */
  float cost[FO_MAX_SEARCH];
  uint8_t cok[FO_MAX_SEARCH];
  int nbest = -1;
  float best = 0.0f;
  for (int n = 0; n < n_steps; ++n) {
    int good = 1;
    float c = 0.0f;
    for (int k = 0; k < win; ++k) {
      good &= ok[n + k];
      float d = line[n + k] - ref[k];
      c = fmaf(d, d, c);
    }
    cok[n] = (uint8_t)good;
    cost[n] = c;
    if (good && (nbest < 0 || c < best)) {
      best = c;
      nbest = n;
    }
  }
  if (nbest < 0) return FO_FAIL_OUT_OF_IMAGE;
  int have2 = 0;
  float second = 0.0f;
  for (int n = 0; n < n_steps; ++n) {
    if (!cok[n]) continue;
    int dn2 = n - nbest;
    if (dn2 < 0) dn2 = -dn2;
    if (dn2 <= p->ambiguity_radius) continue;
    if (!have2 || cost[n] < second) {
      second = cost[n];
      have2 = 1;
    }
  }
  if (best > p->max_cost * (float)win) return FO_FAIL_MAX_COST;
  const float floor_c = p->pixel_noise_var * (float)win;
  if (have2 && second < p->ambiguity_ratio * fmaxf(best, floor_c)) return FO_FAIL_AMBIGUOUS_MATCH;
  /* sub-pixel parabola */
  float delta = 0.0f;
  if (nbest > 0 && nbest + 1 < n_steps && cok[nbest - 1] && cok[nbest + 1]) {
    float cm = cost[nbest - 1], c0 = cost[nbest], cp = cost[nbest + 1];
    float den = (cm - 2.0f * c0) + cp;
    if (den > 1e-12f) {
      delta = (0.5f * (cm - cp)) / den;
      delta = fminf(fmaxf(delta, -0.5f), 0.5f);
    }
  }
  const float sstar = (s0 + (float)nbest) + delta;
  const float ucx = fmaf(sstar, lx, umx), ucy = fmaf(sstar, ly, umy);
  ucmp[0] = ucx;
  ucmp[1] = ucy;
  /* measurement + variance */
  const int use_x = fabsf(lx) >= fabsf(ly);
  const float xi_m = idepth_at(ucx, ucy, use_x, P0x, P0y, P0z, bx, by, bz);
  const float sp = sstar + 0.5f, sm = sstar - 0.5f;
  const float xi_p = idepth_at(fmaf(sp, lx, umx), fmaf(sp, ly, umy), use_x, P0x, P0y, P0z, bx, by, bz);
  const float xi_n = idepth_at(fmaf(sm, lx, umx), fmaf(sm, ly, umy), use_x, P0x, P0y, P0z, bx, by, bz);
  const float alpha = xi_p - xi_n;
  const float g2 = fmaf(gx, gx, gy * gy);
  const float gl = fmaf(gx, lrx, gy * lry);
  const float var_geo = p->epipolar_line_var * ((g2 + 1e-6f) / fmaf(gl, gl, 1e-6f));
  const float var_photo = (2.0f * p->pixel_noise_var) / (grad2 + 1e-6f);
  const float var_m = (alpha * alpha) * (var_geo + var_photo);
  if (!(var_m <= p->meas_var_max) || !(xi_m == xi_m)) return FO_FAIL_MAX_VAR;
  const float den = v + var_m;
  *mu_io = fmaf(var_m, m, v * xi_m) / den;
  *var_io = (v * var_m) / den;
  return FO_SUCCESS;
}

void fo_idepth_update(int W, int H, int n_slots, const uint8_t* imgs, const float* poses,
                      const float* K, int cmp_slot, int N, const int32_t* ref_slot,
                      const float* u_ref, float* mu, float* var, int32_t* dropouts,
                      int32_t* alive, int32_t* status, float* u_cmp, int32_t* counters,
                      const fo_epi_params* p, int nthreads) {
  float* G = (float*)malloc(sizeof(float) * 15 * (size_t)n_slots);
  for (int s = 0; s < n_slots; ++s)
    fo_epi_geometry(K, poses + 7 * s, poses + 7 * cmp_slot, G + 15 * s);
  const size_t fsz = (size_t)W * (size_t)H;
  (void)nthreads;
#ifdef _OPENMP
  int nt = nthreads > 1 ? nthreads : 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt) if (nt > 1)
#endif
  for (int f = 0; f < N; ++f) {
    u_cmp[2 * f] = NAN;
    u_cmp[2 * f + 1] = NAN;
    if (!alive[f]) {
      status[f] = FO_SKIPPED;
      continue;
    }
    int r = ref_slot[f];
    int st;
    if (r == cmp_slot) {
      st = FO_NO_PARALLAX;
    } else {
      st = update_one(W, H, imgs + fsz * (size_t)r, imgs + fsz * (size_t)cmp_slot, G + 15 * r,
                      u_ref[2 * f], u_ref[2 * f + 1], &mu[f], &var[f], &u_cmp[2 * f], p);
    }
    if (st == FO_SUCCESS) {
      dropouts[f] = 0;
    } else if (st != FO_NO_PARALLAX) {
      u_cmp[2 * f] = NAN;
      u_cmp[2 * f + 1] = NAN;
      dropouts[f] += 1;
      if (dropouts[f] > p->max_dropouts) {
        alive[f] = 0;
        st = FO_FAIL_MAX_DROPOUTS;
      }
    }
    status[f] = st;
  }
  for (int k = 0; k < FO_NUM_COUNTERS; ++k) counters[k] = 0;
  for (int f = 0; f < N; ++f)
    if (status[f] >= 0 && status[f] < FO_NUM_COUNTERS) counters[status[f]]++;
  free(G);
}

/* ======================================================================== */
/* Feature projection into the current frame (row a10 / project_features)   */
/* ======================================================================== */

void fo_project_features(int W, int H, int n_slots, const float* poses, const float* K,
                         int cur_slot, int N, const int32_t* ref_slot, const float* u_ref,
                         const float* mu, const float* var, const int32_t* alive,
                         float* u_cur, float* mu_cur, float* var_cur, int32_t* valid) {
  float* G = (float*)malloc(sizeof(float) * 15 * (size_t)n_slots);
  for (int s = 0; s < n_slots; ++s)
    fo_epi_geometry(K, poses + 7 * s, poses + 7 * cur_slot, G + 15 * s);
  for (int f = 0; f < N; ++f) {
    valid[f] = 0;
    u_cur[2 * f] = NAN;
    u_cur[2 * f + 1] = NAN;
    mu_cur[f] = NAN;
    var_cur[f] = NAN;
    if (!alive[f]) continue;
    const float* g = G + 15 * ref_slot[f];
    float ux = u_ref[2 * f], uy = u_ref[2 * f + 1], m = mu[f];
    float px = fmaf(m, g[9], fmaf(g[0], ux, fmaf(g[1], uy, g[2])));
    float py = fmaf(m, g[10], fmaf(g[3], ux, fmaf(g[4], uy, g[5])));
    float pz = fmaf(m, g[11], fmaf(g[6], ux, fmaf(g[7], uy, g[8])));
    if (!(pz > 1e-6f)) continue;
    float x = px / pz, y = py / pz;
    if (!(x >= 0.0f && y >= 0.0f && x <= (float)(W - 1) && y <= (float)(H - 1))) continue;
    /* depth in cur = pz / mu (ref ray has unit z) => idepth_cur = mu / pz */
    float r = 1.0f / pz;
    float r2 = r * r;
    u_cur[2 * f] = x;
    u_cur[2 * f + 1] = y;
    mu_cur[f] = m * r;
    var_cur[f] = var[f] * (r2 * r2);
    valid[f] = 1;
  }
  free(G);
}

/* ======================================================================== */
/* Frame creation + detection (row f2)                                       */
/* ======================================================================== */

void fo_gradient_mag(int W, int H, const uint8_t* img, float* mag) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float g = 0.0f;
      if (x > 0 && y > 0 && x < W - 1 && y < H - 1) {
        float gx = 0.5f * ((float)img[y * W + x + 1] - (float)img[y * W + x - 1]);
        float gy = 0.5f * ((float)img[(y + 1) * W + x] - (float)img[(y - 1) * W + x]);
        g = sqrtf(fmaf(gx, gx, gy * gy));
      }
      mag[y * W + x] = g;
    }
}

void fo_pyr_down(int W, int H, const uint8_t* img, uint8_t* out) {
  int w2 = W / 2, h2 = H / 2;
  for (int y = 0; y < h2; ++y)
    for (int x = 0; x < w2; ++x) {
      int s = img[(2 * y) * W + 2 * x] + img[(2 * y) * W + 2 * x + 1] +
              img[(2 * y + 1) * W + 2 * x] + img[(2 * y + 1) * W + 2 * x + 1];
      out[y * w2 + x] = (uint8_t)((s + 2) >> 2);
    }
}

int fo_detect_features_rows(int W, int H, const float* mag, int win, int border, float min_grad_mag,
                            const uint8_t* occupied, int y_lo, int y_hi, float* det_xy, int32_t* det_ok) {
  int cx = W / win, cy = H / win, n = 0;
  for (int j = 0; j < cy; ++j)
    for (int i = 0; i < cx; ++i) {
      int c = j * cx + i;
      det_ok[c] = 0;
      det_xy[2 * c] = 0.0f;
      det_xy[2 * c + 1] = 0.0f;
      if (occupied && occupied[c]) continue;
      float best = -1.0f;
      int bx = -1, by = -1;
      for (int y = j * win; y < (j + 1) * win; ++y)
        for (int x = i * win; x < (i + 1) * win; ++x) {
          if (x < border || y < border || x >= W - border || y >= H - border || y < y_lo || y >= y_hi) continue;
          float g = mag[y * W + x];
          if (g > best) {
            best = g;
            bx = x;
            by = y;
          }
        }
      if (bx >= 0 && best >= min_grad_mag) {
        det_ok[c] = 1;
        det_xy[2 * c] = (float)bx;
        det_xy[2 * c + 1] = (float)by;
        ++n;
      }
    }
  return n;
}

int fo_detect_features(int W, int H, const float* mag, int win, int border, float min_grad_mag,
                       const uint8_t* occupied, float* det_xy, int32_t* det_ok) {
  return fo_detect_features_rows(W, H, mag, win, border, min_grad_mag, occupied, 0, H, det_xy, det_ok);
}

/* ======================================================================== */
/* Mesh -> dense inverse depth (row f1)                                      */
/* ======================================================================== */

void fo_triangle_validity(int W, int H, const float* K, int V, const float* vtx,
                          const float* idepth, int T, const int32_t* tri,
                          const fo_tri_filter_params* fp, uint8_t* valid) {
  (void)H;
  (void)V;
  const float fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  const float cos_thresh = (float)cos((double)fp->oblique_normal_thresh);
  const float len_max = fp->edge_length_thresh * (float)W;
  const float len2_max = len_max * len_max;
  for (int t = 0; t < T; ++t) {
    int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    float d[3] = {idepth[a], idepth[b], idepth[c]};
    float px[3] = {vtx[2 * a], vtx[2 * b], vtx[2 * c]};
    float py[3] = {vtx[2 * a + 1], vtx[2 * b + 1], vtx[2 * c + 1]};
    int ok = 1;
    for (int k = 0; k < 3; ++k)
      if (!(d[k] > 0.0f)) ok = 0; /* NaN or non-positive idepth: never valid */
    if (ok && fp->do_idepth)
      for (int k = 0; k < 3; ++k)
        if (d[k] < fp->min_triangle_idepth) ok = 0;
    if (ok && fp->do_edge_length)
      for (int k = 0; k < 3; ++k) {
        int k2 = (k + 1) % 3;
        float ex = px[k] - px[k2], ey = py[k] - py[k2];
        if (fmaf(ex, ex, ey * ey) > len2_max) ok = 0;
      }
    if (ok && fp->do_oblique) {
      float dmax = fmaxf(d[0], fmaxf(d[1], d[2]));
      float dmin = fminf(d[0], fminf(d[1], d[2]));
      float diff = dmax - dmin;
      if (diff > fmaxf(fp->oblique_idepth_diff_factor * dmax, fp->oblique_idepth_diff_abs)) ok = 0;
      /* back-project: P = ((u-cx)/fx, (v-cy)/fy, 1) / idepth */
      float P[3][3];
      for (int k = 0; k < 3; ++k) {
        float zk = 1.0f / d[k];
        P[k][0] = ((px[k] - cx) / fx) * zk;
        P[k][1] = ((py[k] - cy) / fy) * zk;
        P[k][2] = zk;
      }
      float e1[3], e2[3], n[3], ctr[3];
      for (int k = 0; k < 3; ++k) {
        e1[k] = P[1][k] - P[0][k];
        e2[k] = P[2][k] - P[0][k];
        ctr[k] = (P[0][k] + P[1][k]) + P[2][k];
      }
      n[0] = fmaf(e1[1], e2[2], -(e1[2] * e2[1]));
      n[1] = fmaf(e1[2], e2[0], -(e1[0] * e2[2]));
      n[2] = fmaf(e1[0], e2[1], -(e1[1] * e2[0]));
      float nn = fmaf(n[0], n[0], fmaf(n[1], n[1], n[2] * n[2]));
      float cc = fmaf(ctr[0], ctr[0], fmaf(ctr[1], ctr[1], ctr[2] * ctr[2]));
      float dot = fmaf(n[0], ctr[0], fmaf(n[1], ctr[1], n[2] * ctr[2]));
      /* |cos| < cos_thresh  <=>  dot^2 < cos_thresh^2 * nn * cc (cos_thresh >= 0) */
      if (cos_thresh > 0.0f && dot * dot < (cos_thresh * cos_thresh) * (nn * cc)) ok = 0;
    }
    valid[t] = (uint8_t)ok;
  }
}

static inline float edge_fn(float ax, float ay, float bx, float by, float px, float py) {
  return fmaf(bx - ax, py - ay, -((by - ay) * (px - ax)));
}

/* Rows [ya, yb) of the map: triangles in index order, first (smallest-index) cover wins. */
static void rasterize_band(int W, int H, const float* vtx, const float* idepth, int T, const int32_t* tri,
                           const uint8_t* valid, float* map, int ya, int yb) {
  for (int i = ya * W; i < yb * W; ++i) map[i] = NAN;
  for (int t = 0; t < T; ++t) {
    if (valid && !valid[t]) continue;
    int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    float ax = vtx[2 * a], ay = vtx[2 * a + 1];
    float bx = vtx[2 * b], by = vtx[2 * b + 1];
    float cx = vtx[2 * c], cy = vtx[2 * c + 1];
    float ymin = fminf(ay, fminf(by, cy)), ymax = fmaxf(ay, fmaxf(by, cy));
    int y0 = (int)ceilf(fmaxf(ymin, 0.0f)), y1 = (int)floorf(fminf(ymax, (float)(H - 1)));
    if (y0 < ya) y0 = ya;
    if (y1 > yb - 1) y1 = yb - 1;
    if (y1 < y0) continue;
    float area = edge_fn(ax, ay, bx, by, cx, cy);
    if (area == 0.0f || !(area == area)) continue;
    float xmin = fminf(ax, fminf(bx, cx)), xmax = fmaxf(ax, fmaxf(bx, cx));
    int x0 = (int)ceilf(fmaxf(xmin, 0.0f)), x1 = (int)floorf(fminf(xmax, (float)(W - 1)));
    float inv = 1.0f / area;
    for (int y = y0; y <= y1; ++y)
      for (int x = x0; x <= x1; ++x) {
        float* out = &map[y * W + x];
        if (*out == *out) continue; /* first (smallest-index) triangle wins */
        float px = (float)x, py = (float)y;
        float w0 = edge_fn(bx, by, cx, cy, px, py) * inv;
        float w1 = edge_fn(cx, cy, ax, ay, px, py) * inv;
        float w2 = edge_fn(ax, ay, bx, by, px, py) * inv;
        if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f)
          *out = fmaf(w0, idepth[a], fmaf(w1, idepth[b], w2 * idepth[c]));
      }
  }
}

/* nthreads > 1: the image is cut into horizontal bands, one per thread; every band walks the
 * triangles in index order restricted to its rows, so the result does not depend on nthreads. */
void fo_rasterize_idepth_mt(int W, int H, int V, const float* vtx, const float* idepth, int T,
                            const int32_t* tri, const uint8_t* valid, float* map, int nthreads) {
  (void)V;
  int nb = nthreads > 1 ? nthreads : 1;
  if (nb > H) nb = H;
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nb) if (nb > 1)
#endif
  for (int k = 0; k < nb; ++k)
    rasterize_band(W, H, vtx, idepth, T, tri, valid, map, (int)((long long)H * k / nb), (int)((long long)H * (k + 1) / nb));
}

void fo_rasterize_idepth(int W, int H, int V, const float* vtx, const float* idepth, int T,
                         const int32_t* tri, const uint8_t* valid, float* map) {
  fo_rasterize_idepth_mt(W, H, V, vtx, idepth, T, tri, valid, map, 1);
}
