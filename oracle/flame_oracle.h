/*
 * flame_oracle.h -- CPU ORACLE for the FLaME hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the un-vendored,
 * un-pinned third-party library robustrobotics/flame (GPLv3 core; pulled by
 * `find_package(flame REQUIRED)` with no version, /root/reference/CMakeLists.txt:57,
 * cloned at default-branch HEAD, /root/reference/README.md:73).  It is absent from
 * /root/reference and the reference ships no tests, fixtures or golden vectors
 * for it.  This file therefore restates the published
 * algorithm (Greene & Roy, ICCV'17: NLTGV2-L1 energy + Chambolle-Pock; LSD-SLAM
 * style epipolar line stereo + Gaussian inverse-depth filter) anchored on the
 * reference's call sites, parameter names and defaults:
 *   rparams.{data_factor,step_x,step_q,theta}   /root/reference/src/flame_nodelet.cc:256-259
 *   defaults 0.15 / 0.001 / 125 / 0.25           /root/reference/cfg/flame_nodelet.yaml:86-89
 *   fparams.{min_grad_mag,win_size}, zparams.{win_size,epipolar_line_var},
 *   max_dropouts                                  /root/reference/src/flame_nodelet.cc:227-245
 *   failure counters / status taxonomy            /root/reference/src/utils.cc:124-129
 *   cost observables nltgv2_*_cost                /root/reference/src/utils.cc:131-136
 *   output filters (oblique / long edge / idepth) /root/reference/src/flame_nodelet.cc:182-206
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (flame_ros_b200/) never
 * links, imports or calls it.
 *
 * Float discipline: fp32 everywhere, built with -ffp-contract=off; fused
 * multiply-adds appear only where fmaf() is written out.  The CUDA kernels are
 * built with --fmad=false and use the same expressions in the same order, so
 * integer/index outputs are bit-exact and float outputs are bit-exact wherever
 * the kernel accumulates in the same (CSR, ascending edge id) order.
 */
#ifndef FLAME_ORACLE_H_
#define FLAME_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- NLTGV2-L1 */

typedef struct {
  float data_factor; /* lambda  (rparams.data_factor, default 0.15)  */
  float step_x;      /* tau     (rparams.step_x,      default 0.001) */
  float step_q;      /* sigma   (rparams.step_q,      default 125)   */
  float theta;       /* theta   (rparams.theta,       default 0.25)  */
  float x_min;       /* box on inverse depth (OUR CHOICE, default 0) */
  float x_max;       /* (OUR CHOICE, default 10)                     */
} fo_nltgv2_params;

/*
 * Run `iters` Chambolle-Pock iterations of NLTGV2-L1 on a graph.
 *   pos      [2V]  vertex pixel positions (x,y interleaved)
 *   edge_ij  [2E]  canonical edges (i<j, sorted by (i,j)); i = source, j = target
 *   alpha,beta [E] per-edge weights
 *   z, wt    [V]   data term and data weight
 * State (in/out): x,w1,w2,xb,w1b,w2b [V]; q1,q2,q3 [E].
 * Per-vertex accumulation of K^T q runs over incident edges in ascending edge id.
 * nthreads<=1: serial; otherwise OpenMP over edges / vertices (same results).
 */
void fo_nltgv2_solve(int V, int E, const float* pos, const int32_t* edge_ij,
                     const float* alpha, const float* beta, const float* z,
                     const float* wt, float* x, float* w1, float* w2, float* xb,
                     float* w1b, float* w2b, float* q1, float* q2, float* q3,
                     const fo_nltgv2_params* p, int iters, int nthreads);

/* smoothness = sum_e |k1|+|k2|+|k3| at (x,w); data = sum_v lambda*wt*|x-z|.
 * Per-term arithmetic in fp32, accumulated in fp64. */
void fo_nltgv2_costs(int V, int E, const float* pos, const int32_t* edge_ij,
                     const float* alpha, const float* beta, const float* z,
                     const float* wt, const float* x, const float* w1,
                     const float* w2, float data_factor, double* smoothness,
                     double* data);

/* ------------------------------------------------- epipolar idepth update  */

enum {
  FO_SUCCESS = 0,
  FO_FAIL_REF_PATCH_GRADIENT = 1,
  FO_FAIL_AMBIGUOUS_MATCH = 2,
  FO_FAIL_MAX_COST = 3,
  FO_FAIL_MAX_VAR = 4,
  FO_FAIL_MAX_DROPOUTS = 5,
  FO_FAIL_OUT_OF_IMAGE = 6,
  FO_NO_PARALLAX = 7, /* baseline too small to search: feature left untouched */
  FO_NUM_COUNTERS = 8,
  FO_SKIPPED = 8 /* dead feature slot: untouched, not counted */
};

typedef struct {
  int win_size;            /* fparams.win_size / zparams.win_size, default 5 (odd, <= 15) */
  float min_grad_mag;      /* fparams.min_grad_mag, default 5.0                    */
  float epipolar_line_var; /* zparams.epipolar_line_var, default 4.0               */
  int max_dropouts;        /* params.max_dropouts, default 5                       */
  float search_sigma;      /* OUR CHOICE: search mu +- k*sigma, default 2          */
  float max_cost;          /* OUR CHOICE: best mean squared residual above this fails (default 400) */
  float ambiguity_ratio;   /* OUR CHOICE: second_best < ratio*max(best,floor) fails (default 1.5) */
  int ambiguity_radius;    /* OUR CHOICE: |n-n*| <= radius excluded from 2nd best (default 2) */
  float pixel_noise_var;   /* OUR CHOICE: sigma_I^2 in photometric variance (default 4.0) */
  float meas_var_max;      /* OUR CHOICE: measurement variance above this fails (default 1.0) */
  float idepth_min;        /* OUR CHOICE: lower clamp of search interval (>=0), default 0.0 */
  float idepth_max;        /* OUR CHOICE: upper clamp of search interval, default 10 */
  int max_search_px;       /* OUR CHOICE: cap on candidates along the segment (default 64, <= 256) */
  float min_parallax;      /* OUR CHOICE: px of image motion per unit idepth below which the
                              feature is skipped with FO_NO_PARALLAX (default 0.5) */
} fo_epi_params;

/*
 * One epipolar inverse-depth update of N features of one stream.
 *   imgs          [n_slots][H][W] uint8 frames (tightly packed)
 *   poses         [n_slots][7]   camera-in-world (qx,qy,qz,qw, tx,ty,tz), RDF optical frame
 *   K             [9] row-major pinhole intrinsics (Kinv is formed analytically)
 *   cmp_slot      slot index of the comparison (new) frame
 *   ref_slot [N]  slot index of each feature's poseframe
 *   u_ref   [2N]  feature pixel in its poseframe
 *   mu,var   [N]  in/out inverse-depth mean / variance
 *   dropouts [N]  in/out consecutive failure count
 *   alive    [N]  in/out 1 = live, 0 = dead (dead slots are skipped)
 *   status   [N]  out   FO_* code (FO_SKIPPED for dead slots)
 *   u_cmp   [2N]  out   matched pixel in cmp (NaN unless a match was found)
 *   counters [8]  out   histogram of status codes 0..7 over live features
 */
void fo_idepth_update(int W, int H, int n_slots, const uint8_t* imgs,
                      const float* poses, const float* K, int cmp_slot, int N,
                      const int32_t* ref_slot, const float* u_ref, float* mu,
                      float* var, int32_t* dropouts, int32_t* alive,
                      int32_t* status, float* u_cmp, int32_t* counters,
                      const fo_epi_params* p, int nthreads);

/* Relative geometry used by fo_idepth_update, exposed for tests:
 * G[0..8] = A = K*R*Kinv (row-major), G[9..11] = b = K*t, with (R,t) = T_cmp<-ref,
 * G[12..14] = e = K * (-R^T t): homogeneous image of cmp's centre in ref. */
void fo_epi_geometry(const float* K, const float* pose_ref, const float* pose_cmp,
                     float* G /*[15]*/);

/* ------------------------------------- feature projection (data assembly)  */
/*
 * Project features from their poseframe into the current frame
 * (rows a10 / project_features): for each live feature
 *   p = A*(u,v,1) + mu*b ;  u_cur = (p.x/p.z, p.y/p.z)
 *   mu_cur = mu / pz_n  with pz_n = p.z (K has last row 0 0 1 so p.z is the depth ratio)
 *   var_cur = var / pz_n^4 ... see flame_oracle.c for the exact expression order.
 * Out-of-image or behind-camera features get valid=0.
 */
void fo_project_features(int W, int H, int n_slots, const float* poses,
                         const float* K, int cur_slot, int N,
                         const int32_t* ref_slot, const float* u_ref,
                         const float* mu, const float* var, const int32_t* alive,
                         float* u_cur, float* mu_cur, float* var_cur,
                         int32_t* valid);

/* ------------------------------------- frame creation + feature detection  */
/*
 * Gradient magnitude image used by the detector (row f2):
 *   gx = (I(x+1,y) - I(x-1,y))/2, gy likewise (0 on the 1-px border), mag = sqrt(gx^2+gy^2).
 */
void fo_gradient_mag(int W, int H, const uint8_t* img, float* mag);

/*
 * Half-resolution pyramid level: out(x,y) = (I(2x,2y)+I(2x+1,2y)+I(2x,2y+1)+I(2x+1,2y+1)+2)>>2.
 */
void fo_pyr_down(int W, int H, const uint8_t* img, uint8_t* out /*[H/2][W/2]*/);

/*
 * Grid detector (features/detection/{win_size,min_grad_mag}, /root/reference/cfg/flame_nodelet.yaml:68-71):
 * the image is tiled into win x win cells (partial cells at the right/bottom edges are dropped);
 * a cell whose `occupied` flag is set is skipped; otherwise its pixel with the largest gradient
 * magnitude (ties: smallest y, then smallest x; `border` px of the image edge excluded) is a
 * detection when that magnitude >= min_grad_mag.
 *   occupied [cells_y*cells_x] in   1 = a live feature already projects into the cell
 *   det_xy   [2*cells]         out  detection pixel per cell (valid where det_ok)
 *   det_ok   [cells]           out
 * returns the number of detections.
 */
int fo_detect_features(int W, int H, const float* mag, int win, int border,
                       float min_grad_mag, const uint8_t* occupied,
                       float* det_xy, int32_t* det_ok);
/* The same restricted to the rows [y_lo, y_hi) (features/do_letterbox). */
int fo_detect_features_rows(int W, int H, const float* mag, int win, int border,
                            float min_grad_mag, const uint8_t* occupied, int y_lo, int y_hi,
                            float* det_xy, int32_t* det_ok);

/* ------------------------------------- mesh -> dense inverse-depth map     */
/*
 * Triangle validity filters (output/filter_* params, /root/reference/cfg/flame_nodelet.yaml:31-46)
 * and barycentric rasterisation (row f1 `interpolate`).
 *   vtx [2V] pixel positions, idepth [V], tri [3T] vertex ids
 * Filters (each enabled by its flag):
 *   oblique : normal of the back-projected triangle (Kinv rays / idepth) makes an angle with the
 *             viewing ray of the centroid whose |cos| < cos(oblique_normal_thresh)  -> invalid, or
 *             (max-min idepth) > max(oblique_idepth_diff_factor * max idepth... see .c
 *   long edge: any edge longer than edge_length_thresh * W px -> invalid
 *   idepth  : any vertex idepth < min_triangle_idepth -> invalid
 * Rasterisation: pixel centres (integer coords) inside or on a valid triangle get the
 * barycentric interpolation of vertex idepths; ties between triangles sharing an edge are
 * resolved to the smallest triangle index; uncovered pixels are NaN.
 */
typedef struct {
  int do_oblique;
  float oblique_normal_thresh;
  float oblique_idepth_diff_factor;
  float oblique_idepth_diff_abs;
  int do_edge_length;
  float edge_length_thresh;
  int do_idepth;
  float min_triangle_idepth;
} fo_tri_filter_params;

void fo_triangle_validity(int W, int H, const float* K, int V, const float* vtx,
                          const float* idepth, int T, const int32_t* tri,
                          const fo_tri_filter_params* fp, uint8_t* valid /*[T]*/);

void fo_rasterize_idepth(int W, int H, int V, const float* vtx, const float* idepth,
                         int T, const int32_t* tri, const uint8_t* valid /*NULL = all*/,
                         float* idepthmap /*[H][W]*/);

/* Same result with the image cut into `nthreads` horizontal bands (OpenMP). */
void fo_rasterize_idepth_mt(int W, int H, int V, const float* vtx, const float* idepth, int T,
                            const int32_t* tri, const uint8_t* valid, float* idepthmap, int nthreads);

/* ------------------------------------- triangulation (flame_pipeline.c)    */
/*
 * Delaunay triangulation of n pixel positions (snapped to a 1/64 px lattice, exact predicates):
 * the `triangulate` stage (/root/reference/src/utils.cc:154).  Canonical output: co-circular point
 * sets fan out from their smallest index, identical points keep the smallest index, triangles
 * (v0 smallest, counter-clockwise in stored coordinates) sorted by (v0, v1), edges (i<j) sorted.
 * tris capacity 3*2n, edges capacity 2*3n.  Returns 0, -1 when degenerate.
 */
int fo_delaunay(int n, const float* pts, int32_t* tris, int32_t* n_tris, int32_t* edges,
                int32_t* n_edges);

/* ------------------------------------- whole per-frame pipeline            */
/* flame::Flame::update (/root/reference/src/flame_nodelet.cc:634) restated on the CPU. */
typedef struct {
  int detection_win_size;     /* features/detection/win_size, 16 */
  float min_grad_mag;         /* features/detection/min_grad_mag, 5.0 */
  int detection_border;       /* OUR CHOICE: 8 px */
  float idepth_init;          /* OUR CHOICE: prior mean of a new feature, 0.5 */
  float idepth_var_init;      /* OUR CHOICE: prior variance, 0.25 */
  float idepth_var_max_graph; /* regularization/nltgv2/idepth_var_max, 0.01 */
  int adaptive_data_weights;
  int init_with_prediction;
  int do_nltgv2;
  int iters;                  /* OUR CHOICE: iterations per frame, 50 */
  fo_nltgv2_params rparams;
  /* regularization/nltgv2/{rescale_data,min_height,max_height,check_sticky_obstacles},
   * features/{do_letterbox,detection/min_error} (/root/reference/src/flame_nodelet.cc:225-231,251,260-263).
   * Semantics are OUR CHOICE (the reference only names them):
   *   rescale_data  the data term, x and w are divided by mean(z) before the iterations and
   *                 multiplied back after them ("Rescale data to have mean 1")
   *   min/max_height  a feature enters the graph only when the world z of its 3-D point (current
   *                 camera pose, depth 1/idepth) lies in [min_height, max_height]
   *   do_letterbox  detection only in the middle third of the rows [H/3, 2H/3)
   *   min_error, check_sticky_obstacles  not restated: only their defaults (100, 0) are accepted */
  int rescale_data;
  float min_height, max_height;
  int check_sticky_obstacles;
  float min_error;
  int do_letterbox;
} fo_update_params;
void fo_default_update_params(fo_update_params* p);

enum { FO_STAGE_UPDATE = 0, FO_STAGE_FRAME, FO_STAGE_IDEPTH, FO_STAGE_PROJECT, FO_STAGE_SYNC,
       FO_STAGE_TRIANGULATE, FO_STAGE_SOLVE, FO_STAGE_INTERP, FO_STAGE_DETECT, FO_STAGE_NUM };

typedef struct fo_pipeline fo_pipeline;
fo_pipeline* fo_pipeline_create(int W, int H, const float* K /*[9]*/, int n_slots, int max_features,
                                int max_vertices, const fo_update_params* up, const fo_epi_params* ep,
                                int nthreads);
void fo_pipeline_destroy(fo_pipeline* P);
/* One frame; returns 1 when the mesh / dense map were updated. */
int fo_pipeline_update(fo_pipeline* P, int img_id, const float* pose /*[7]*/, const uint8_t* gray,
                       int is_poseframe);
void fo_pipeline_sizes(const fo_pipeline* P, int32_t* V, int32_t* T, int32_t* E);
void fo_pipeline_mesh(const fo_pipeline* P, float* vtx, float* idepth, int32_t* tris, int32_t* edges,
                      int32_t* vert_feat);
void fo_pipeline_idepthmap(const fo_pipeline* P, const fo_tri_filter_params* filter, float* out);
void fo_pipeline_features(const fo_pipeline* P, float* u_ref, int32_t* ref_slot, float* mu, float* var,
                          int32_t* dropouts, int32_t* alive, int32_t* valid);
void fo_pipeline_stage_ms(const fo_pipeline* P, double* ms /*[FO_STAGE_NUM]*/);

#ifdef __cplusplus
}
#endif
#endif /* FLAME_ORACLE_H_ */
