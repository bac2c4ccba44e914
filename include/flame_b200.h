/*
 * flame_b200.h -- C-ABI of the B200-native FLaME hot path (libflame_b200.so).
 *
 * This is the drop-in boundary below `flame::Flame` (include/flame/flame.h): plain pointers and
 * sizes, no C++ / torch / OpenCV / Eigen types.  Every entry point names the reference interface
 * it stands in for.  The reference tree (/root/reference) is only the ROS wrapper; the functions
 * replaced here live in the external `flame` core library that the wrapper links
 * (/root/reference/CMakeLists.txt:57, src/CMakeLists.txt:11), so citations give the wrapper's call
 * site / parameter source for each one.
 *
 * A context owns a BATCH of `n_streams` independent camera streams of identical image size.  All
 * compute entry points process every stream of the batch in one launch (block-diagonal layout);
 * n_streams = 1 is the reference's single-camera use.  All work is enqueued on one CUDA stream
 * (the caller's, or one the context creates); calls return after enqueueing unless documented as
 * synchronising.  Host pointers are caller-owned and may be pageable (pinned is faster).
 *
 * Threads and devices: a context belongs to the device it was created on; every entry point makes
 * that device current for the CALLING host thread (and leaves it current), so a context may be driven
 * from any thread -- one thread at a time per context; different contexts are independent (the
 * reference runs one flame::Flame per camera thread).
 *
 * Return values: 0 = OK, negative = error (FB_E_*); fb_last_error() returns a description.
 * There is NO CPU fallback: without a CUDA device fb_create() fails.
 */
#ifndef FLAME_B200_H_
#define FLAME_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_OK 0
#define FB_E_ARG (-1)     /* bad argument / out of range */
#define FB_E_CUDA (-2)    /* CUDA runtime error (see fb_last_error) */
#define FB_E_STATE (-3)   /* call sequence error (e.g. solve before graph_set) */
#define FB_E_NOMEM (-4)   /* capacity exceeded / allocation failure */

typedef struct fb_ctx fb_ctx;

/* flame::Params::rparams {data_factor, step_x, step_q, theta}
 * (/root/reference/src/flame_nodelet.cc:256-259; defaults cfg/flame_nodelet.yaml:86-89). */
typedef struct {
  float data_factor; /* lambda, 0.15  */
  float step_x;      /* tau,    0.001 */
  float step_q;      /* sigma,  125   */
  float theta;       /* theta,  0.25  */
  float x_min;       /* box on inverse depth, 0  */
  float x_max;       /* 10 */
} fb_nltgv2_params;

/* flame::Params::{fparams,zparams,max_dropouts}
 * (/root/reference/src/flame_nodelet.cc:227,237-245; defaults cfg/flame_nodelet.yaml:69-75),
 * plus the thresholds the reference does not expose (documented in DESIGN.md). */
typedef struct {
  int win_size;            /* 5, odd, <= 15 */
  float min_grad_mag;      /* 5.0 */
  float epipolar_line_var; /* 4.0 */
  int max_dropouts;        /* 5 */
  float search_sigma;      /* 2.0 */
  float max_cost;          /* 400 (mean squared residual per patch sample) */
  float ambiguity_ratio;   /* 1.5 */
  int ambiguity_radius;    /* 2 */
  float pixel_noise_var;   /* 4.0 */
  float meas_var_max;      /* 1.0 */
  float idepth_min;        /* 0.0 */
  float idepth_max;        /* 10.0 */
  int max_search_px;       /* 64, <= 256 */
  float min_parallax;      /* 0.5 px per unit inverse depth */
} fb_epi_params;

/* Feature status codes; names follow the reference's failure counters
 * (/root/reference/src/utils.cc:124-129, msg/FlameStats.msg:14-19). */
enum {
  FB_SUCCESS = 0,
  FB_FAIL_REF_PATCH_GRADIENT = 1,
  FB_FAIL_AMBIGUOUS_MATCH = 2,
  FB_FAIL_MAX_COST = 3,
  FB_FAIL_MAX_VAR = 4,
  FB_FAIL_MAX_DROPOUTS = 5,
  FB_FAIL_OUT_OF_IMAGE = 6,
  FB_NO_PARALLAX = 7,
  FB_NUM_COUNTERS = 8,
  FB_SKIPPED = 8
};

/* Output filters of getFilteredInverseDepthMap (/root/reference/src/flame_nodelet.cc:182-206). */
typedef struct {
  int do_oblique;
  float oblique_normal_thresh;
  float oblique_idepth_diff_factor;
  float oblique_idepth_diff_abs;
  int do_edge_length;
  float edge_length_thresh;
  int do_idepth;
  float min_triangle_idepth;
} fb_tri_filter_params;

/* ------------------------------------------------------------------ lifecycle
 * Stands in for flame::Flame::Flame(width, height, K, Kinv, params)
 * (/root/reference/src/flame_nodelet.cc:523-527).  Capacities are per stream.
 * cuda_stream: a cudaStream_t to enqueue on, or NULL to create one. */
fb_ctx* fb_create(int device, int n_streams, int width, int height, int n_slots,
                  int max_features, int max_vertices, int max_edges, void* cuda_stream);
void fb_destroy(fb_ctx* ctx);
const char* fb_last_error(const fb_ctx* ctx); /* ctx may be NULL: error of the last failed fb_create */
int fb_sync(fb_ctx* ctx);                     /* cudaStreamSynchronize */
int fb_version(void);

/* Pinned host memory for callers that want fast, truly asynchronous transfers. */
void* fb_host_alloc(size_t bytes);
void fb_host_free(void* p);

int fb_set_intrinsics(fb_ctx* ctx, int stream, const float K[9]); /* row-major 3x3 pinhole */
int fb_set_epi_params(fb_ctx* ctx, const fb_epi_params* p);
void fb_default_epi_params(fb_epi_params* p);
void fb_default_nltgv2_params(fb_nltgv2_params* p);
void fb_default_tri_filter_params(fb_tri_filter_params* p);

/* ------------------------------------------------------------------ NLTGV2-L1 solver
 * Stands in for flame::optimizers::nltgv2_l1_graph_regularizer (graph types + step()), driven
 * from inside flame::Flame::update (/root/reference/src/flame_nodelet.cc:634). */

/* Topology (re)upload for one stream: V vertices at pixel positions pos_xy[2V], E canonical edges
 * edge_ij[2E] (i<j, sorted by (i,j); i = source, j = target), per-edge weights. Builds the CSR
 * incidence (ascending edge id per vertex). State is reset to zero; call fb_graph_state_set or
 * fb_graph_data_set (+ init) afterwards. */
int fb_graph_set(fb_ctx* ctx, int stream, int V, int E, const float* pos_xy,
                 const int32_t* edge_ij, const float* alpha, const float* beta);
/* Data term z[V] and data weight wt[V] (NULL = all ones). */
int fb_graph_data_set(fb_ctx* ctx, int stream, const float* z, const float* wt);
/* Warm start. x[V], w[2V] (w1,w2 interleaved), q[3E] (q1,q2,q3 interleaved).
 * NULL x => x = z; NULL w/q => zeros.  The extragradient point is reset to (x, w). */
int fb_graph_state_set(fb_ctx* ctx, int stream, const float* x, const float* w, const float* q);
/* Copies state out (synchronises). Any pointer may be NULL. xbar[3V] = (xb,w1b,w2b) interleaved. */
int fb_graph_state_get(fb_ctx* ctx, int stream, float* x, float* w, float* q, float* xbar);
/* x of every stream into x_all[n_streams * max_vertices] (stream-major); synchronises. */
int fb_graph_x_get_all(fb_ctx* ctx, float* x_all);
/* `iters` Chambolle-Pock iterations on every stream's graph (one batched launch sequence).
 * variant: 0 = auto, 1 = streaming (two kernels per iteration, any size),
 *          2 = persistent thread-block-cluster kernel (graph resident in shared memory, one
 *              cluster of <= 16 CTAs per stream, DSMEM exchange),
 *          3 = resident kernel with ONE halo exchange per iteration (cut edges held by both
 *              sides): a cluster of <= 16 CTAs per stream with DSMEM st.async hand-over when the
 *              graphs fit, else every co-resident CTA of the device with tagged 128-bit
 *              mailboxes in L2 (cooperative launch).
 *          4 = plan-free resident kernel: one cluster per stream, state in registers, exchange
 *              through L2 behind the cluster barrier; needs no per-topology tables, so it serves
 *              graphs that change every frame (fb_update).
 *          5 = tile-resident kernel planned on the device: a k-d split of the vertex positions
 *              into 16 tiles (one small kernel), then one cluster of 16 CTAs per stream with the
 *              state in registers and the exchange in (distributed) shared memory.
 * auto picks 3 when the batch fits, else 2, else 1 (5, else 4, else 1, for graphs built on the device
 * by fb_update).  All variants give bit-identical results. */
int fb_nltgv2_solve(fb_ctx* ctx, int iters, const fb_nltgv2_params* p, int variant);
/* nltgv2_total_{smoothness,data}_cost (/root/reference/src/utils.cc:131-136); synchronises. */
int fb_costs(fb_ctx* ctx, int stream, float data_factor, double* smoothness, double* data);

/* ------------------------------------------------------------------ epipolar inverse-depth update
 * Stands in for flame::stereo::inverse_depth_filter::{search,update} + InverseDepthMeasModel,
 * i.e. the `update_idepths` stage of flame::Flame::update (/root/reference/src/utils.cc:150). */

/* Upload one frame (gray uint8, `pitch` bytes per row) and its camera-in-world pose
 * (qx,qy,qz,qw,tx,ty,tz; RDF optical frame, /root/reference/README.md:168-171) into a slot. */
int fb_frame_set(fb_ctx* ctx, int stream, int slot, const uint8_t* gray, int pitch,
                 const float pose[7]);
/* Pose-only update of a slot (flame::Flame::updatePoseFramePoses,
 * /root/reference/src/flame_nodelet.cc:474). */
int fb_frame_pose_set(fb_ctx* ctx, int stream, int slot, const float pose[7]);
/* Device-resident frame pool (bench `value` leg: inputs already in HBM before timing starts):
 * fb_pool_upload stores a frame in pool entry `idx`; fb_frame_from_pool copies it device-to-device
 * into a slot. */
int fb_pool_reserve(fb_ctx* ctx, int n_entries);
int fb_pool_upload(fb_ctx* ctx, int idx, const uint8_t* gray, int pitch);
int fb_frame_from_pool(fb_ctx* ctx, int stream, int slot, int idx, const float pose[7]);

/* Feature table of one stream (device resident). NULL dropouts => 0, NULL alive => 1. */
int fb_features_set(fb_ctx* ctx, int stream, int N, const float* u_ref, const int32_t* ref_slot,
                    const float* mu, const float* var, const int32_t* dropouts,
                    const int32_t* alive);
/* Copies out (synchronises); any pointer may be NULL. Backs flame::Flame::getRawIDepths
 * (/root/reference/src/flame_nodelet.cc:721-723). */
int fb_features_get(fb_ctx* ctx, int stream, float* mu, float* var, int32_t* dropouts,
                    int32_t* alive, int32_t* status, float* u_cmp);
/* Device-side re-initialisation of every feature of every stream s with ref_slot[s] >= 0:
 * mu = mu0, var = var0, dropouts = 0, alive = 1, ref_slot = ref_slot[s].  Emulates the detector
 * handing a fresh feature set to the filter on a new poseframe without a host round trip. */
int fb_features_reinit(fb_ctx* ctx, const int32_t* ref_slot, float mu0, float var0);
/* One epipolar update of every live feature of every stream against that stream's comparison
 * slot cmp_slot[s] (cmp_slot[s] < 0 skips stream s). */
int fb_idepth_update(fb_ctx* ctx, const int32_t* cmp_slot);
/* Status histogram of the last update of `stream` (synchronises):
 * num_idepth_updates, num_fail_* (/root/reference/src/utils.cc:124-129). */
int fb_idepth_counters(fb_ctx* ctx, int stream, int32_t counters[FB_NUM_COUNTERS]);
/* Project live features into the frame held by cur_slot (project_features stage,
 * /root/reference/src/utils.cc:151). Outputs are host arrays [N]; synchronises. */
int fb_project_features(fb_ctx* ctx, int stream, int cur_slot, float* u_cur, float* mu_cur,
                        float* var_cur, int32_t* valid);

/* ------------------------------------------------------------------ data-term assembly (sync_graph)
 * vertex_feature[V]: index of the feature that backs each vertex of `stream`'s graph. */
int fb_graph_bind_features(fb_ctx* ctx, int stream, const int32_t* vertex_feature);
/* For every stream: z[v] = mu[feature(v)], wt[v] = adaptive ? 1/var : 1
 * (regularization/nltgv2/adaptive_data_weights, /root/reference/cfg/flame_nodelet.yaml:82).
 * Vertices whose feature is dead keep their previous data term with weight 0. */
int fb_graph_data_from_features(fb_ctx* ctx, int adaptive_weights);

/* ------------------------------------------------------------------ one frame of every stream
 * The per-frame hot path of flame::Flame::update (/root/reference/src/flame_nodelet.cc:634) for a
 * batch, in one call: [new poseframe: load it, re-initialise the filters] -> load the new frame ->
 * epipolar update -> data-term assembly -> `iters` NLTGV2-L1 iterations -> optional D2H of x.
 * Images come from host memory (pitch == width) or, when the image pointer arrays are NULL, from
 * the device pool. With x_out == NULL nothing is copied back and the call does not synchronise. */
typedef struct {
  int new_poseframe;
  int ref_slot, cmp_slot;
  const uint8_t* const* ref_images; /* [n_streams] host images, or NULL: use ref_pool_idx */
  const uint8_t* const* cmp_images; /* [n_streams] host images, or NULL: use cmp_pool_idx */
  const int32_t* ref_pool_idx;      /* [n_streams] */
  const int32_t* cmp_pool_idx;      /* [n_streams] */
  const float* ref_poses;           /* [n_streams*7], read when new_poseframe */
  const float* cmp_poses;           /* [n_streams*7] */
  float mu0, var0;                  /* prior of re-initialised features */
  int adaptive_weights;
  int iters, variant;
  fb_nltgv2_params rparams;
  float* x_out;                     /* host [n_streams*max_vertices] or NULL */
  int pipelined;                    /* 0: as documented above.  1: streaming mode -- frames are brought
                                       into their slots on a copy stream (overlapping the previous
                                       frame's kernels; the caller alternates cmp_slot between two
                                       slots), the solve of frame k overlaps the epipolar update of
                                       frame k+1, x_out (pinned) is filled asynchronously and the
                                       call never blocks; fb_results_wait() waits for a frame's x_out. */
  int ref_from_slot;                /* new_poseframe steps: 1 + slot whose frame BECOMES the poseframe (device
                                       copy, no second upload) -- in the reference a poseframe is the current
                                       frame flagged is_poseframe (/root/reference/src/flame_nodelet.cc:634), not
                                       a separate image.  0 = upload ref_images / ref_pool_idx as before. */
} fb_step_desc;
int fb_hotpath_step(fb_ctx* ctx, const fb_step_desc* d);
/* Waits until the x_out of the pipelined step issued `lag` calls ago (0 = the latest) has landed. */
int fb_results_wait(fb_ctx* ctx, int lag);
/* Device-side join: the context's main stream waits for everything the pipelined steps enqueued on
 * the auxiliary streams (does not block the host). */
int fb_pipeline_join(fb_ctx* ctx);

/* ------------------------------------------------------------------ flame::Flame::update and getters
 * fb_update is the whole per-frame pipeline of flame::Flame::update(time, img_id, T_world_cam, gray,
 * is_poseframe) (/root/reference/src/flame_nodelet.cc:634; src/flame_offline_tum.cc:578): frame
 * upload -> epipolar update of the feature pool -> projection into the new frame -> graph sync
 * (vertex selection by idepth_var_max_graph, Delaunay, device-side carry-over of x/w/q) -> NLTGV2-L1
 * iterations -> dense interpolation -> on poseframes: ring insertion + grid detection.
 * Returns 1 when the mesh/depth outputs were updated, 0 when not yet (first frames), <0 on error:
 * the reference's `bool update()` (/root/reference/src/flame_nodelet.cc:636-642). */
typedef struct {
  int detection_win_size;     /* features/detection/win_size, 16 (cfg/flame_nodelet.yaml:71) */
  float min_grad_mag;         /* features/detection/min_grad_mag, 5.0 */
  int detection_border;       /* px of image border without detections, 8 */
  float idepth_init;          /* prior mean of a new feature when no prediction exists, 0.5 */
  float idepth_var_init;      /* prior variance of a new feature, 0.25 */
  float idepth_var_max_graph; /* regularization/nltgv2/idepth_var_max, 0.01 */
  int adaptive_data_weights;  /* 0 */
  int init_with_prediction;   /* 1 */
  int do_nltgv2;              /* 1 */
  int iters;                  /* NLTGV2 iterations per frame, 50 (BASELINE configs) */
  fb_nltgv2_params rparams;
  int triangulator;           /* sync_graph + triangulate: 0 = on the device (per-vertex Delaunay stars,
                                 default, no host round trip), 1 = host (incremental Bowyer-Watson).
                                 Both give the same canonical mesh. */
  /* regularization/nltgv2/{rescale_data,min_height,max_height,check_sticky_obstacles} and
   * features/{do_letterbox,detection/min_error} (/root/reference/src/flame_nodelet.cc:225-231,251,260-263;
   * defaults cfg/flame_nodelet.yaml:67,70,83,90-92).  The reference only names these options; the
   * semantics below are this library's:
   *   rescale_data   z, x and w are divided by mean(z) before the iterations and multiplied back after
   *   min/max_height a feature enters the graph only when the world z of its 3-D point lies in the band
   *   do_letterbox   detection only in the middle third of the rows [H/3, 2H/3)
   *   min_error, check_sticky_obstacles: not implemented -- fb_set_update_params rejects any value
   *   other than the defaults (100, 0) with FB_E_ARG instead of ignoring it. */
  int rescale_data;           /* 0 */
  float min_height;           /* -1e14 */
  float max_height;           /* 1e14 */
  int check_sticky_obstacles; /* 0 */
  float min_error;            /* 100 */
  int do_letterbox;           /* 0 */
} fb_update_params;
void fb_default_update_params(fb_update_params* p);
int fb_set_update_params(fb_ctx* ctx, const fb_update_params* p);
int fb_update(fb_ctx* ctx, int stream, double time, int img_id, const float pose[7],
              const uint8_t* gray, int pitch, int is_poseframe);
/* getInverseDepthMesh (/root/reference/src/flame_nodelet.cc:669-676). fb_get_mesh_sizes first;
 * arrays: vtx_xy[2V], idepth[V], normals[3V], tris[3T], tri_valid[T], edges[2E]; any may be NULL.
 * filter == NULL: every triangle with positive idepths is valid. Synchronises. */
int fb_get_mesh_sizes(fb_ctx* ctx, int stream, int32_t* V, int32_t* T, int32_t* E);
int fb_get_mesh(fb_ctx* ctx, int stream, const fb_tri_filter_params* filter, float* vtx_xy,
                float* idepth, float* normals, int32_t* tris, uint8_t* tri_valid, int32_t* edges);
/* getInverseDepthMap / getFilteredInverseDepthMap (/root/reference/src/flame_nodelet.cc:682-688):
 * filter == NULL -> unfiltered. out[H*W], NaN = no depth. Synchronises. */
int fb_get_idepthmap(fb_ctx* ctx, int stream, const fb_tri_filter_params* filter, float* out);
/* A camera thread's loop in one call: for k in [k0, k1): fb_update(time = k / 30, img_id = k, poses + 7 k,
 * frames + k * frame_stride, pitch = W, is_poseframe = (k % poseframe_every == 0)) and, when it returns 1,
 * fb_get_idepthmap(filter, out_map) -- exactly what the nodelet does per image
 * (/root/reference/src/flame_nodelet.cc:634,682-683), driven from C so that a host in another language
 * (bench.py: Python, one thread per camera) does not pay interpreter time and lock contention per frame.
 * Returns the number of frames that produced a map, or a negative error. */
int fb_update_run(fb_ctx* ctx, int stream, int k0, int k1, const uint8_t* frames, size_t frame_stride,
                  const float* poses, int poseframe_every, const fb_tri_filter_params* filter, float* out_map);
/* getRawIDepths (/root/reference/src/flame_nodelet.cc:721-723): live features projected into the
 * current frame. Arrays sized max_features; *N receives the count. Synchronises. */
int fb_get_raw_idepths(fb_ctx* ctx, int stream, int32_t* N, float* xy, float* mu, float* var);
/* stats()/timings() of flame::utils::StatsTracker (/root/reference/src/flame_nodelet.cc:747-749):
 * value of one key of the last update (ms for stage names, counts otherwise); FB_E_ARG if unknown. */
int fb_get_stat(fb_ctx* ctx, int stream, const char* key, double* value);
/* updatePoseFramePoses / prunePoseFrames (/root/reference/src/flame_nodelet.cc:474-475). */
int fb_update_poseframe_poses(fb_ctx* ctx, int stream, int n, const int32_t* img_ids, const float* poses);
int fb_prune_poseframes(fb_ctx* ctx, int stream, int n, const int32_t* img_ids_to_keep);
/* Frame creation / detection building blocks, exposed for parity tests (host outputs; synchronise):
 * gradient magnitude and half-resolution pyramid level of the frame held in `slot`; grid detection
 * with an optional host occupancy mask [cells]. */
int fb_frame_gradient(fb_ctx* ctx, int stream, int slot, float* mag);
int fb_frame_pyr_down(fb_ctx* ctx, int stream, int slot, uint8_t* out);
int fb_detect(fb_ctx* ctx, int stream, int slot, int win, int border, float min_grad_mag,
              const uint8_t* occupied, float* det_xy, int32_t* det_ok, int32_t* n_det);
/* Feature pool of the update pipeline (device -> host copy of every slot; arrays sized
 * max_features): for parity tests against the oracle-side mirror. Synchronises. */
int fb_get_feature_pool(fb_ctx* ctx, int stream, float* u_ref, int32_t* ref_slot, float* mu,
                        float* var, int32_t* dropouts, int32_t* alive);

/* ------------------------------------------------------------------ triangulation (host, no GPU needed)
 * Stands in for the `triangulate` stage (/root/reference/src/utils.cc:154; the external core wraps
 * Shewchuk's Triangle).  Exact-predicate incremental Delaunay of n pixel positions pts_xy[2n]
 * (snapped to 1/64 px).  tris: capacity 3*2n ints, edges: capacity 2*3n ints (canonical: i<j,
 * sorted by (i,j) -- directly usable by fb_graph_set).  Duplicate points are left unreferenced.
 * Returns FB_OK, or FB_E_ARG when the points are degenerate (n < 3 or all collinear). */
int fb_delaunay(int n, const float* pts_xy, int32_t* tris, int32_t* n_tris, int32_t* edges,
                int32_t* n_edges);

/* The same triangulation computed on the device by the kernels fb_update uses (one warp per vertex
 * computes that vertex's Delaunay star with exact predicates; csrc/delaunay_star.h).  Same output
 * contract and bit-identical output as fb_delaunay; n <= max_features and max_vertices of the
 * context.  FB_E_STATE when a star exceeds 32 neighbours.  Leaves stream `stream` without a graph. */
int fb_delaunay_device(fb_ctx* ctx, int stream, int n, const float* pts_xy, int32_t* tris, int32_t* n_tris,
                       int32_t* edges, int32_t* n_edges);

/* ------------------------------------------------------------------ mesh -> dense inverse depth
 * Stands in for the `interpolate` stage + flame::Flame::getInverseDepthMap /
 * getFilteredInverseDepthMap (/root/reference/src/flame_nodelet.cc:682-688). */
int fb_mesh_set(fb_ctx* ctx, int stream, int T, const int32_t* tri /*[3T] vertex ids*/);
/* Rasterise the current x of `stream` over its mesh. filter = NULL: no triangle filtering.
 * idepthmap [H*W] host (NaN = no depth), tri_valid [T] host (may be NULL); synchronises. */
int fb_interpolate(fb_ctx* ctx, int stream, const fb_tri_filter_params* filter, float* idepthmap,
                   uint8_t* tri_valid);

/* ------------------------------------------------------------------ profiling (CUDA events)
 * Sections: 0 = nltgv2 solve, 1 = idepth update, 2 = frame upload, 3 = data assembly,
 * 4 = interpolate. Times are accumulated between fb_profile_reset calls. */
enum { FB_PROF_SOLVE = 0, FB_PROF_IDEPTH = 1, FB_PROF_UPLOAD = 2, FB_PROF_ASSEMBLY = 3,
       FB_PROF_INTERP = 4, FB_PROF_NUM = 5 };
int fb_profile_enable(fb_ctx* ctx, int enable); /* 0 = off, 1 = all sections, 2 + section = that section only */
int fb_profile_reset(fb_ctx* ctx);
/* Synchronises; total_ms = sum of device time of the section, launches = kernels launched. */
int fb_profile_get(fb_ctx* ctx, int section, float* total_ms, int64_t* calls, int64_t* launches);
/* Total kernels launched by this context since creation. */
int64_t fb_launch_count(const fb_ctx* ctx);
/* Host-only self-check of the variant-3 partitioner (no device, no context): cuts the graph into
 * `parts` CTAs' worth of tables (cluster != 0: capacities of the cluster transport, parts <= 16;
 * else of the L2 transport) and verifies the invariants the kernel relies on (every vertex owned
 * once, register rows holding each vertex's out-edges in ascending edge id, every slot written
 * exactly once in CSR order, halo vertices published by their
 * and pushed by their owners, every edge written back exactly once).  0 = OK, 1 = does not fit this part count, other
 * > 0 = violated invariant (fb_last_error(NULL) says which), < 0 = bad argument.
 * stats[12] (optional) = {max own vertices, max generic edges per part, max halo, cut edges (held by
 * both sides), max slots, shared-memory bytes, boundary vertices, out-edges beyond the register rows, ideal / estimated
 * load / estimated store shared-memory wavefronts of the register rows' target-side accesses, 0}. */
int fb_grid_plan_verify(int V, int E, const float* pos, const int32_t* ij, int parts, int cluster,
                        int32_t* stats);
/* Which solver variant the last fb_nltgv2_solve used (1, 2 or 3). */
int fb_last_solver_variant(const fb_ctx* ctx);
/* CTAs per stream of the last variant-2 (cluster size) or variant-3 (parts per stream) launch. */
int fb_last_cluster_size(const fb_ctx* ctx);
/* Halo transport of the last variant-3 launch: 1 = thread-block cluster (DSMEM st.async +
 * mbarrier), 2 = tagged 128-bit mailboxes in L2 (cooperative launch); 0 = none yet. */
int fb_last_solver_transport(const fb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FLAME_B200_H_ */
