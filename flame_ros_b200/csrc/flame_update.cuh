// flame_update.cuh -- the per-frame orchestrator behind flame::Flame::update.
//
// Stands in for flame::Flame::update(time, img_id, T_world_cam, gray, is_poseframe)
// (/root/reference/src/flame_nodelet.cc:634, src/flame_offline_tum.cc:578) and the stages whose
// timing keys the wrapper exports (/root/reference/src/utils.cc:143-156):
//   frame_creation -> update_idepths -> project_features -> sync_graph (+ triangulate) -> NLTGV2-L1
//   iterations -> interpolate -> [poseframe: keyframe + detection].
// Host C++ sequences the stages; every per-feature / per-vertex / per-pixel computation is a CUDA
// kernel.  The only per-frame host work is the graph bookkeeping + Delaunay triangulation
// (delaunay.h), fed by one small D2H of the projected feature table.
//
// State carried frame to frame on the device: the feature pool (fixed maxF slots, `alive` flags),
// the poseframe ring (slots 0..n_slots-2; the last slot holds the current frame), the graph with
// its primal/dual state, which k_remap_state re-threads through every topology change so (x, w, q)
// never visit the host.
#pragma once

#include <unordered_map>

#include "common.cuh"
#include "delaunay.h"
#include "delaunay_gpu.cuh"
#include "epipolar.cuh"
#include "frontend.cuh"
#include "nltgv2.cuh"
#include "raster.cuh"

struct UpdateStream {
  bool have_poseframe = false;
  int pf_next = 0;                 // next ring slot to (over)write
  std::vector<int> pf_img_id;      // img_id held by each ring slot, -1 = empty
  int64_t frames = 0;
  // graph bookkeeping (host): feature index backing every vertex, canonical edges
  std::vector<int32_t> vert_feat;
  std::vector<int32_t> edges;      // 2E
  std::vector<int32_t> tris;       // 3T
  bool have_graph = false;
  bool dev_graph = false;          // the current graph was built on the device (delaunay_gpu.cuh)
  int64_t builds = 0;              // device graph builds so far (parity of the f2v tables)
  // The steady-state frame (upload .. readback) captured as a CUDA graph, one per f2v parity: a frame
  // is ~24 stream operations, and with one flame::Flame per camera and a thread per camera their
  // launches serialise on the driver's context lock; replayed as a graph a frame is ONE launch.
  cudaGraphExec_t frame_graph[2] = {nullptr, nullptr};
  int64_t graph_launches[2] = {0, 0};
  // the display filter of the last getFilteredInverseDepthMap call: once known, every update renders that map too
  // (one claim + one shading pass for both maps instead of a second rasterisation per getter call)
  bool spec_on = false;
  fb_tri_filter_params spec_filter{};
  uint64_t spec_epoch = ~0ull;     // c->mut_epoch right after the update that rendered idmap_f
  // stats of the last update (names follow msg/FlameStats.msg)
  std::unordered_map<std::string, double> stats;
};

struct UpdateState {
  std::vector<UpdateStream> st;
  std::vector<fbdel::Triangulator> tri;  // one per stream: scratch capacity survives from frame to frame
  // device scratch (per stream base s*maxF unless noted)
  float2* f_ucur = nullptr;
  float* f_mucur = nullptr;
  float* f_varcur = nullptr;
  int32_t* f_valid = nullptr;
  uint8_t* occupied = nullptr;   // [maxCells]
  float2* det_xy = nullptr;      // [maxCells]
  int32_t* det_ok = nullptr;     // [maxCells]
  int32_t* det_rank = nullptr;   // [maxCells]
  int32_t* free_flag = nullptr;  // [maxF]
  int32_t* free_rank = nullptr;  // [maxF]
  int32_t* free_list = nullptr;  // [maxF]
  int32_t* det_list = nullptr;   // [maxCells]
  int32_t* counts = nullptr;     // [2] n_det, n_free
  // old graph state stash for the remap
  float* o_x = nullptr;
  float* o_w1 = nullptr;
  float* o_w2 = nullptr;
  float4* o_vbar = nullptr;
  float4* o_q4 = nullptr;
  int32_t* map_v = nullptr;      // [maxV] new vertex -> old vertex or -1
  int32_t* map_e = nullptr;      // [maxE] new edge -> old edge or -1
  int maxCells = 0;
  fb_update_params up;
  std::vector<int32_t> cmp_scratch;
  DelGpu del;                    // device-side sync_graph + triangulate (delaunay_gpu.cuh)
  int32_t* misc = nullptr;       // [S*4] device: covered pixels, live projected features, 0, 0
  int32_t* h_read = nullptr;     // pinned [S*(DSG_META+4)]: per-frame readback of meta + misc
  // side branch of the graph build: the old state is set aside and the solver's tiles are cut while
  // the star kernel runs (fork after k_ds_prepare, join before k_ds_scan)
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint8_t* h_img = nullptr;      // pinned [S][H*W]: staging of the frame the graph uploads
  float* h_geo = nullptr;        // pinned [S][n_slots*7 + 1]: poses + comparison slot the graph uploads
  bool use_graph = true;         // FB_UPDATE_GRAPH=0 disables
};

// ------------------------------------------------------------------------------------ kernels
// alive features that project outside the current frame leave the pool
__global__ void __launch_bounds__(256)
k_kill_invalid(int N, int32_t* __restrict__ alive, const int32_t* __restrict__ valid,
               int32_t* __restrict__ n_valid = nullptr) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const bool v = f < N && valid[f];
  if (f < N && alive[f] && !v) alive[f] = 0;
  if (n_valid) {  // `num_feats` stat (/root/reference/src/utils.cc:117)
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_valid, __popc(m));
  }
}

__global__ void __launch_bounds__(256)
k_kill_ref_slot(int N, int slot, int32_t* __restrict__ alive, const int32_t* __restrict__ ref_slot) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < N && alive[f] && ref_slot[f] == slot) alive[f] = 0;
}

// project_features + kill of the features that left the frame + live count, one launch
__global__ void __launch_bounds__(256)
k_project_kill(const float* __restrict__ geo, int n_slots, int s, int N, int W, int H,
               const float2* __restrict__ u_ref, const int32_t* __restrict__ ref_slot,
               const float* __restrict__ mu, const float* __restrict__ var, int32_t* __restrict__ alive,
               float2* u_cur, float* mu_cur, float* var_cur, int32_t* valid, int32_t* __restrict__ n_valid) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  bool v = false;
  if (f < N) {
    const float qnan = __int_as_float(0x7fc00000);
    float2 uc = make_float2(qnan, qnan);
    float mc = qnan, vc = qnan;
    if (alive[f]) {
      const float* g = geo + ((size_t)s * n_slots + ref_slot[f]) * FB_GEO_STRIDE;
      const float ux = u_ref[f].x, uy = u_ref[f].y, m = mu[f];
      const float px = fmaf(m, g[9], fmaf(g[0], ux, fmaf(g[1], uy, g[2])));
      const float py = fmaf(m, g[10], fmaf(g[3], ux, fmaf(g[4], uy, g[5])));
      const float pz = fmaf(m, g[11], fmaf(g[6], ux, fmaf(g[7], uy, g[8])));
      if (pz > 1e-6f) {
        const float x = px / pz, y = py / pz;
        if (x >= 0.0f && y >= 0.0f && x <= (float)(W - 1) && y <= (float)(H - 1)) {
          const float r = 1.0f / pz;
          const float r2 = r * r;
          uc = make_float2(x, y);
          mc = m * r;
          vc = var[f] * (r2 * r2);
          v = true;
        }
      }
      if (!v) alive[f] = 0;
    }
    u_cur[f] = uc; mu_cur[f] = mc; var_cur[f] = vc;
    valid[f] = v ? 1 : 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_valid, __popc(m));
}

// z = projected idepth of the backing feature, wt = 1 or 1/var (adaptive_data_weights)
__global__ void __launch_bounds__(256)
k_data_from_projection(int V, float* __restrict__ z, float* __restrict__ wt,
                       const int32_t* __restrict__ vfeat, const float* __restrict__ mu_cur,
                       const float* __restrict__ var_cur, int adaptive) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int f = vfeat[v];
  z[v] = mu_cur[f];
  wt[v] = adaptive ? (1.0f / var_cur[f]) : 1.0f;
}

// Re-thread the solver state through a topology change: persisting vertices / edges keep
// (x, w, xbar) / q; new vertices start at the prediction (dense idepthmap at the vertex, when valid
// and init_with_prediction) or at their data term, with w = 0; new edges start at q = 0.
__global__ void __launch_bounds__(256)
k_remap_state(int V, int E, const int32_t* __restrict__ map_v, const int32_t* __restrict__ map_e,
              const float* __restrict__ o_x, const float* __restrict__ o_w1,
              const float* __restrict__ o_w2, const float4* __restrict__ o_vbar,
              const float4* __restrict__ o_q4, const float* __restrict__ z,
              const float2* __restrict__ vpos, const float* __restrict__ idmap, int W, int H,
              int use_prediction, float* x, float* w1, float* w2, float4* vbar, float4* q4) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < V) {
    const int o = map_v[t];
    if (o >= 0) {
      x[t] = o_x[o];
      w1[t] = o_w1[o];
      w2[t] = o_w2[o];
      vbar[t] = o_vbar[o];
    } else {
      float x0 = z[t];
      if (use_prediction && idmap) {
        const int px = (int)rintf(vpos[t].x), py = (int)rintf(vpos[t].y);
        if (px >= 0 && py >= 0 && px < W && py < H) {
          const float p = idmap[py * W + px];
          if (p == p && p > 0.0f) x0 = p;
        }
      }
      x[t] = x0;
      w1[t] = 0.0f;
      w2[t] = 0.0f;
      vbar[t] = make_float4(x0, 0.0f, 0.0f, 0.0f);
    }
  }
  if (t < E) {
    const int o = map_e[t];
    q4[t] = (o >= 0) ? o_q4[o] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// free_list[rank] = slot for every free slot, det_list[rank] = cell for every detection
__global__ void __launch_bounds__(256)
k_scatter_ranked(int n, const int32_t* __restrict__ flag, const int32_t* __restrict__ rank,
                 int32_t* __restrict__ list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) list[rank[i]] = i;
}

__global__ void __launch_bounds__(256)
k_free_flags(int N, const int32_t* __restrict__ alive, int32_t* __restrict__ free_flag) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < N) free_flag[f] = alive[f] ? 0 : 1;
}

// k-th detection (cell order) -> k-th free slot (ascending slot index)
__global__ void __launch_bounds__(256)
k_spawn_features(const int32_t* __restrict__ counts, const int32_t* __restrict__ det_list,
                 const int32_t* __restrict__ free_list, const float2* __restrict__ det_xy,
                 const float* __restrict__ idmap, int W, int H, int ref, float mu0, float var0,
                 int use_prediction, float2* u_ref, int32_t* ref_slot, float* mu, float* var,
                 int32_t* dropouts, int32_t* alive, int32_t* status, int have_host = 1,
                 const int32_t* __restrict__ have_dev = nullptr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(counts[0], counts[1]);
  if (k >= n) return;
  // a dense prediction exists when an earlier frame produced one (host flag) or this frame did
  if (!(have_host || (have_dev && *have_dev > 0))) idmap = nullptr;
  const int f = free_list[k];
  const float2 p = det_xy[det_list[k]];
  float m = mu0;
  if (use_prediction && idmap) {
    const float q = idmap[(int)p.y * W + (int)p.x];
    if (q == q && q > 0.0f) m = q;
  }
  u_ref[f] = p;
  ref_slot[f] = ref;
  mu[f] = m;
  var[f] = var0;
  dropouts[f] = 0;
  alive[f] = 1;
  status[f] = FB_SKIPPED;
}

// ------------------------------------------------------------------------------------ host side
static int update_alloc(fb_ctx* c) {
  if (c->upd) return FB_OK;
  UpdateState* U = new UpdateState();
  c->upd = U;
  U->st.resize(c->S);
  for (auto& s : U->st) s.pf_img_id.assign(c->n_slots - 1, -1);
  const size_t S = c->S, nf = S * c->maxF, nv = S * c->maxV, ne = S * c->maxE;
  U->maxCells = (c->W / 4) * (c->H / 4);  // detection win >= 4
  bool ok = true;
  auto A = [&](cudaError_t r) { ok = ok && r == cudaSuccess; };
  A(dalloc(&U->f_ucur, nf)); A(dalloc(&U->f_mucur, nf)); A(dalloc(&U->f_varcur, nf)); A(dalloc(&U->f_valid, nf));
  A(dalloc(&U->occupied, U->maxCells)); A(dalloc(&U->det_xy, U->maxCells)); A(dalloc(&U->det_ok, U->maxCells));
  A(dalloc(&U->det_rank, U->maxCells)); A(dalloc(&U->det_list, U->maxCells));
  A(dalloc(&U->free_flag, c->maxF)); A(dalloc(&U->free_rank, c->maxF)); A(dalloc(&U->free_list, c->maxF));
  A(dalloc(&U->counts, 2));
  A(dalloc(&U->o_x, c->maxV)); A(dalloc(&U->o_w1, c->maxV)); A(dalloc(&U->o_w2, c->maxV));
  A(dalloc(&U->o_vbar, c->maxV)); A(dalloc(&U->o_q4, c->maxE));
  A(dalloc(&U->map_v, c->maxV)); A(dalloc(&U->map_e, c->maxE));
  DelGpu& D = U->del;
  A(dalloc(&D.vxy, nv)); A(dalloc(&D.sxy, nv)); A(dalloc(&D.sid, nv)); A(dalloc(&D.vorder, nv));
  A(dalloc(&D.cell_start, S * (DSG_MAXCELLS + 1))); A(dalloc(&D.star, nv * DS_MAXD));
  A(dalloc(&D.deg, nv)); A(dalloc(&D.od, nv)); A(dalloc(&D.tc, nv));
  A(dalloc(&D.eoff, S * (c->maxV + 1))); A(dalloc(&D.toff, S * (c->maxV + 1)));
  A(dalloc(&D.meta, S * DSG_META)); A(dalloc(&D.f2v, 2 * nf));
  A(dalloc(&D.o_x, nv)); A(dalloc(&D.o_w1, nv)); A(dalloc(&D.o_w2, nv)); A(dalloc(&D.o_vbar, nv));
  A(dalloc(&D.o_q4, ne)); A(dalloc(&D.o_eij, ne)); A(dalloc(&D.o_eoff, S * (c->maxV + 1)));
  A(dalloc(&U->misc, S * 4));
  A(cudaMallocHost((void**)&U->h_read, sizeof(int32_t) * S * (DSG_META + 4)));
  if (!getenv("FB_UPDATE_NO_FORK")) {
    A(cudaStreamCreateWithFlags(&U->aux, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&U->ev_fork, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&U->ev_join, cudaEventDisableTiming));
  }
  A(cudaMallocHost((void**)&U->h_img, S * (size_t)c->W * c->H));
  A(cudaMallocHost((void**)&U->h_geo, sizeof(float) * S * ((size_t)c->n_slots * 7 + 1)));
  if (const char* e = getenv("FB_UPDATE_GRAPH")) U->use_graph = atoi(e) != 0;
  if (!ok) FB_FAIL(c, FB_E_NOMEM, "fb_update: scratch allocation failed");
  FB_CUDA(c, cudaMemsetAsync(D.meta, 0, sizeof(int32_t) * S * DSG_META, c->stream));
  FB_CUDA(c, cudaMemsetAsync(D.eoff, 0, sizeof(int32_t) * S * (c->maxV + 1), c->stream));
  FB_CUDA(c, cudaMemsetAsync(D.f2v, 0xff, sizeof(int32_t) * 2 * nf, c->stream));
  FB_CUDA(c, cudaMemsetAsync(U->misc, 0, sizeof(int32_t) * S * 4, c->stream));
  // the feature pool is a fixed set of maxF slots: mark them all dead
  FB_CUDA(c, cudaMemsetAsync(c->f_alive, 0, sizeof(int32_t) * nf, c->stream));
  FB_CUDA(c, cudaMemsetAsync(c->f_status, 0, sizeof(int32_t) * nf, c->stream));
  for (int s = 0; s < c->S; ++s) c->hF[s] = c->maxF;
  FB_CUDA(c, cudaMemcpyAsync(c->nF, c->hF.data(), sizeof(int32_t) * c->S, cudaMemcpyHostToDevice, c->stream));
  fb_update_params& p = U->up;
  fb_default_update_params(&p);
  return FB_OK;
}

static void update_free(fb_ctx* c) {
  UpdateState* U = c->upd;
  if (!U) return;
  cudaFree(U->f_ucur); cudaFree(U->f_mucur); cudaFree(U->f_varcur); cudaFree(U->f_valid);
  cudaFree(U->occupied); cudaFree(U->det_xy); cudaFree(U->det_ok); cudaFree(U->det_rank);
  cudaFree(U->det_list); cudaFree(U->free_flag); cudaFree(U->free_rank); cudaFree(U->free_list);
  cudaFree(U->counts); cudaFree(U->o_x); cudaFree(U->o_w1); cudaFree(U->o_w2); cudaFree(U->o_vbar);
  cudaFree(U->o_q4); cudaFree(U->map_v); cudaFree(U->map_e);
  DelGpu& D = U->del;
  cudaFree(D.vxy); cudaFree(D.sxy); cudaFree(D.sid); cudaFree(D.vorder); cudaFree(D.cell_start); cudaFree(D.star);
  cudaFree(D.deg); cudaFree(D.od); cudaFree(D.tc); cudaFree(D.eoff); cudaFree(D.toff); cudaFree(D.meta);
  cudaFree(D.f2v); cudaFree(D.o_x); cudaFree(D.o_w1); cudaFree(D.o_w2); cudaFree(D.o_vbar);
  cudaFree(D.o_q4); cudaFree(D.o_eij); cudaFree(D.o_eoff);
  cudaFree(U->misc);
  if (U->aux) cudaStreamDestroy(U->aux);
  if (U->ev_fork) cudaEventDestroy(U->ev_fork);
  if (U->ev_join) cudaEventDestroy(U->ev_join);
  if (U->h_read) cudaFreeHost(U->h_read);
  if (U->h_img) cudaFreeHost(U->h_img);
  if (U->h_geo) cudaFreeHost(U->h_geo);
  for (auto& st : U->st)
    for (int k = 0; k < 2; ++k)
      if (st.frame_graph[k]) cudaGraphExecDestroy(st.frame_graph[k]);
  delete U;
  c->upd = nullptr;
}

struct StageTimer {
  std::unordered_map<std::string, double>& m;
  const char* key;
  std::chrono::steady_clock::time_point t0;
  StageTimer(std::unordered_map<std::string, double>& mm, const char* k)
      : m(mm), key(k), t0(std::chrono::steady_clock::now()) {}
  ~StageTimer() {
    m[key] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

// Detection on the frame in `img_slot` for cells without a live feature; spawns features that
// reference poseframe slot `ref`.
static int update_detect(fb_ctx* c, int s, int img_slot, int ref) {
  UpdateState* U = c->upd;
  const fb_update_params& p = U->up;
  const int win = p.detection_win_size, cx = c->W / win, cy = c->H / win, cells = cx * cy;
  if (cells > U->maxCells) FB_FAIL(c, FB_E_ARG, "fb_update: detection_win_size too small");
  const size_t fb = (size_t)s * c->maxF, npx = (size_t)c->W * c->H;
  cudaStream_t st = c->stream;
  const uint8_t* img = c->imgs + ((size_t)s * c->n_slots + img_slot) * npx;
  FB_CUDA(c, cudaMemsetAsync(U->occupied, 0, cells, st));
  k_mark_occupied<<<fb_div_up(c->maxF, 256), 256, 0, st>>>(c->W, c->H, win, c->maxF, U->f_ucur + fb, U->f_valid + fb, U->occupied);
  k_detect_features<<<fb_div_up(cells * 32, 256), 256, 0, st>>>(c->W, c->H, win, p.detection_border, p.min_grad_mag, img, U->occupied, U->det_xy, U->det_ok,
                                                               p.do_letterbox ? c->H / 3 : 0, p.do_letterbox ? (2 * c->H) / 3 : c->H);
  k_scan_flags<<<1, 1024, 0, st>>>(cells, U->det_ok, U->det_rank, U->counts);
  k_scatter_ranked<<<fb_div_up(cells, 256), 256, 0, st>>>(cells, U->det_ok, U->det_rank, U->det_list);
  k_free_flags<<<fb_div_up(c->maxF, 256), 256, 0, st>>>(c->maxF, c->f_alive + fb, U->free_flag);
  k_scan_flags<<<1, 1024, 0, st>>>(c->maxF, U->free_flag, U->free_rank, U->counts + 1);
  k_scatter_ranked<<<fb_div_up(c->maxF, 256), 256, 0, st>>>(c->maxF, U->free_flag, U->free_rank, U->free_list);
  // the dense prediction: present when an earlier frame produced a map (host flag), or -- device
  // path, where this frame's outcome is not known on the host yet -- when this frame did
  UpdateStream& S = U->st[s];
  const int32_t* have_dev = S.dev_graph ? U->del.meta + (size_t)s * DSG_META + DSG_NT : nullptr;
  k_spawn_features<<<fb_div_up(std::min(cells, c->maxF), 256), 256, 0, st>>>(
      U->counts, U->det_list, U->free_list, U->det_xy, c->idmap + (size_t)s * npx, c->W, c->H, ref, p.idepth_init,
      p.idepth_var_init, p.init_with_prediction, c->f_uref + fb, c->f_ref + fb, c->f_mu + fb, c->f_var + fb,
      c->f_drop + fb, c->f_alive + fb, c->f_status + fb, S.have_graph ? 1 : 0, have_dev);
  c->launches += 8;
  FB_CUDA(c, cudaGetLastError());
  return FB_OK;
}

// Installs the current frame as a new poseframe in the ring and detects new features in it.
static int update_new_poseframe(fb_ctx* c, int s, int img_id) {
  UpdateState* U = c->upd;
  UpdateStream& S = U->st[s];
  const int cur = c->n_slots - 1, slot = S.pf_next;
  const size_t fb = (size_t)s * c->maxF, npx = (size_t)c->W * c->H;
  cudaStream_t st = c->stream;
  if (S.pf_img_id[slot] >= 0) {  // ring wraps: features anchored in the evicted poseframe die
    k_kill_ref_slot<<<fb_div_up(c->maxF, 256), 256, 0, st>>>(c->maxF, slot, c->f_alive + fb, c->f_ref + fb);
    c->launches++;
  }
  uint8_t* base = c->imgs + (size_t)s * c->n_slots * npx;
  FB_CUDA(c, cudaMemcpyAsync(base + (size_t)slot * npx, base + (size_t)cur * npx, npx, cudaMemcpyDeviceToDevice, st));
  memcpy(&c->h_pose[((size_t)s * c->n_slots + slot) * 7], &c->h_pose[((size_t)s * c->n_slots + cur) * 7], sizeof(float) * 7);
  S.pf_img_id[slot] = img_id;
  S.pf_next = (slot + 1) % (c->n_slots - 1);
  S.have_poseframe = true;
  return update_detect(c, s, cur, slot);
}

// Unfiltered dense map of the stream's current mesh into c->idmap (kept on the device: it is the
// prediction source of the next frames).  Tdev != NULL: the triangle count lives on the device.
static int update_interpolate(fb_ctx* c, int s, const int32_t* Tdev, int32_t* covered) {
  const int T = Tdev ? c->maxT : c->hT[s];
  const size_t npx = (size_t)c->W * c->H, vb = (size_t)s * c->maxV;
  int32_t* owner = c->owner + (size_t)s * npx;
  const int32_t* tri = c->tri + (size_t)s * c->maxT * 3;
  cudaStream_t st = c->stream;
  const UpdateStream& S = c->upd->st[s];
  const bool spec = S.spec_on && c->owner2 && c->idmap_f;
  uint8_t* valid = c->tri_valid + (size_t)s * c->maxT;
  int32_t* owner2 = spec ? c->owner2 + (size_t)s * npx : nullptr;
  float* map2 = spec ? c->idmap_f + (size_t)s * npx : nullptr;
  ProfScope ps(c, FB_PROF_INTERP);
  // (the ownership maps are kept clean by the shading pass; the coverage counter is zeroed by the claim pass)
  if (covered && !T) FB_CUDA(c, cudaMemsetAsync(covered, 0, sizeof(int32_t), st));
  if (T) {  // unfiltered: every triangle is valid, no validity pass; the filtered map rides along when asked for
    if (spec) {
      const float cos_thresh = (float)cos((double)S.spec_filter.oblique_normal_thresh);
      k_tri_validity<<<fb_div_up(T, 256), 256, 0, st>>>(c->W, c->d_K + 9 * s, c->vpos + vb, c->x + vb, T, tri, S.spec_filter, cos_thresh, 1, valid, Tdev);
      c->launches += 1;
    }
    k_raster_claim<<<fb_div_up(T * 32, 256), 256, 0, st>>>(c->W, c->H, c->vpos + vb, T, tri, nullptr, owner, Tdev, spec ? valid : nullptr, owner2, covered);
    c->launches += 1;
  }
  k_raster_shade<<<fb_div_up((int)npx, 256), 256, 0, st>>>(c->W, c->H, c->vpos + vb, c->x + vb, tri, owner, c->idmap + (size_t)s * npx, Tdev, covered);
  c->launches++;
  if (spec && T) {  // (one shading pass writing both maps was measured 3x slower than this second launch)
    k_raster_shade<<<fb_div_up((int)npx, 256), 256, 0, st>>>(c->W, c->H, c->vpos + vb, c->x + vb, tri, owner2, map2, Tdev, nullptr);
    c->launches++;
  }
  FB_CUDA(c, cudaGetLastError());
  return FB_OK;
}

// regularization/nltgv2/rescale_data around the solve of stream s (dir 0 before, 1 after).
static int update_rescale(fb_ctx* c, int s, int dir, float* z_keep) {
  UpdateState* U = c->upd;
  const size_t vb = (size_t)s * c->maxV;
  cudaStream_t st = c->stream;
  float* scale = reinterpret_cast<float*>(U->misc + 4 * s + 2);
  if (dir == 0) {
    k_rescale_mean<<<1, DSG_THREADS, 0, st>>>(c->nV, s, c->z + vb, scale);
    c->launches++;
  }
  k_rescale_apply<<<fb_div_up(c->maxV, 256), 256, 0, st>>>(c->nV, s, dir, scale, c->z + vb, z_keep, c->x + vb, c->w1 + vb,
                                                          c->w2 + vb, c->vbar + vb);
  c->launches++;
  FB_CUDA(c, cudaGetLastError());
  return FB_OK;
}

// sync_graph + triangulate on the device (delaunay_gpu.cuh): six launches, no host involvement.
static int update_graph_device(fb_ctx* c, int s) {
  UpdateState* U = c->upd;
  UpdateStream& S = U->st[s];
  const fb_update_params& p = U->up;
  DelGpu& D = U->del;
  const size_t fb = (size_t)s * c->maxF, vb = (size_t)s * c->maxV, eb = (size_t)s * c->maxE;
  const size_t npx = (size_t)c->W * c->H, nf = (size_t)c->S * c->maxF;
  cudaStream_t st = c->stream;
  const int par = (int)(S.builds & 1);
  DsgGraph gg{c->x, c->w1, c->w2, c->vbar, c->q4, c->eij};
  const int use_height = (p.min_height > -1e13f || p.max_height < 1e13f) ? 1 : 0;
  DsgSelect q{U->f_ucur + fb, U->f_varcur + fb, U->f_valid + fb, p.idepth_var_max_graph, c->maxF, c->maxV, c->W, c->H,
              use_height, U->f_mucur + fb, c->d_K + 9 * s, c->d_pose + ((size_t)s * c->n_slots + (c->n_slots - 1)) * 7,
              p.min_height, p.max_height};
  k_ds_prepare<<<1, DSG_THREADS, 0, st>>>(q, D, s, c->vfeat + vb, c->vpos + vb, D.f2v + par * nf + fb, c->nV + s);
  // per-topology tables of the resident solvers (variants 2 / 3) do not describe this graph
  if (c->plan && s < (int)c->plan->topo.size()) { c->plan->topo[s].V = 0; c->plan->topo[s].dirty = true; }
  if (c->gplan) { c->gplan->topo[s].dirty = true; c->gplan->topo[s].planned = 0; c->gplan->version++; }
  tile_plan_mark(c, s);  // positions moved: the tiles of variant 5 are re-cut (one small kernel)
  // Side branch while the stars are computed (they take the whole GPU for ~75 us, these two a few SMs
  // for ~17 us): the previous graph's state aside (k_ds_prepare left its counts in OLD_NV / OLD_NE; the
  // copy must be through before k_ds_scan rewrites the edge offsets) and the solver's tiles from the
  // new vertex positions.  Inside a captured frame the events become fork / join edges of the graph.
  cudaStream_t side = U->aux ? U->aux : st;
  if (U->aux) {
    FB_CUDA(c, cudaEventRecord(U->ev_fork, st));
    FB_CUDA(c, cudaStreamWaitEvent(U->aux, U->ev_fork, 0));
  }
  k_ds_stash<<<64, 256, 0, side>>>(gg, D, s, c->maxV, c->maxE);
  tile_assign_early(c, s, side);
  if (U->aux) FB_CUDA(c, cudaEventRecord(U->ev_join, U->aux));
  k_ds_stars<<<dsg_stars_grid(c->device, c->maxV), DSG_WARPS * 32, 0, st>>>(D, s, c->maxV);
  if (U->aux) FB_CUDA(c, cudaStreamWaitEvent(st, U->ev_join, 0));
  k_ds_scan<<<1, DSG_THREADS, 0, st>>>(D, s, c->maxV, c->maxE, c->maxT, c->row + (size_t)s * (c->maxV + 1), c->nE + s, c->nT + s);
  k_ds_emit<<<fb_div_up(c->maxV, 4), 128, 0, st>>>(D, s, c->maxV, c->maxE, c->maxT, c->vpos + vb, c->eij + eb, c->ec + eb,
                                                    c->tri + (size_t)s * c->maxT * 3);
  DsgCarry cq{U->f_mucur + fb, U->f_varcur + fb, D.f2v + (1 - par) * nf + fb,
              S.have_graph ? c->idmap + (size_t)s * npx : nullptr, c->W, c->H, p.adaptive_data_weights, p.init_with_prediction};
  k_ds_csr<<<fb_div_up(c->maxE > c->maxV ? c->maxE : c->maxV, 128), 128, 0, st>>>(D, cq, s, c->maxV, c->maxE, c->vfeat + vb, c->vpos + vb, c->eij + eb,
                                                   c->row + (size_t)s * (c->maxV + 1), c->inc + 2 * eb, c->z + vb, c->wt + vb,
                                                   c->x + vb, c->w1 + vb, c->w2 + vb, c->vbar + vb, c->q4 + eb, c->epos + eb, c->vnin + vb);
  c->launches += 6;
  FB_CUDA(c, cudaGetLastError());
  S.builds++;
  S.dev_graph = true;
  return FB_OK;
}

// sync_graph + triangulate on the host (delaunay.h) -- the reference path the device one is tested
// against; selected with fb_update_params::triangulator = 1.  Returns 1 when a graph was built.
static int update_graph_host(fb_ctx* c, int s) {
  UpdateState* U = c->upd;
  UpdateStream& S = U->st[s];
  const fb_update_params& p = U->up;
  const size_t fb = (size_t)s * c->maxF, vb = (size_t)s * c->maxV, eb = (size_t)s * c->maxE;
  const size_t npx = (size_t)c->W * c->H;
  cudaStream_t st = c->stream;
  int rc;
  std::vector<float2> h_u(c->maxF);
  std::vector<float> h_var(c->maxF), h_mu;
  std::vector<int32_t> h_valid(c->maxF);
  const bool use_height = p.min_height > -1e13f || p.max_height < 1e13f;
  {
    StageTimer t(S.stats, "project_features");
    FB_CUDA(c, cudaMemcpyAsync(h_u.data(), U->f_ucur + fb, sizeof(float2) * c->maxF, cudaMemcpyDeviceToHost, st));
    FB_CUDA(c, cudaMemcpyAsync(h_var.data(), U->f_varcur + fb, sizeof(float) * c->maxF, cudaMemcpyDeviceToHost, st));
    FB_CUDA(c, cudaMemcpyAsync(h_valid.data(), U->f_valid + fb, sizeof(int32_t) * c->maxF, cudaMemcpyDeviceToHost, st));
    if (use_height) {
      h_mu.resize(c->maxF);
      FB_CUDA(c, cudaMemcpyAsync(h_mu.data(), U->f_mucur + fb, sizeof(float) * c->maxF, cudaMemcpyDeviceToHost, st));
    }
    FB_CUDA(c, cudaStreamSynchronize(st));
  }
  std::vector<int32_t> vfeat;
  std::vector<float> pos;
  int n_valid = 0;
  for (int f = 0; f < c->maxF; ++f) {
    n_valid += h_valid[f] ? 1 : 0;
    if (h_valid[f] && h_var[f] < p.idepth_var_max_graph && (int)vfeat.size() < c->maxV) {  // valid implies alive
      if (use_height) {
        const float h = dsg_world_height(&c->h_K[9 * s], &c->h_pose[((size_t)s * c->n_slots + (c->n_slots - 1)) * 7], h_u[f].x, h_u[f].y, h_mu[f]);
        if (!(h >= p.min_height && h <= p.max_height)) continue;
      }
      vfeat.push_back(f);
      pos.push_back(h_u[f].x);
      pos.push_back(h_u[f].y);
    }
  }
  S.stats["num_feats"] = n_valid;
  const int V = (int)vfeat.size();
  S.stats["num_vtx"] = V;
  std::vector<int> tris, edges;
  bool have_tri = false;
  if (V >= 3) {
    StageTimer t(S.stats, "triangulate");
    if ((int)U->tri.size() < c->S) U->tri.resize(c->S);
    have_tri = U->tri[s].run(V, pos.data(), tris, edges);
    if ((int)edges.size() / 2 > c->maxE || (int)tris.size() / 3 > c->maxT) have_tri = false;
    if (tris.empty()) have_tri = false;
  }
  if (!have_tri) return 0;
  const int E = (int)edges.size() / 2;
  StageTimer t(S.stats, "sync_graph");
  // maps new -> old for the state carry-over
  std::vector<int32_t> map_v(V, -1), map_e(E, -1);
  if (S.have_graph) {
    std::vector<int32_t> f2v(c->maxF, -1);
    for (size_t k = 0; k < S.vert_feat.size(); ++k) f2v[S.vert_feat[k]] = (int32_t)k;
    for (int k = 0; k < V; ++k) map_v[k] = f2v[vfeat[k]];
    // vertices are listed in ascending feature index in both graphs, so both canonical edge
    // lists are sorted by (feature_i, feature_j): one linear merge matches the persisting edges
    const size_t oEn = S.edges.size() / 2;
    size_t o = 0;
    for (int e = 0; e < E; ++e) {
      const uint64_t key = ((uint64_t)vfeat[edges[2 * e]] << 32) | (uint32_t)vfeat[edges[2 * e + 1]];
      while (o < oEn && (((uint64_t)S.vert_feat[S.edges[2 * o]] << 32) | (uint32_t)S.vert_feat[S.edges[2 * o + 1]]) < key) ++o;
      if (o < oEn && (((uint64_t)S.vert_feat[S.edges[2 * o]] << 32) | (uint32_t)S.vert_feat[S.edges[2 * o + 1]]) == key)
        map_e[e] = (int32_t)o;
    }
    // stash the old state (device to device)
    const int oV = (int)S.vert_feat.size(), oE = (int)S.edges.size() / 2;
    FB_CUDA(c, cudaMemcpyAsync(U->o_x, c->x + vb, sizeof(float) * oV, cudaMemcpyDeviceToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(U->o_w1, c->w1 + vb, sizeof(float) * oV, cudaMemcpyDeviceToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(U->o_w2, c->w2 + vb, sizeof(float) * oV, cudaMemcpyDeviceToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(U->o_vbar, c->vbar + vb, sizeof(float4) * oV, cudaMemcpyDeviceToDevice, st));
    if (oE) FB_CUDA(c, cudaMemcpyAsync(U->o_q4, c->q4 + eb, sizeof(float4) * oE, cudaMemcpyDeviceToDevice, st));
  }
  // edge weights: alpha = 1/|delta| in pixels, beta = 1
  std::vector<float> alpha(E), beta(E, 1.0f);
  for (int e = 0; e < E; ++e) {
    const float dx = pos[2 * edges[2 * e]] - pos[2 * edges[2 * e + 1]];
    const float dy = pos[2 * edges[2 * e] + 1] - pos[2 * edges[2 * e + 1] + 1];
    alpha[e] = 1.0f / sqrtf(dx * dx + dy * dy);
  }
  rc = fb_graph_set(c, s, V, E, pos.data(), edges.data(), alpha.data(), beta.data());
  if (rc) return rc;
  FB_CUDA(c, cudaMemcpyAsync(c->vfeat + vb, vfeat.data(), sizeof(int32_t) * V, cudaMemcpyHostToDevice, st));
  FB_CUDA(c, cudaMemcpyAsync(U->map_v, map_v.data(), sizeof(int32_t) * V, cudaMemcpyHostToDevice, st));
  if (E) FB_CUDA(c, cudaMemcpyAsync(U->map_e, map_e.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice, st));
  k_data_from_projection<<<fb_div_up(V, 256), 256, 0, st>>>(V, c->z + vb, c->wt + vb, c->vfeat + vb, U->f_mucur + fb, U->f_varcur + fb, p.adaptive_data_weights);
  const float* idmap = S.have_graph ? c->idmap + (size_t)s * npx : nullptr;
  k_remap_state<<<fb_div_up(std::max(V, E), 256), 256, 0, st>>>(
      V, E, U->map_v, U->map_e, U->o_x, U->o_w1, U->o_w2, U->o_vbar, U->o_q4, c->z + vb, c->vpos + vb, idmap,
      c->W, c->H, p.init_with_prediction, c->x + vb, c->w1 + vb, c->w2 + vb, c->vbar + vb, c->q4 + eb);
  c->launches += 2;
  FB_CUDA(c, cudaGetLastError());
  FB_CUDA(c, cudaStreamSynchronize(st));  // the pageable staging vectors above go out of scope
  S.vert_feat.assign(vfeat.begin(), vfeat.end());
  S.edges.assign(edges.begin(), edges.end());
  S.tris.assign(tris.begin(), tris.end());
  S.dev_graph = false;
  rc = fb_mesh_set(c, s, (int)S.tris.size() / 3, S.tris.data());
  if (rc) return rc;
  S.stats["num_edges"] = E;
  S.stats["num_tris"] = (double)S.tris.size() / 3;
  return 1;
}

// update_idepths + project_features of stream s against the current frame.  captured: the poses and
// the comparison slot come from the stream's pinned staging record (inputs of the frame graph) and
// only this stream's geometry / search kernels are launched.
static int update_idepths_enqueue(fb_ctx* c, int s, bool captured) {
  UpdateState* U = c->upd;
  UpdateStream& S = U->st[s];
  const int cur = c->n_slots - 1;
  const size_t fb = (size_t)s * c->maxF;
  cudaStream_t st = c->stream;
  int32_t* misc = U->misc + 4 * s;
  StageTimer t(S.stats, "update_idepths");
  if (!captured) {
    std::vector<int32_t>& cmp = U->cmp_scratch;
    cmp.assign(c->S, -1);
    cmp[s] = cur;
    const int rc = fb_idepth_update(c, cmp.data());  // also refreshes the geometry table against `cur`
    if (rc) return rc;
  } else {
    const size_t np = (size_t)c->n_slots * 7;
    const float* hg = U->h_geo + (size_t)s * (np + 1);
    // poses + comparison slot are read by the kernel from the pinned staging record of the frame graph
    k_epi_geometry<<<1, std::max(32, c->n_slots), 0, st>>>(hg, c->d_K, reinterpret_cast<const int32_t*>(hg + np), c->n_slots,
                                                           c->d_geo, s, c->d_pose, c->d_cmp, c->counters);
    EpiArgs a;
    a.imgs = c->imgs; a.geo = c->d_geo; a.cmp_slot = c->d_cmp; a.u_ref = c->f_uref;
    a.ref_slot = c->f_ref; a.mu = c->f_mu; a.var = c->f_var; a.dropouts = c->f_drop;
    a.alive = c->f_alive; a.status = c->f_status; a.u_cmp = c->f_ucmp; a.nF = c->nF;
    a.counters = c->counters; a.W = c->W; a.H = c->H; a.n_slots = c->n_slots; a.maxF = c->maxF;
    a.s0 = s;
    a.cmp_frames = nullptr;
    a.p = c->epi;
    const int wpb = 8;
    const size_t smem = sizeof(float) * wpb * FB_EPI_GROUPS * (2 * c->epi.max_search_px + 2 * FB_MAX_WIN + 2);
    const dim3 grid(fb_div_up(c->maxF, wpb * FB_EPI_GROUPS), 1);
    k_epipolar_search<<<grid, wpb * 32, smem, st>>>(a);
    c->launches += 2;
  }
  FB_CUDA(c, cudaMemsetAsync(misc, 0, sizeof(int32_t) * 4, st));
  k_project_kill<<<fb_div_up(c->maxF, 256), 256, 0, st>>>(
      c->d_geo, c->n_slots, s, c->maxF, c->W, c->H, c->f_uref + fb, c->f_ref + fb, c->f_mu + fb, c->f_var + fb,
      c->f_alive + fb, U->f_ucur + fb, U->f_mucur + fb, U->f_varcur + fb, U->f_valid + fb, misc + 1);
  c->launches += 1;
  FB_CUDA(c, cudaGetLastError());
  return FB_OK;
}

// The steady-state frame of the device path on c->stream: [frame upload from the staging buffer]
// epipolar update -> projection -> device triangulation + graph sync -> solve -> interpolation ->
// readback of the counts.  Everything is asynchronous; captured = being recorded into the frame graph.
static int update_frame_enqueue(fb_ctx* c, int s, bool captured) {
  UpdateState* U = c->upd;
  UpdateStream& S = U->st[s];
  const fb_update_params& p = U->up;
  cudaStream_t st = c->stream;
  int32_t* misc = U->misc + 4 * s;
  int rc;
  // (the frame itself is uploaded by the caller in front of the graph launch: its source is the caller's
  // buffer when that is pinned, so it cannot be a node with a fixed address)
  rc = update_idepths_enqueue(c, s, captured);
  if (rc) return rc;
  {
    StageTimer t(S.stats, "sync_graph");
    rc = update_graph_device(c, s);
    if (rc) return rc;
  }
  if (p.do_nltgv2 && p.iters > 0) {
    StageTimer t(S.stats, "nltgv2");
    float* z_keep = U->del.o_x + (size_t)s * c->maxV;  // free between k_ds_csr and the next frame's stash
    if (p.rescale_data && (rc = update_rescale(c, s, 0, z_keep)) != 0) return rc;
    rc = fb_nltgv2_solve_stream(c, s, p.iters, &p.rparams);
    if (rc) return rc;
    if (p.rescale_data && (rc = update_rescale(c, s, 1, z_keep)) != 0) return rc;
  }
  {
    StageTimer t(S.stats, "interpolate");
    rc = update_interpolate(c, s, c->nT + s, misc);
    if (rc) return rc;
  }
  int32_t* h = U->h_read + (size_t)s * (DSG_META + 4);
  FB_CUDA(c, cudaMemcpyAsync(h, U->del.meta + (size_t)s * DSG_META, sizeof(int32_t) * DSG_META, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(h + DSG_META, misc, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, st));
  return FB_OK;
}

static int fb_update_impl(fb_ctx* c, int s, double /*time*/, int img_id, const float pose[7],
                          const uint8_t* gray, int pitch, int is_poseframe) {
  int rc = update_alloc(c);
  if (rc) return rc;
  UpdateState* U = c->upd;
  UpdateStream& S = U->st[s];
  const fb_update_params& p = U->up;
  const int cur = c->n_slots - 1;
  const size_t fb = (size_t)s * c->maxF;
  const size_t npx = (size_t)c->W * c->H;
  cudaStream_t st = c->stream;
  const bool dev = p.triangulator == 0;
  S.stats.clear();
  StageTimer t_all(S.stats, "update");
  // the pose always goes to the host mirror; the image upload is enqueued below (directly, or from the
  // pinned staging buffer when the frame is replayed as a graph)
  rc = fb_frame_pose_set(c, s, cur, pose);
  if (rc) return rc;
  if (pitch < c->W) FB_FAIL(c, FB_E_ARG, "fb_update: pitch < width");
  S.frames++;
  if (!S.have_poseframe) {  // very first frame: it becomes the first poseframe, nothing to estimate yet
    rc = fb_frame_set(c, s, cur, gray, pitch, pose);
    if (rc) return rc;
    StageTimer t(S.stats, "detection");
    // no projections yet: every cell is free
    FB_CUDA(c, cudaMemsetAsync(U->f_valid + fb, 0, sizeof(int32_t) * c->maxF, st));
    rc = update_new_poseframe(c, s, img_id);
    if (rc) return rc;
    FB_CUDA(c, cudaStreamSynchronize(st));
    return 0;
  }

  int32_t* misc = U->misc + 4 * s;
  int updated = 0;
  if (dev) {
    // ---- device path: the whole frame is enqueued; ONE synchronisation at the end --------------
    int32_t* h = U->h_read + (size_t)s * (DSG_META + 4);
    // Steady state (a dense map exists, every lazy allocation has happened, no event profiling): the
    // frame is replayed as a CUDA graph captured once per f2v parity.
    const bool graph_ok = U->use_graph && !c->prof && S.have_graph && S.builds >= 4 && pitch == c->W;
    if (graph_ok) {
      StageTimer t(S.stats, "sync_graph");
      const int par = (int)(S.builds & 1);
      if (!S.frame_graph[par]) {
        cudaGraph_t graph = nullptr;
        const int64_t l0 = c->launches;
        FB_CUDA(c, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        rc = update_frame_enqueue(c, s, true);
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc || ce != cudaSuccess) {
          if (graph) cudaGraphDestroy(graph);
          cudaGetLastError();
          U->use_graph = false;
          if (rc) return rc;
          FB_FAIL(c, FB_E_CUDA, std::string("fb_update: graph capture failed: ") + cudaGetErrorString(ce));
        }
        ce = cudaGraphInstantiate(&S.frame_graph[par], graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
          S.frame_graph[par] = nullptr;
          FB_FAIL(c, FB_E_CUDA, std::string("fb_update: cudaGraphInstantiate: ") + cudaGetErrorString(ce));
        }
        S.graph_launches[par] = c->launches - l0;
        c->launches = l0;
        // the capture advanced the build counter and marked the plans; undo the double count
        S.builds--;
      }
      // inputs of the graph: the frame, uploaded in front of it -- straight from the caller's buffer when
      // that is pinned / registered memory (a capture driver's ring), through the staging buffer otherwise
      // (copying 300 kB on the host costs ~15 us during which the GPU has nothing to do) -- and the poses
      // in their pinned staging record
      {
        const uint8_t* src = gray;
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, gray) != cudaSuccess || pa.type != cudaMemoryTypeHost) {
          cudaGetLastError();
          memcpy(U->h_img + (size_t)s * npx, gray, npx);
          src = U->h_img + (size_t)s * npx;
        }
        FB_CUDA(c, cudaMemcpyAsync(c->imgs + ((size_t)s * c->n_slots + cur) * npx, src, npx, cudaMemcpyHostToDevice, st));
      }
      float* hg = U->h_geo + (size_t)s * ((size_t)c->n_slots * 7 + 1);
      memcpy(hg, &c->h_pose[(size_t)s * c->n_slots * 7], sizeof(float) * c->n_slots * 7);
      const int32_t cur32 = cur;
      memcpy(hg + (size_t)c->n_slots * 7, &cur32, sizeof(int32_t));
      FB_CUDA(c, cudaGraphLaunch(S.frame_graph[par], st));
      c->launches += S.graph_launches[par];
      S.builds++;
      S.dev_graph = true;
      tile_plan_mark(c, s);
      if (c->tplan && s < (int)c->tplan->dirty.size()) c->tplan->dirty[s] = 0;  // k_tile_assign ran inside the graph
    } else {
      {
        StageTimer t(S.stats, "frame_creation");
        rc = fb_frame_set(c, s, cur, gray, pitch, pose);
        if (rc) return rc;
      }
      rc = update_frame_enqueue(c, s, false);
      if (rc) return rc;
    }
    if (is_poseframe) {
      StageTimer t(S.stats, "detection");
      rc = update_new_poseframe(c, s, img_id);
      if (rc) return rc;
    }
    FB_CUDA(c, cudaStreamSynchronize(st));
    if (fb_tile_failed(c)) {
      // this frame was not solved; the following ones use the next solver in line
      tile_clear_failure(c, true);
      FB_FAIL(c, FB_E_STATE, "fb_update: the tile-resident solver rejected this frame's graph (a tile exceeded its capacity); it is disabled for this context");
    }
    if (fb_coop_failed(c)) FB_FAIL(c, FB_E_STATE, "fb_update: resident solver rejected the graph size");
    const int err = h[DSG_ERR];
    c->hV[s] = h[DSG_NV];
    c->hE[s] = h[DSG_NE];
    c->hT[s] = h[DSG_NT];
    if (err) {
      char msg[96];
      snprintf(msg, sizeof(msg), "fb_update: device triangulation failed (flags 0x%x); set triangulator = 1", err);
      FB_FAIL(c, FB_E_STATE, msg);
    }
    updated = h[DSG_NT] > 0 ? 1 : 0;
    if (updated) S.have_graph = true;
    // the filtered map rendered beside the unfiltered one belongs to this state of the context
    S.spec_epoch = (updated && S.spec_on && c->idmap_f) ? c->mut_epoch : ~0ull;
    S.stats["num_feats"] = h[DSG_META + 1];
    S.stats["num_vtx"] = h[DSG_NV];
    S.stats["num_edges"] = h[DSG_NE];
    S.stats["num_tris"] = h[DSG_NT];
    S.stats["coverage"] = (double)h[DSG_META] / (double)npx;
  } else {
    // ---- host path: D2H of the projected features, host Delaunay, upload --------------------
    rc = fb_frame_set(c, s, cur, gray, pitch, pose);
    if (rc) return rc;
    rc = update_idepths_enqueue(c, s, false);
    if (rc) return rc;
    rc = update_graph_host(c, s);
    if (rc < 0) return rc;
    if (rc == 1) {
      S.have_graph = true;
      if (p.do_nltgv2 && p.iters > 0) {
        StageTimer t(S.stats, "nltgv2");
        // solve only this stream's graph: other streams of the context keep their own cadence
        if (p.rescale_data && (rc = update_rescale(c, s, 0, U->o_x)) != 0) return rc;
        rc = fb_nltgv2_solve_stream(c, s, p.iters, &p.rparams);
        if (rc) return rc;
        if (p.rescale_data && (rc = update_rescale(c, s, 1, U->o_x)) != 0) return rc;
      }
      {
        StageTimer t(S.stats, "interpolate");
        rc = update_interpolate(c, s, nullptr, misc);
        if (rc) return rc;
      }
      updated = 1;
    }
    if (is_poseframe) {
      StageTimer t(S.stats, "detection");
      rc = update_new_poseframe(c, s, img_id);
      if (rc) return rc;
    }
    int32_t* h = U->h_read + (size_t)s * (DSG_META + 4);
    FB_CUDA(c, cudaMemcpyAsync(h + DSG_META, misc, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, st));
    FB_CUDA(c, cudaStreamSynchronize(st));
    if (updated) S.stats["coverage"] = (double)h[DSG_META] / (double)npx;
    S.spec_epoch = (updated && c->hT[s] > 0 && S.spec_on && c->idmap_f) ? c->mut_epoch : ~0ull;
    if (!S.stats.count("num_edges")) { S.stats["num_edges"] = 0.0; S.stats["num_tris"] = 0.0; }
  }
  // aliases of round 1 (kept for callers of fb_get_stat)
  S.stats["num_vertices"] = S.stats["num_vtx"];
  S.stats["num_triangles"] = S.stats["num_tris"];
  if (is_poseframe) S.stats["keyframe"] = S.stats.count("detection") ? S.stats["detection"] : 0.0;
  return updated;
}
