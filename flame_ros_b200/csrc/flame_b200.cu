// flame_b200.cu -- C-ABI of libflame_b200.so (see include/flame_b200.h for the contract and the
// reference interfaces each entry point replaces).  Host code is plain C++17; all arithmetic of the
// hot path runs in the CUDA kernels of nltgv2*.cuh / epipolar.cuh / raster.cuh.  There is no CPU
// fallback: every compute entry point launches kernels on ctx->stream.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "delaunay.h"
#include "epipolar.cuh"
#include "nltgv2.cuh"
#include "nltgv2_cluster.cuh"
#include "nltgv2_grid.cuh"
#include "nltgv2_coop.cuh"
#include "nltgv2_tile.cuh"
#include "raster.cuh"
#include "frontend.cuh"

static std::string g_create_error;
static void update_free(fb_ctx* c);
static void update_mark_host_graph(fb_ctx* c, int s);
static void update_invalidate_graphs(fb_ctx* c);
static void tile_plan_free(fb_ctx* c);
static void tile_plan_mark(fb_ctx* c, int s);
static void tile_assign_early(fb_ctx* c, int s, cudaStream_t st);
static bool fb_tile_failed(const fb_ctx* c);
static void tile_clear_failure(fb_ctx* c, bool disable);

// ------------------------------------------------------------------------------------ helpers

static void prof_fold(fb_ctx* c, int section) {
  ProfSection& s = c->sec[section];
  for (size_t i = 0; i + 1 < s.ev.size(); i += 2) {
    float ms = 0.f;
    cudaEventSynchronize(s.ev[i + 1]);
    if (cudaEventElapsedTime(&ms, s.ev[i], s.ev[i + 1]) == cudaSuccess) s.total_ms += ms;
    c->prof_free.push_back(s.ev[i]);
    c->prof_free.push_back(s.ev[i + 1]);
  }
  s.ev.clear();
}

// Work issued by the pipelined fb_hotpath_step lives on auxiliary streams; every other entry point
// first waits for it so the single-stream ordering the rest of the API assumes still holds.
static inline void pipeline_drain(fb_ctx* c) {
  if (!c->pipe_dirty || c->pipe_hold) return;
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->solve_stream) cudaStreamSynchronize(c->solve_stream);
  if (c->out_stream) cudaStreamSynchronize(c->out_stream);
  cudaStreamSynchronize(c->stream);
  c->pipe_dirty = false;
}
// The current CUDA device is per host thread: a context created on device 1 and then driven from a
// worker thread (whose current device is still 0) would launch on streams that do not exist there
// ("invalid resource handle": found by the 2-GPU bench, rank 1).  Every entry point therefore makes
// the context's device current for the calling thread first.
static inline void bind_device(const fb_ctx* c) {
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess || d != c->device) cudaSetDevice(c->device);
}
#define CHECK_CTX_NODRAIN(c) \
  if (!(c)) return FB_E_ARG; \
  bind_device(c)
#define CHECK_CTX(c)           \
  if (!(c)) return FB_E_ARG;   \
  bind_device(c);              \
  ++(c)->mut_epoch;            \
  pipeline_drain(c)
// read-only entry points (getters): they leave the epoch alone, so the filtered map fb_update rendered
// stays current across them
#define CHECK_CTX_RO(c)        \
  if (!(c)) return FB_E_ARG;   \
  bind_device(c);              \
  pipeline_drain(c)
#define CHECK_STREAM(c, s) \
  if ((s) < 0 || (s) >= (c)->S) FB_FAIL(c, FB_E_ARG, "stream index out of range")

// ------------------------------------------------------------------------------------ lifecycle
extern "C" int fb_version(void) { return 100; }

extern "C" void fb_default_epi_params(fb_epi_params* p) {
  // win 5 / min_grad 5 / line var 4 / dropouts 5: /root/reference/cfg/flame_nodelet.yaml:69-75
  *p = fb_epi_params{5, 5.0f, 4.0f, 5, 2.0f, 400.0f, 1.5f, 2, 4.0f, 1.0f, 0.0f, 10.0f, 64, 0.5f};
}
extern "C" void fb_default_nltgv2_params(fb_nltgv2_params* p) {
  // /root/reference/cfg/flame_nodelet.yaml:86-89
  *p = fb_nltgv2_params{0.15f, 0.001f, 125.0f, 0.25f, 0.0f, 10.0f};
}
extern "C" void fb_default_tri_filter_params(fb_tri_filter_params* p) {
  // /root/reference/cfg/flame_nodelet.yaml:31-46
  *p = fb_tri_filter_params{1, 1.57f, 0.35f, 0.1f, 1, 0.333f, 1, 0.01f};
}

extern "C" const char* fb_last_error(const fb_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" void* fb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void fb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

static void free_all(fb_ctx* c) {
  if (c->z_buf[0]) { c->z = c->z_buf[0]; c->wt = c->wt_buf[0]; }  // the originals; the second buffers below
  cudaFree(c->z_buf[1]); cudaFree(c->wt_buf[1]);
  for (int b = 0; b < 2; ++b) if (c->ev_zfree[b]) cudaEventDestroy(c->ev_zfree[b]);
  cudaFree(c->vbar); cudaFree(c->x); cudaFree(c->w1); cudaFree(c->w2); cudaFree(c->z);
  cudaFree(c->wt); cudaFree(c->ec); cudaFree(c->eij); cudaFree(c->q4); cudaFree(c->row);
  cudaFree(c->inc); cudaFree(c->epos); cudaFree(c->vnin); cudaFree(c->nV); cudaFree(c->nE); cudaFree(c->vfeat); cudaFree(c->vpos);
  cudaFree(c->costs); cudaFree(c->imgs); cudaFree(c->d_pose); cudaFree(c->d_K);
  cudaFree(c->d_cmp); cudaFree(c->d_geo); cudaFree(c->pool); cudaFree(c->f_uref);
  cudaFree(c->f_mu); cudaFree(c->f_var); cudaFree(c->f_drop); cudaFree(c->f_alive);
  cudaFree(c->f_ref); cudaFree(c->f_status); cudaFree(c->f_ucmp); cudaFree(c->nF);
  cudaFree(c->counters); cudaFree(c->tri); cudaFree(c->nT); cudaFree(c->tri_valid);
  cudaFree(c->owner); cudaFree(c->idmap);
  cluster_plan_free(c);
  grid_plan_free(c);
  tile_plan_free(c);
  update_free(c);
  for (cudaGraphExec_t e : c->solve_exec) if (e) cudaGraphExecDestroy(e);
  cudaFree(c->coop_contrib); cudaFree(c->idmap_scratch); cudaFree(c->incoming);
  cudaFree(c->x_stage[0]); cudaFree(c->x_stage[1]);
  cudaFree(c->idmap_f); cudaFree(c->owner2);
  if (c->coop_err) cudaFreeHost(c->coop_err);
  for (int k = 0; k < FB_PROF_NUM; ++k)
    for (cudaEvent_t e : c->sec[k].ev) cudaEventDestroy(e);
  for (cudaEvent_t e : c->prof_free) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_ready) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_free) cudaEventDestroy(e);
  for (int k = 0; k < 4; ++k) if (c->ev_result[k]) cudaEventDestroy(c->ev_result[k]);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_join2) cudaEventDestroy(c->ev_join2);
  if (c->ev_epi) cudaEventDestroy(c->ev_epi);
  if (c->ev_asm) cudaEventDestroy(c->ev_asm);
  if (c->stage) cudaFreeHost(c->stage);
  for (cudaEvent_t e : c->stage_ev) if (e) cudaEventDestroy(e);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->solve_stream) cudaStreamDestroy(c->solve_stream);
  if (c->out_stream) cudaStreamDestroy(c->out_stream);
  if (c->ev_solved) cudaEventDestroy(c->ev_solved);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
}

extern "C" fb_ctx* fb_create(int device, int n_streams, int width, int height, int n_slots,
                             int max_features, int max_vertices, int max_edges,
                             void* cuda_stream) {
  if (n_streams < 1 || width < 16 || height < 16 || n_slots < 2 || n_slots > 1024 ||
      max_features < 0 || max_vertices < 1 || max_edges < 0) {
    g_create_error = "fb_create: bad argument";
    return nullptr;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("fb_create: no CUDA device (") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                     "); libflame_b200 has no CPU fallback";
    return nullptr;
  }
  if (device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) {
    g_create_error = "fb_create: bad device index";
    return nullptr;
  }
  fb_ctx* c = new fb_ctx();
  c->device = device; c->S = n_streams; c->W = width; c->H = height; c->n_slots = n_slots;
  c->maxF = max_features; c->maxV = max_vertices; c->maxE = max_edges;
  c->maxT = 2 * max_vertices;
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      g_create_error = "fb_create: cudaStreamCreate failed";
      delete c;
      return nullptr;
    }
    c->own_stream = true;
  }
  const size_t S = n_streams, nv = S * max_vertices, ne = S * max_edges, nf = S * max_features;
  const size_t npx = (size_t)width * height;
  bool ok = true;
  auto A = [&](cudaError_t r) { if (r != cudaSuccess) { ok = false; g_create_error = std::string("fb_create: ") + cudaGetErrorString(r); } };
  A(dalloc(&c->vbar, nv)); A(dalloc(&c->x, nv)); A(dalloc(&c->w1, nv)); A(dalloc(&c->w2, nv));
  A(dalloc(&c->z, nv)); A(dalloc(&c->wt, nv)); A(dalloc(&c->ec, ne)); A(dalloc(&c->eij, ne));
  A(dalloc(&c->q4, ne)); A(dalloc(&c->row, S * (max_vertices + 1))); A(dalloc(&c->inc, 2 * ne));
  A(dalloc(&c->epos, ne)); A(dalloc(&c->vnin, nv));
  A(dalloc(&c->nV, S)); A(dalloc(&c->nE, S)); A(dalloc(&c->vfeat, nv)); A(dalloc(&c->vpos, nv));
  A(dalloc(&c->costs, 2 * S));
  A(dalloc(&c->imgs, S * n_slots * npx)); A(dalloc(&c->d_pose, S * n_slots * 7));
  A(dalloc(&c->d_K, S * 9)); A(dalloc(&c->d_cmp, S));
  A(dalloc(&c->d_geo, S * n_slots * FB_GEO_STRIDE));
  A(dalloc(&c->f_uref, nf)); A(dalloc(&c->f_mu, nf)); A(dalloc(&c->f_var, nf));
  A(dalloc(&c->f_drop, nf)); A(dalloc(&c->f_alive, nf)); A(dalloc(&c->f_ref, nf));
  A(dalloc(&c->f_status, nf)); A(dalloc(&c->f_ucmp, nf)); A(dalloc(&c->nF, S));
  A(dalloc(&c->counters, S * FB_NUM_COUNTERS));
  A(dalloc(&c->tri, S * (size_t)c->maxT * 3)); A(dalloc(&c->nT, S));
  A(dalloc(&c->tri_valid, S * (size_t)c->maxT)); A(dalloc(&c->owner, S * npx));
  A(cudaMemset(c->owner, 0x7f, sizeof(int32_t) * S * npx));  // FB_OWNER_NONE; every shading pass leaves it that way
  A(dalloc(&c->idmap, S * npx));
  if (!ok) {
    free_all(c);
    delete c;
    return nullptr;
  }
  cudaMemsetAsync(c->nV, 0, sizeof(int32_t) * S, c->stream);
  cudaMemsetAsync(c->nE, 0, sizeof(int32_t) * S, c->stream);
  cudaMemsetAsync(c->nF, 0, sizeof(int32_t) * S, c->stream);
  cudaMemsetAsync(c->nT, 0, sizeof(int32_t) * S, c->stream);
  cudaMemsetAsync(c->counters, 0, sizeof(int32_t) * S * FB_NUM_COUNTERS, c->stream);
  cudaMemsetAsync(c->vfeat, 0xff, sizeof(int32_t) * nv, c->stream);
  cudaMemsetAsync(c->imgs, 0, S * n_slots * npx, c->stream);
  c->hV.assign(S, 0); c->hE.assign(S, 0); c->hF.assign(S, 0); c->hT.assign(S, 0);
  c->h_pose.assign(S * n_slots * 7, 0.f);
  for (size_t k = 0; k < S * n_slots; ++k) c->h_pose[7 * k + 3] = 1.f;
  c->h_K.assign(S * 9, 0.f);
  for (size_t s = 0; s < S; ++s) {
    float* K = &c->h_K[9 * s];
    K[0] = K[4] = (float)width; K[2] = 0.5f * (width - 1); K[5] = 0.5f * (height - 1); K[8] = 1.f;
  }
  cudaMemcpyAsync(c->d_K, c->h_K.data(), sizeof(float) * S * 9, cudaMemcpyHostToDevice, c->stream);
  fb_default_epi_params(&c->epi);
  if (const char* e = getenv("FB_CLUSTER_MIN")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 16) c->cluster_min = v;
  }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
    g_create_error = "fb_create: initialisation failed";
    free_all(c);
    delete c;
    return nullptr;
  }
  return c;
}

extern "C" void fb_destroy(fb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  c->pipe_dirty = true;
  pipeline_drain(c);
  cudaStreamSynchronize(c->stream);
  free_all(c);
  delete c;
}

extern "C" int fb_sync(fb_ctx* c) {
  CHECK_CTX_RO(c);
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (grid_watchdog_fired(c)) FB_FAIL(c, FB_E_STATE, "grid-resident solver: mailbox exchange timed out (watchdog)");
  if (fb_tile_failed(c)) {
    tile_clear_failure(c, false);
    FB_FAIL(c, FB_E_STATE, "tile-resident solver: a tile exceeded its capacity (nothing was solved)");
  }
  return FB_OK;
}

extern "C" int fb_set_intrinsics(fb_ctx* c, int s, const float K[9]) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (!K || !(K[0] > 0.f) || !(K[4] > 0.f)) FB_FAIL(c, FB_E_ARG, "fb_set_intrinsics: bad K");
  memcpy(&c->h_K[9 * s], K, sizeof(float) * 9);
  FB_CUDA(c, cudaMemcpyAsync(c->d_K + 9 * s, &c->h_K[9 * s], sizeof(float) * 9,
                             cudaMemcpyHostToDevice, c->stream));
  return FB_OK;
}

extern "C" int fb_set_epi_params(fb_ctx* c, const fb_epi_params* p) {
  CHECK_CTX(c);
  if (!p || p->win_size < 3 || p->win_size > FB_MAX_WIN || (p->win_size & 1) == 0 ||
      p->max_search_px < 8 || p->max_search_px > FB_MAX_SEARCH || p->ambiguity_radius < 0)
    FB_FAIL(c, FB_E_ARG, "fb_set_epi_params: win_size must be odd in [3,15], max_search_px in [8,256]");
  c->epi = *p;
  update_invalidate_graphs(c);  // the parameters are baked into the captured frame
  return FB_OK;
}

// ------------------------------------------------------------------------------------ graph
extern "C" int fb_graph_set(fb_ctx* c, int s, int V, int E, const float* pos,
                            const int32_t* ij, const float* alpha, const float* beta) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (V < 0 || E < 0 || (V > 0 && !pos) || (E > 0 && (!ij || !alpha || !beta)))
    FB_FAIL(c, FB_E_ARG, "fb_graph_set: null input");
  if (V > c->maxV || E > c->maxE) FB_FAIL(c, FB_E_NOMEM, "fb_graph_set: V/E exceed context capacity");
  std::vector<float4> ec(E);
  std::vector<int2> eij(E);
  std::vector<int32_t> row(V + 1, 0), inc(2 * (size_t)E);
  for (int e = 0; e < E; ++e) {
    const int i = ij[2 * e], j = ij[2 * e + 1];
    if (i < 0 || j < 0 || i >= V || j >= V || i >= j)
      FB_FAIL(c, FB_E_ARG, "fb_graph_set: edges must satisfy 0 <= i < j < V (canonical orientation)");
    if (e > 0 && (ij[2 * e - 2] > i || (ij[2 * e - 2] == i && ij[2 * e - 1] >= j)))
      FB_FAIL(c, FB_E_ARG, "fb_graph_set: edges must be sorted by (i,j) without duplicates");
    // dx = pos_i - pos_j in fp32, once, so every kernel sees the same delta
    ec[e] = make_float4(alpha[e], beta[e], pos[2 * i] - pos[2 * j], pos[2 * i + 1] - pos[2 * j + 1]);
    eij[e] = make_int2(i, j);
    row[i + 1]++;
    row[j + 1]++;
  }
  for (int v = 0; v < V; ++v) row[v + 1] += row[v];
  {
    std::vector<int32_t> fill(row.begin(), row.begin() + V);
    for (int e = 0; e < E; ++e) {
      inc[fill[eij[e].x]++] = (e << 1);
      inc[fill[eij[e].y]++] = (e << 1) | 1;
    }
  }
  // position of every edge in its target's row and the in-degrees (in-edges come first in a row:
  // edges are sorted by (i,j) with i < j), used by the device-planned tile solver
  std::vector<int32_t> epos(E), vnin(V, 0);
  for (int v = 0; v < V; ++v)
    for (int k = row[v]; k < row[v + 1] && (inc[k] & 1); ++k) {
      epos[inc[k] >> 1] = k - row[v];
      vnin[v]++;
    }
  const size_t vb = (size_t)s * c->maxV, eb = (size_t)s * c->maxE;
  cudaStream_t st = c->stream;
  if (E) {
    FB_CUDA(c, cudaMemcpyAsync(c->ec + eb, ec.data(), sizeof(float4) * E, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(c->eij + eb, eij.data(), sizeof(int2) * E, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(c->inc + 2 * eb, inc.data(), sizeof(int32_t) * 2 * E, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(c->epos + eb, epos.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemsetAsync(c->q4 + eb, 0, sizeof(float4) * E, st));
  }
  FB_CUDA(c, cudaMemcpyAsync(c->row + (size_t)s * (c->maxV + 1), row.data(), sizeof(int32_t) * (V + 1), cudaMemcpyHostToDevice, st));
  if (V) {
    FB_CUDA(c, cudaMemcpyAsync(c->vpos + vb, pos, sizeof(float2) * V, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(c->vnin + vb, vnin.data(), sizeof(int32_t) * V, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemsetAsync(c->vbar + vb, 0, sizeof(float4) * V, st));
    FB_CUDA(c, cudaMemsetAsync(c->x + vb, 0, sizeof(float) * V, st));
    FB_CUDA(c, cudaMemsetAsync(c->w1 + vb, 0, sizeof(float) * V, st));
    FB_CUDA(c, cudaMemsetAsync(c->w2 + vb, 0, sizeof(float) * V, st));
    FB_CUDA(c, cudaMemsetAsync(c->z + vb, 0, sizeof(float) * V, st));
    std::vector<float> ones(V, 1.0f);
    FB_CUDA(c, cudaMemcpyAsync(c->wt + vb, ones.data(), sizeof(float) * V, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemsetAsync(c->vfeat + vb, 0xff, sizeof(int32_t) * V, st));
  }
  c->hV[s] = V;
  c->hE[s] = E;
  update_mark_host_graph(c, s);
  tile_plan_mark(c, s);
  FB_CUDA(c, cudaMemcpyAsync(c->nV + s, &c->hV[s], sizeof(int32_t), cudaMemcpyHostToDevice, st));
  FB_CUDA(c, cudaMemcpyAsync(c->nE + s, &c->hE[s], sizeof(int32_t), cudaMemcpyHostToDevice, st));
  int rc = cluster_plan_build(c, s, V, E, eij.data(), row.data(), inc.data());
  if (rc != FB_OK) return rc;
  rc = grid_plan_set(c, s, V, pos);
  if (rc != FB_OK) return rc;
  // the staging vectors above are pageable: the runtime has copied them before returning
  FB_CUDA(c, cudaStreamSynchronize(st));
  return FB_OK;
}

extern "C" int fb_graph_data_set(fb_ctx* c, int s, const float* z, const float* wt) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  const int V = c->hV[s];
  if (!z) FB_FAIL(c, FB_E_ARG, "fb_graph_data_set: null z");
  const size_t vb = (size_t)s * c->maxV;
  FB_CUDA(c, cudaMemcpyAsync(c->z + vb, z, sizeof(float) * V, cudaMemcpyHostToDevice, c->stream));
  if (wt) {
    FB_CUDA(c, cudaMemcpyAsync(c->wt + vb, wt, sizeof(float) * V, cudaMemcpyHostToDevice, c->stream));
  } else {
    std::vector<float> ones(V, 1.0f);
    FB_CUDA(c, cudaMemcpyAsync(c->wt + vb, ones.data(), sizeof(float) * V, cudaMemcpyHostToDevice, c->stream));
  }
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FB_OK;
}

extern "C" int fb_graph_state_set(fb_ctx* c, int s, const float* x, const float* w, const float* q) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  const int V = c->hV[s], E = c->hE[s];
  const size_t vb = (size_t)s * c->maxV, eb = (size_t)s * c->maxE;
  cudaStream_t st = c->stream;
  std::vector<float> w1, w2;
  std::vector<float4> q4;
  if (x) FB_CUDA(c, cudaMemcpyAsync(c->x + vb, x, sizeof(float) * V, cudaMemcpyHostToDevice, st));
  if (w) {
    w1.resize(V); w2.resize(V);
    for (int v = 0; v < V; ++v) { w1[v] = w[2 * v]; w2[v] = w[2 * v + 1]; }
    FB_CUDA(c, cudaMemcpyAsync(c->w1 + vb, w1.data(), sizeof(float) * V, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(c->w2 + vb, w2.data(), sizeof(float) * V, cudaMemcpyHostToDevice, st));
  }
  if (q) {
    q4.resize(E);
    for (int e = 0; e < E; ++e) q4[e] = make_float4(q[3 * e], q[3 * e + 1], q[3 * e + 2], 0.f);
    FB_CUDA(c, cudaMemcpyAsync(c->q4 + eb, q4.data(), sizeof(float4) * E, cudaMemcpyHostToDevice, st));
  }
  const int n = std::max(V, E);
  if (n > 0) {
    k_state_init<<<fb_div_up(n, 256), 256, 0, st>>>(graph_view(c), s, V, E, x ? 0 : 1, w ? 0 : 1, q ? 0 : 1);
    c->launches++;
    FB_CUDA(c, cudaGetLastError());
  }
  FB_CUDA(c, cudaStreamSynchronize(st));
  return FB_OK;
}

extern "C" int fb_graph_state_get(fb_ctx* c, int s, float* x, float* w, float* q, float* xbar) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  const int V = c->hV[s], E = c->hE[s];
  const size_t vb = (size_t)s * c->maxV, eb = (size_t)s * c->maxE;
  cudaStream_t st = c->stream;
  std::vector<float> w1(V), w2(V);
  std::vector<float4> q4(E), vb4(V);
  if (x) FB_CUDA(c, cudaMemcpyAsync(x, c->x + vb, sizeof(float) * V, cudaMemcpyDeviceToHost, st));
  if (w) {
    FB_CUDA(c, cudaMemcpyAsync(w1.data(), c->w1 + vb, sizeof(float) * V, cudaMemcpyDeviceToHost, st));
    FB_CUDA(c, cudaMemcpyAsync(w2.data(), c->w2 + vb, sizeof(float) * V, cudaMemcpyDeviceToHost, st));
  }
  if (q) FB_CUDA(c, cudaMemcpyAsync(q4.data(), c->q4 + eb, sizeof(float4) * E, cudaMemcpyDeviceToHost, st));
  if (xbar) FB_CUDA(c, cudaMemcpyAsync(vb4.data(), c->vbar + vb, sizeof(float4) * V, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  if (grid_watchdog_fired(c)) FB_FAIL(c, FB_E_STATE, "grid-resident solver: mailbox exchange timed out (watchdog)");
  if (fb_tile_failed(c)) {
    tile_clear_failure(c, false);
    FB_FAIL(c, FB_E_STATE, "tile-resident solver: a tile exceeded its capacity (nothing was solved)");
  }
  if (w) for (int v = 0; v < V; ++v) { w[2 * v] = w1[v]; w[2 * v + 1] = w2[v]; }
  if (q) for (int e = 0; e < E; ++e) { q[3 * e] = q4[e].x; q[3 * e + 1] = q4[e].y; q[3 * e + 2] = q4[e].z; }
  if (xbar) for (int v = 0; v < V; ++v) { xbar[3 * v] = vb4[v].x; xbar[3 * v + 1] = vb4[v].y; xbar[3 * v + 2] = vb4[v].z; }
  return FB_OK;
}

extern "C" int fb_graph_x_get_all(fb_ctx* c, float* x_all) {
  CHECK_CTX(c);
  if (!x_all) FB_FAIL(c, FB_E_ARG, "fb_graph_x_get_all: null output");
  FB_CUDA(c, cudaMemcpyAsync(x_all, c->x, sizeof(float) * (size_t)c->S * c->maxV, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FB_OK;
}

static int solve_streaming(fb_ctx* c, int iters, const fb_nltgv2_params* p, int only = -1) {
  // The launch sequence is captured once per (only, iters, params) and replayed as one graph; grid
  // extents cover the context capacity and counts are read on the device, so a captured graph
  // survives topology changes.
  static_assert(sizeof(fb_nltgv2_params) == 24, "params layout");
  if (c->solve_exec.empty()) {
    c->solve_exec.assign(2 * (c->S + 1), nullptr);
    c->solve_iters.assign(2 * (c->S + 1), 0);
    c->solve_params.assign(2 * (c->S + 1), fb_nltgv2_params{});
    c->solve_z.assign(2 * (c->S + 1), nullptr);
  }
  // the captured launches hold the data-term pointers: one cached graph per buffer of the pipelined step
  const int slot = 2 * (only + 1) + ((c->z_buf[1] && c->z == c->z_buf[1]) ? 1 : 0);
  const bool reuse = c->solve_exec[slot] && c->solve_iters[slot] == iters && c->solve_z[slot] == c->z &&
                     memcmp(&c->solve_params[slot], p, sizeof(*p)) == 0;
  if (!reuse) {
    if (c->solve_exec[slot]) { cudaGraphExecDestroy(c->solve_exec[slot]); c->solve_exec[slot] = nullptr; }
    GraphView g = graph_view(c);
    g.only = only;
    const dim3 ge(fb_div_up(std::max(c->maxE, 1), 256), c->S), gv(fb_div_up(c->maxV, 256), c->S);
    const float tl = p->step_x * p->data_factor;
    cudaGraph_t graph = nullptr;
    FB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    for (int it = 0; it < iters; ++it) {
      k_dual_edges<<<ge, 256, 0, c->stream>>>(g, p->step_q);
      k_primal_vertices<<<gv, 256, 0, c->stream>>>(g, p->step_x, tl, p->theta, p->x_min, p->x_max);
    }
    FB_CUDA(c, cudaStreamEndCapture(c->stream, &graph));
    cudaError_t e = cudaGraphInstantiate(&c->solve_exec[slot], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { c->solve_exec[slot] = nullptr; FB_FAIL(c, FB_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    c->solve_iters[slot] = iters;
    c->solve_params[slot] = *p;
    c->solve_z[slot] = c->z;
  }
  ProfScope ps(c, FB_PROF_SOLVE);
  FB_CUDA(c, cudaGraphLaunch(c->solve_exec[slot], c->stream));
  c->launches += 2 * (int64_t)iters;
  return FB_OK;
}

// ---- variant 5: tile-resident solver planned on the device (nltgv2_tile.cuh) ----------------------
static void tile_plan_free(fb_ctx* c) {
  TilePlan* T = c->tplan;
  if (!T) return;
  cudaFree(T->vtile); cudaFree(T->vloc); cudaFree(T->tlist); cudaFree(T->toff); cudaFree(T->lrow); cudaFree(T->derr);
  if (T->err) cudaFreeHost(T->err);
  delete T;
  c->tplan = nullptr;
}
static void tile_plan_mark(fb_ctx* c, int s) {
  if (c->tplan && s < (int)c->tplan->dirty.size()) c->tplan->dirty[s] = 1;
}
static bool fb_tile_failed(const fb_ctx* c) { return c->tplan && c->tplan->err && *c->tplan->err != 0; }
// A capacity verdict is reported ONCE (the launch that raised it solved nothing); the flags are then
// cleared so the context stays usable.  disable: fb_update stops choosing this solver for the context
// (its frames fall back to the plan-free / streaming kernels) and drops the captured frames that hold it.
static void tile_clear_failure(fb_ctx* c, bool disable) {
  if (!c->tplan) return;
  if (c->tplan->err) *c->tplan->err = 0;
  if (c->tplan->derr) cudaMemsetAsync(c->tplan->derr, 0, sizeof(int), c->stream);
  if (disable) {
    c->tplan->available = 0;
    update_invalidate_graphs(c);
  }
}
static bool tile_available(fb_ctx* c);
// Allocates the plan and probes the launch configuration (cluster of 16, ~148 KB shared memory).
static bool tile_available(fb_ctx* c) {
  if (c->tplan && c->tplan->available >= 0) return c->tplan->available == 1;
  if (!c->tplan) c->tplan = new TilePlan();
  TilePlan* T = c->tplan;
  T->available = 0;
  if (const char* e = getenv("FB_TILE_DISABLE")) if (atoi(e)) return false;
  // tiles hold V/16 vertices give or take the granularity of the split: keep 35 % head room
  if ((long long)c->maxV * 100 > (long long)FBT_C * FBT_VCAP * 65) return false;
  const size_t S = c->S, nv = S * c->maxV;
  if (dalloc(&T->vtile, nv) != cudaSuccess || dalloc(&T->vloc, nv) != cudaSuccess || dalloc(&T->tlist, nv) != cudaSuccess ||
      dalloc(&T->toff, S * (FBT_C + 1)) != cudaSuccess || dalloc(&T->lrow, nv) != cudaSuccess ||
      dalloc(&T->derr, 1) != cudaSuccess || cudaMemset(T->derr, 0, sizeof(int)) != cudaSuccess ||
      cudaHostAlloc((void**)&T->err, sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  *T->err = 0;
  T->dirty.assign(c->S, 1);
  T->smem = fbt_smem_bytes();
  const void* kern = (const void*)k_nltgv2_tile;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T->smem) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t q{};
  q.gridDim = dim3(FBT_C);
  q.blockDim = dim3(FBT_THREADS);
  q.dynamicSmemBytes = T->smem;
  cudaLaunchAttribute qa[1];
  qa[0].id = cudaLaunchAttributeClusterDimension;
  qa[0].val.clusterDim.x = FBT_C; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
  q.attrs = qa;
  q.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess) { cudaGetLastError(); n = 0; }
  T->available = n >= 1 ? 1 : 0;
  return T->available == 1;
}

// The tiles of one stream cut ahead of the solve, on a side stream (fb_update: beside the star kernel).
// A no-op when the tile-resident solver is not the one fb_update will pick.
static void tile_assign_early(fb_ctx* c, int s, cudaStream_t st) {
  if (!tile_available(c)) return;
  TilePlan* T = c->tplan;
  if (!T->dirty[s]) return;
  k_tile_assign<<<1, 1024, 0, st>>>(s, c->maxV, c->nV, c->vpos, T->vtile, T->vloc, T->tlist, T->toff);
  c->launches++;
  T->dirty[s] = 0;
}

static int solve_tile(fb_ctx* c, int iters, const fb_nltgv2_params* p, int only = -1) {
  if (!tile_available(c)) FB_FAIL(c, FB_E_STATE, "fb_nltgv2_solve: the tile-resident solver is not available for this context");
  TilePlan* T = c->tplan;
  for (int s = 0; s < c->S; ++s) {
    if ((only >= 0 && s != only) || !T->dirty[s]) continue;
    k_tile_assign<<<1, 1024, 0, c->stream>>>(s, c->maxV, c->nV, c->vpos, T->vtile, T->vloc, T->tlist, T->toff);
    c->launches++;
    T->dirty[s] = 0;
  }
  TileArgs a;
  a.g = graph_view(c);
  a.g.only = only;
  a.vtile = T->vtile; a.vloc = T->vloc; a.tlist = T->tlist; a.toff = T->toff; a.lrow = T->lrow; a.err = T->err; a.derr = T->derr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((only >= 0 ? 1 : c->S) * FBT_C));
  cfg.blockDim = dim3(FBT_THREADS);
  cfg.dynamicSmemBytes = T->smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FBT_C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  float sq = p->step_q, sx = p->step_x, tl = p->step_x * p->data_factor, th = p->theta, x0 = p->x_min, x1 = p->x_max;
  int itv = iters;
  void* args[] = {&a, &itv, &sq, &sx, &tl, &th, &x0, &x1};
  ProfScope ps(c, FB_PROF_SOLVE);
  FB_CUDA(c, cudaLaunchKernelExC(&cfg, (const void*)k_nltgv2_tile, args));
  c->launches++;
  c->last_cluster = FBT_C;
  return FB_OK;
}

// ---- variant 4: plan-free resident solver (nltgv2_coop.cuh) ---------------------------------------
static const void* coop_kernel(int ept) {
  return ept <= 3 ? (const void*)k_nltgv2_coop<3, 1> : (const void*)k_nltgv2_coop<6, 2>;
}
// Cluster size and instantiation for this context's capacities; coop_cluster = 0 when unavailable.
static void coop_probe(fb_ctx* c) {
  if (c->coop_cluster >= 0) return;
  c->coop_cluster = 0;
  const int sizes[2] = {16, 8};
  for (int k = 0; k < 2 && !c->coop_cluster; ++k) {
    const int C = sizes[k], NT = C * FBK_THREADS;
    int ept = 0, vpt = 0;
    if (c->maxE <= 3 * NT && c->maxV <= NT) { ept = 3; vpt = 1; }
    else if (c->maxE <= 6 * NT && c->maxV <= 2 * NT) { ept = 6; vpt = 2; }
    else continue;
    const void* kern = coop_kernel(ept);
    if (C > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    cudaLaunchConfig_t q{};
    q.gridDim = dim3((unsigned)C);
    q.blockDim = dim3(FBK_THREADS);
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = (unsigned)C; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (n >= 1) { c->coop_cluster = C; c->coop_ept = ept; c->coop_vpt = vpt; }
  }
  if (const char* e = getenv("FB_COOP_DISABLE")) if (atoi(e)) c->coop_cluster = 0;
}
static bool fb_coop_failed(const fb_ctx* c) { return c->coop_err && *c->coop_err != 0; }

static int solve_coop(fb_ctx* c, int iters, const fb_nltgv2_params* p, int only = -1) {
  coop_probe(c);
  if (!c->coop_cluster) FB_FAIL(c, FB_E_STATE, "fb_nltgv2_solve: the plan-free resident solver does not fit this context");
  if (!c->coop_contrib) {
    FB_CUDA(c, dalloc(&c->coop_contrib, (size_t)c->S * 2 * c->maxE));
    FB_CUDA(c, cudaHostAlloc((void**)&c->coop_err, sizeof(int), cudaHostAllocMapped));
    *c->coop_err = 0;
  }
  GraphView g = graph_view(c);
  g.only = only;
  const int C = c->coop_cluster;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((only >= 0 ? 1 : c->S) * C));
  cfg.blockDim = dim3(FBK_THREADS);
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  float sq = p->step_q, sx = p->step_x, tl = p->step_x * p->data_factor, th = p->theta, x0 = p->x_min, x1 = p->x_max;
  int itv = iters;
  float4* cb = c->coop_contrib;
  int* er = c->coop_err;
  void* args[] = {&g, &cb, &er, &itv, &sq, &sx, &tl, &th, &x0, &x1};
  ProfScope ps(c, FB_PROF_SOLVE);
  FB_CUDA(c, cudaLaunchKernelExC(&cfg, coop_kernel(c->coop_ept), args));
  c->launches++;
  c->last_cluster = C;
  return FB_OK;
}

// True when a stream's current graph was built on the device (fb_update, delaunay_gpu.cuh): the
// per-topology tables of variants 2 / 3 then do not exist for it.
static bool any_device_graph(const fb_ctx* c);

extern "C" int fb_nltgv2_solve(fb_ctx* c, int iters, const fb_nltgv2_params* p, int variant) {
  CHECK_CTX(c);
  if (!p || iters < 0 || variant < 0 || variant > 5) FB_FAIL(c, FB_E_ARG, "fb_nltgv2_solve: bad argument");
  if (iters == 0) return FB_OK;
  int v = variant;
  const bool devg = any_device_graph(c);
  if (devg && (v == 2 || v == 3)) FB_FAIL(c, FB_E_STATE, "fb_nltgv2_solve: variants 2 / 3 need a host-set topology (fb_graph_set)");
  if (v == 5 || (v == 0 && devg)) {
    if (tile_available(c)) {
      c->last_variant = 5;
      const int rc = solve_tile(c, iters, p);
      if (rc == FB_OK || v == 5) return rc;
      cudaGetLastError();
      c->tplan->available = 0;
    } else if (v == 5) {
      FB_FAIL(c, FB_E_STATE, "fb_nltgv2_solve: the tile-resident solver is not available for this context");
    }
  }
  if (v == 4 || (v == 0 && devg)) {
    coop_probe(c);
    if (c->coop_cluster) {
      c->last_variant = 4;
      const int rc = solve_coop(c, iters, p);
      if (rc == FB_OK || v == 4) return rc;
      cudaGetLastError();
      c->coop_cluster = 0;
    } else if (v == 4) {
      FB_FAIL(c, FB_E_STATE, "fb_nltgv2_solve: the plan-free resident solver does not fit this context");
    }
    c->last_variant = 1;
    return solve_streaming(c, iters, p);
  }
  if ((v == 0 && !c->grid_disabled) || v == 3) {
    int nper = 0;
    size_t smem = 0;
    {
      int rc = grid_prepare(c, iters, -1, &nper, &smem);
      if (rc) return rc;
    }
    if (nper > 0) {
      c->last_variant = 3;
      const int rc = solve_grid(c, iters, p, nper, smem);
      if (rc == FB_OK || v == 3) return rc;
      // auto: a launch configuration this device / driver refuses (cluster size, cooperative launch)
      // is an error of the launch call itself -- nothing ran; remember it and use the other variants
      cudaGetLastError();
      c->grid_disabled = true;
    } else if (v == 3) {
      FB_FAIL(c, FB_E_STATE, "fb_nltgv2_solve: the batch does not fit the grid-resident solver");
    }
  }
  if (v == 0) v = cluster_plan_ready(c) ? 2 : 1;
  if (v == 2 && !cluster_plan_ready(c))
    FB_FAIL(c, FB_E_STATE, "fb_nltgv2_solve: a graph of this context does not fit the cluster-resident solver");
  c->last_variant = v;
  if (v == 2) return solve_cluster(c, iters, p);
  return solve_streaming(c, iters, p);
}

// One stream of a batch (fb_update drives streams independently).  fb_update re-triangulates every
// frame (~2 % of the edges change), so per-topology tables are never reusable: the plan-free resident
// kernel (variant 4) takes the solve, the streaming kernels (graph replay) when it does not fit.
static int fb_nltgv2_solve_stream(fb_ctx* c, int s, int iters, const fb_nltgv2_params* p) {
  if (iters <= 0) return FB_OK;
  const int only = c->S > 1 ? s : -1;
  if (tile_available(c)) {
    c->last_variant = 5;
    if (solve_tile(c, iters, p, only) == FB_OK) return FB_OK;
    cudaGetLastError();  // launch refused: nothing ran
    c->tplan->available = 0;
  }
  coop_probe(c);
  if (c->coop_cluster) {
    c->last_variant = 4;
    if (solve_coop(c, iters, p, only) == FB_OK) return FB_OK;
    cudaGetLastError();  // launch refused: nothing ran
    c->coop_cluster = 0;
  }
  c->last_variant = 1;
  return solve_streaming(c, iters, p, only);
}

// Context-free, host-only: builds the variant-3 partition tables of one graph for `parts` CTAs and
// checks the invariants k_nltgv2_grid relies on (see fbg_verify).  Returns 0 when they hold.
extern "C" int fb_grid_plan_verify(int V, int E, const float* pos, const int32_t* ij, int parts, int cluster,
                                   int32_t* stats) {
  if (V < 0 || E < 0 || parts < 1 || parts > (cluster ? FBG_MAXC : FBG_MAXP) || (V > 0 && !pos) || (E > 0 && !ij)) return FB_E_ARG;
  for (int e = 0; e < E; ++e)
    if (ij[2 * e] < 0 || ij[2 * e] >= ij[2 * e + 1] || ij[2 * e + 1] >= V) return FB_E_ARG;
  std::string why;
  const int rc = fbg_verify(V, E, pos, ij, parts, cluster ? FBG_THREADS_CL : FBG_THREADS_L2,
                            cluster ? FBG_SMEM_LIMIT_CL : FBG_SMEM_LIMIT_L2, stats, why);
  if (rc) g_create_error = "fb_grid_plan_verify: " + why;
  return rc;
}

extern "C" int fb_last_solver_variant(const fb_ctx* c) { return c ? c->last_variant : 0; }
extern "C" int fb_last_cluster_size(const fb_ctx* c) { return c ? c->last_cluster : 0; }
extern "C" int fb_last_solver_transport(const fb_ctx* c) { return c ? c->last_transport : 0; }

extern "C" int fb_costs(fb_ctx* c, int s, float data_factor, double* smooth, double* data) {
  CHECK_CTX_RO(c);
  CHECK_STREAM(c, s);
  FB_CUDA(c, cudaMemsetAsync(c->costs, 0, sizeof(double) * 2 * c->S, c->stream));
  const dim3 grid(std::max(1, std::min(64, fb_div_up(std::max(c->maxE, c->maxV), 256))), c->S);
  k_costs<<<grid, 256, 0, c->stream>>>(graph_view(c), data_factor, c->costs);
  c->launches++;
  FB_CUDA(c, cudaGetLastError());
  std::vector<double> h(2 * c->S);
  FB_CUDA(c, cudaMemcpyAsync(h.data(), c->costs, sizeof(double) * 2 * c->S, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (smooth) *smooth = h[2 * s];
  if (data) *data = h[2 * s + 1];
  return FB_OK;
}

// ------------------------------------------------------------------------------------ frames
static int check_slot(fb_ctx* c, int s, int slot) {
  CHECK_STREAM(c, s);
  if (slot < 0 || slot >= c->n_slots) FB_FAIL(c, FB_E_ARG, "slot index out of range");
  return FB_OK;
}

extern "C" int fb_frame_pose_set(fb_ctx* c, int s, int slot, const float pose[7]) {
  CHECK_CTX(c);
  int rc = check_slot(c, s, slot);
  if (rc) return rc;
  if (!pose) FB_FAIL(c, FB_E_ARG, "null pose");
  memcpy(&c->h_pose[((size_t)s * c->n_slots + slot) * 7], pose, sizeof(float) * 7);
  return FB_OK;
}

extern "C" int fb_frame_set(fb_ctx* c, int s, int slot, const uint8_t* gray, int pitch,
                            const float pose[7]) {
  CHECK_CTX(c);
  int rc = check_slot(c, s, slot);
  if (rc) return rc;
  if (!gray || pitch < c->W) FB_FAIL(c, FB_E_ARG, "fb_frame_set: null image or pitch < width");
  rc = fb_frame_pose_set(c, s, slot, pose);
  if (rc) return rc;
  ProfScope ps(c, FB_PROF_UPLOAD);
  uint8_t* dst = c->imgs + ((size_t)s * c->n_slots + slot) * (size_t)c->W * c->H;
  if (pitch == c->W) {
    FB_CUDA(c, cudaMemcpyAsync(dst, gray, (size_t)c->W * c->H, cudaMemcpyHostToDevice, c->stream));
  } else {
    FB_CUDA(c, cudaMemcpy2DAsync(dst, c->W, gray, pitch, c->W, c->H, cudaMemcpyHostToDevice, c->stream));
  }
  return FB_OK;
}

extern "C" int fb_pool_reserve(fb_ctx* c, int n) {
  CHECK_CTX(c);
  if (n < 0) FB_FAIL(c, FB_E_ARG, "fb_pool_reserve: negative size");
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->pool);
  c->pool = nullptr;
  c->pool_n = 0;
  if (n) {
    FB_CUDA(c, dalloc(&c->pool, (size_t)n * c->W * c->H));
    c->pool_n = n;
  }
  return FB_OK;
}

extern "C" int fb_pool_upload(fb_ctx* c, int idx, const uint8_t* gray, int pitch) {
  CHECK_CTX(c);
  if (idx < 0 || idx >= c->pool_n || !gray || pitch < c->W) FB_FAIL(c, FB_E_ARG, "fb_pool_upload: bad argument");
  FB_CUDA(c, cudaMemcpy2DAsync(c->pool + (size_t)idx * c->W * c->H, c->W, gray, pitch, c->W, c->H, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FB_OK;
}

extern "C" int fb_frame_from_pool(fb_ctx* c, int s, int slot, int idx, const float pose[7]) {
  CHECK_CTX(c);
  int rc = check_slot(c, s, slot);
  if (rc) return rc;
  if (idx < 0 || idx >= c->pool_n) FB_FAIL(c, FB_E_ARG, "fb_frame_from_pool: bad pool index");
  rc = fb_frame_pose_set(c, s, slot, pose);
  if (rc) return rc;
  const size_t fsz = (size_t)c->W * c->H;
  FB_CUDA(c, cudaMemcpyAsync(c->imgs + ((size_t)s * c->n_slots + slot) * fsz, c->pool + (size_t)idx * fsz, fsz, cudaMemcpyDeviceToDevice, c->stream));
  return FB_OK;
}

// ------------------------------------------------------------------------------------ features
extern "C" int fb_features_set(fb_ctx* c, int s, int N, const float* u_ref, const int32_t* ref_slot,
                               const float* mu, const float* var, const int32_t* dropouts,
                               const int32_t* alive) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (N < 0 || N > c->maxF) FB_FAIL(c, FB_E_NOMEM, "fb_features_set: N exceeds context capacity");
  if (N > 0 && (!u_ref || !ref_slot || !mu || !var)) FB_FAIL(c, FB_E_ARG, "fb_features_set: null input");
  for (int f = 0; f < N; ++f)
    if (ref_slot[f] < 0 || ref_slot[f] >= c->n_slots) FB_FAIL(c, FB_E_ARG, "fb_features_set: ref_slot out of range");
  const size_t fb = (size_t)s * c->maxF;
  cudaStream_t st = c->stream;
  std::vector<int32_t> ones;
  if (N) {
    FB_CUDA(c, cudaMemcpyAsync(c->f_uref + fb, u_ref, sizeof(float2) * N, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(c->f_ref + fb, ref_slot, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(c->f_mu + fb, mu, sizeof(float) * N, cudaMemcpyHostToDevice, st));
    FB_CUDA(c, cudaMemcpyAsync(c->f_var + fb, var, sizeof(float) * N, cudaMemcpyHostToDevice, st));
    if (dropouts) FB_CUDA(c, cudaMemcpyAsync(c->f_drop + fb, dropouts, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    else FB_CUDA(c, cudaMemsetAsync(c->f_drop + fb, 0, sizeof(int32_t) * N, st));
    if (alive) {
      FB_CUDA(c, cudaMemcpyAsync(c->f_alive + fb, alive, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    } else {
      ones.assign(N, 1);
      FB_CUDA(c, cudaMemcpyAsync(c->f_alive + fb, ones.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    }
    FB_CUDA(c, cudaMemsetAsync(c->f_status + fb, 0, sizeof(int32_t) * N, st));
    FB_CUDA(c, cudaMemsetAsync(c->f_ucmp + fb, 0xff, sizeof(float2) * N, st));
  }
  c->hF[s] = N;
  FB_CUDA(c, cudaMemcpyAsync(c->nF + s, &c->hF[s], sizeof(int32_t), cudaMemcpyHostToDevice, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  return FB_OK;
}

extern "C" int fb_features_get(fb_ctx* c, int s, float* mu, float* var, int32_t* dropouts,
                               int32_t* alive, int32_t* status, float* u_cmp) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  const int N = c->hF[s];
  const size_t fb = (size_t)s * c->maxF;
  cudaStream_t st = c->stream;
  if (N) {
    if (mu) FB_CUDA(c, cudaMemcpyAsync(mu, c->f_mu + fb, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
    if (var) FB_CUDA(c, cudaMemcpyAsync(var, c->f_var + fb, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
    if (dropouts) FB_CUDA(c, cudaMemcpyAsync(dropouts, c->f_drop + fb, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
    if (alive) FB_CUDA(c, cudaMemcpyAsync(alive, c->f_alive + fb, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
    if (status) FB_CUDA(c, cudaMemcpyAsync(status, c->f_status + fb, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
    if (u_cmp) FB_CUDA(c, cudaMemcpyAsync(u_cmp, c->f_ucmp + fb, sizeof(float2) * N, cudaMemcpyDeviceToHost, st));
  }
  FB_CUDA(c, cudaStreamSynchronize(st));
  return FB_OK;
}

#define FB_STAGE_SLOTS 32
// Next slot of the pinned staging ring (deep enough that a slot is never rewritten while a copy
// from it is still queued: callers block on results at most a few frames behind).
static uint8_t* stage_slot(fb_ctx* c) {
  if (!c->stage) {
    c->stage_slot_bytes = (sizeof(float) * 7 * (size_t)c->S * c->n_slots + sizeof(int32_t) * c->S + 255) & ~(size_t)255;
    if (cudaMallocHost((void**)&c->stage, c->stage_slot_bytes * FB_STAGE_SLOTS) != cudaSuccess) return nullptr;
    c->stage_ev.resize(FB_STAGE_SLOTS);
    for (auto& e : c->stage_ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  // a caller that enqueues far ahead of the GPU waits here until the copy last issued from this
  // slot has executed, so a slot is never rewritten under a queued copy
  cudaEventSynchronize(c->stage_ev[c->stage_next]);
  uint8_t* p = c->stage + c->stage_slot_bytes * (size_t)c->stage_next;
  c->stage_last = c->stage_next;
  c->stage_next = (c->stage_next + 1) % FB_STAGE_SLOTS;
  return p;
}
// Call after the copies from the slot returned by stage_slot() have been enqueued on c->stream.
static void stage_commit(fb_ctx* c) { cudaEventRecord(c->stage_ev[c->stage_last], c->stream); }

static int upload_geometry(fb_ctx* c, const int32_t* cmp_slot, bool zero_counters = false) {
  const size_t np = (size_t)c->S * c->n_slots * 7;
  if (np <= FB_GEO_REC_FLOATS && c->S <= FB_GEO_REC_STREAMS && !getenv("FB_GEO_PINNED")) {
    // poses and slots as kernel parameters (see k_epi_geometry_rec)
    GeoRecord rec;
    memcpy(rec.poses, c->h_pose.data(), sizeof(float) * np);
    memcpy(rec.cmp, cmp_slot, sizeof(int32_t) * c->S);
    k_epi_geometry_rec<<<c->S, std::max(32, c->n_slots), 0, c->stream>>>(rec, c->d_K, c->n_slots, c->d_geo, 0, c->d_pose, c->d_cmp,
                                                                         zero_counters ? c->counters : nullptr);
    c->launches++;
    FB_CUDA(c, cudaGetLastError());
    return FB_OK;
  }
  uint8_t* st = stage_slot(c);
  if (!st) FB_FAIL(c, FB_E_NOMEM, "pinned staging allocation failed");
  memcpy(st, c->h_pose.data(), sizeof(float) * np);
  memcpy(st + sizeof(float) * np, cmp_slot, sizeof(int32_t) * c->S);
  // larger batches: the kernel reads the staged poses / slots from pinned host memory itself and
  // publishes the device copies: no small copies on the host-to-device engine in front of the kernels
  k_epi_geometry<<<c->S, std::max(32, c->n_slots), 0, c->stream>>>(reinterpret_cast<const float*>(st), c->d_K,
                                                                   reinterpret_cast<const int32_t*>(st + sizeof(float) * np),
                                                                   c->n_slots, c->d_geo, 0, c->d_pose, c->d_cmp, zero_counters ? c->counters : nullptr);
  stage_commit(c);
  c->launches++;
  FB_CUDA(c, cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_features_reinit(fb_ctx* c, const int32_t* ref_slot, float mu0, float var0) {
  CHECK_CTX(c);
  if (!ref_slot) FB_FAIL(c, FB_E_ARG, "fb_features_reinit: null ref_slot");
  for (int s = 0; s < c->S; ++s)
    if (ref_slot[s] >= c->n_slots) FB_FAIL(c, FB_E_ARG, "fb_features_reinit: ref_slot out of range");
  if (c->maxF == 0) return FB_OK;
  ProfScope ps(c, FB_PROF_ASSEMBLY);
  const dim3 grid_rec(fb_div_up(c->maxF, 256), c->S);
  if (c->S <= FB_GEO_REC_STREAMS && !getenv("FB_GEO_PINNED")) {
    SlotRecord rec;
    memcpy(rec.v, ref_slot, sizeof(int32_t) * c->S);
    k_features_reinit_rec<<<grid_rec, 256, 0, c->stream>>>(rec, c->nF, c->maxF, mu0, var0, c->f_mu, c->f_var, c->f_drop, c->f_alive, c->f_ref);
    c->launches++;
    FB_CUDA(c, cudaGetLastError());
    return FB_OK;
  }
  uint8_t* st = stage_slot(c);
  if (!st) FB_FAIL(c, FB_E_NOMEM, "pinned staging allocation failed");
  memcpy(st, ref_slot, sizeof(int32_t) * c->S);
  const dim3 grid(fb_div_up(c->maxF, 256), c->S);
  // the per-stream slots are read from the pinned staging record (no copy on the H2D engine)
  k_features_reinit<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const int32_t*>(st), c->nF, c->maxF, mu0, var0, c->f_mu, c->f_var, c->f_drop, c->f_alive, c->f_ref);
  stage_commit(c);
  c->launches++;
  FB_CUDA(c, cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_idepth_update(fb_ctx* c, const int32_t* cmp_slot) {
  CHECK_CTX(c);
  if (!cmp_slot) FB_FAIL(c, FB_E_ARG, "fb_idepth_update: null cmp_slot");
  int maxf = 0;
  for (int s = 0; s < c->S; ++s) {
    if (cmp_slot[s] >= c->n_slots) FB_FAIL(c, FB_E_ARG, "fb_idepth_update: cmp_slot out of range");
    if (cmp_slot[s] >= 0) maxf = std::max(maxf, c->hF[s]);
  }
  ProfScope ps(c, FB_PROF_IDEPTH);
  int rc = upload_geometry(c, cmp_slot, true);
  if (rc) return rc;
  // (the counters of the streams this call updates were zeroed by k_epi_geometry; the others keep theirs)
  if (maxf == 0) return FB_OK;
  EpiArgs a;
  a.imgs = c->imgs; a.geo = c->d_geo; a.cmp_slot = c->d_cmp; a.u_ref = c->f_uref;
  a.ref_slot = c->f_ref; a.mu = c->f_mu; a.var = c->f_var; a.dropouts = c->f_drop;
  a.alive = c->f_alive; a.status = c->f_status; a.u_cmp = c->f_ucmp; a.nF = c->nF;
  a.counters = c->counters; a.W = c->W; a.H = c->H; a.n_slots = c->n_slots; a.maxF = c->maxF;
  a.s0 = 0;
  a.cmp_frames = c->epi_cmp_frames;  // set by the pipelined step when the frames sit in a landing buffer
  c->epi_cmp_frames = nullptr;
  a.p = c->epi;
  const int wpb = 8;
  const size_t smem = sizeof(float) * wpb * FB_EPI_GROUPS * (2 * c->epi.max_search_px + 2 * FB_MAX_WIN + 2);
  const dim3 grid(fb_div_up(maxf, wpb * FB_EPI_GROUPS), c->S);
  if (smem > 48 * 1024)
    FB_CUDA(c, cudaFuncSetAttribute(k_epipolar_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_epipolar_search<<<grid, wpb * 32, smem, c->stream>>>(a);
  c->launches++;
  FB_CUDA(c, cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_idepth_counters(fb_ctx* c, int s, int32_t counters[FB_NUM_COUNTERS]) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  FB_CUDA(c, cudaMemcpyAsync(counters, c->counters + s * FB_NUM_COUNTERS, sizeof(int32_t) * FB_NUM_COUNTERS, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FB_OK;
}

extern "C" int fb_project_features(fb_ctx* c, int s, int cur_slot, float* u_cur, float* mu_cur,
                                   float* var_cur, int32_t* valid) {
  CHECK_CTX(c);
  int rc = check_slot(c, s, cur_slot);
  if (rc) return rc;
  const int N = c->hF[s];
  if (N == 0) return FB_OK;
  std::vector<int32_t> cmp(c->S, -1);
  cmp[s] = cur_slot;
  rc = upload_geometry(c, cmp.data());
  if (rc) return rc;
  float2* d_u = nullptr; float *d_mu = nullptr, *d_var = nullptr; int32_t* d_valid = nullptr;
  FB_CUDA(c, dalloc(&d_u, N)); FB_CUDA(c, dalloc(&d_mu, N)); FB_CUDA(c, dalloc(&d_var, N)); FB_CUDA(c, dalloc(&d_valid, N));
  const size_t fb = (size_t)s * c->maxF;
  k_project_features<<<fb_div_up(N, 256), 256, 0, c->stream>>>(c->d_geo, c->n_slots, s, N, c->W, c->H, c->f_uref + fb, c->f_ref + fb, c->f_mu + fb, c->f_var + fb, c->f_alive + fb, d_u, d_mu, d_var, d_valid);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && u_cur) e = cudaMemcpyAsync(u_cur, d_u, sizeof(float2) * N, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && mu_cur) e = cudaMemcpyAsync(mu_cur, d_mu, sizeof(float) * N, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && var_cur) e = cudaMemcpyAsync(var_cur, d_var, sizeof(float) * N, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && valid) e = cudaMemcpyAsync(valid, d_valid, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_u); cudaFree(d_mu); cudaFree(d_var); cudaFree(d_valid);
  FB_CUDA(c, e);
  return FB_OK;
}

// ------------------------------------------------------------------------------------ assembly
extern "C" int fb_graph_bind_features(fb_ctx* c, int s, const int32_t* vertex_feature) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (!vertex_feature) FB_FAIL(c, FB_E_ARG, "fb_graph_bind_features: null input");
  const int V = c->hV[s];
  FB_CUDA(c, cudaMemcpyAsync(c->vfeat + (size_t)s * c->maxV, vertex_feature, sizeof(int32_t) * V, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FB_OK;
}

extern "C" int fb_graph_data_from_features(fb_ctx* c, int adaptive) {
  CHECK_CTX(c);
  ProfScope ps(c, FB_PROF_ASSEMBLY);
  const dim3 grid(fb_div_up(c->maxV, 256), c->S);
  k_data_from_features<<<grid, 256, 0, c->stream>>>(c->z, c->wt, c->vfeat, c->nV, c->maxV, c->f_mu, c->f_var, c->f_alive, c->nF, c->maxF, adaptive);
  c->launches++;
  FB_CUDA(c, cudaGetLastError());
  return FB_OK;
}

// ------------------------------------------------------------------------------------ batched frame
static int pipeline_init(fb_ctx* c) {
  if (c->copy_stream) return FB_OK;
  FB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  c->ev_ready.resize(c->n_slots);
  c->ev_free.resize(c->n_slots);
  c->slot_landing.assign(c->n_slots, -1);
  c->slot_is_ref.assign(c->n_slots, 1);  // unknown history: treat every slot as a poseframe once
  FB_CUDA(c, dalloc(&c->incoming, 2 * (size_t)c->S * c->W * c->H));
  for (int k = 0; k < c->n_slots; ++k) {
    FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_ready[k], cudaEventDisableTiming));
    FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_free[k], cudaEventDisableTiming));
  }
  for (int k = 0; k < 4; ++k) FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_result[k], cudaEventDisableTiming));
  FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming));
  if (!getenv("FB_PIPE_SINGLE_STAGE")) {
    int lo = 0, hi = 0;
    FB_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    FB_CUDA(c, cudaStreamCreateWithPriority(&c->solve_stream, cudaStreamNonBlocking, hi));
    FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_epi, cudaEventDisableTiming));
    FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_asm, cudaEventDisableTiming));
    FB_CUDA(c, cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
    FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_solved, cudaEventDisableTiming));
    if (!getenv("FB_PIPE_SINGLE_DATA")) {
      const size_t nv = (size_t)c->S * c->maxV;
      c->z_buf[0] = c->z; c->wt_buf[0] = c->wt;
      FB_CUDA(c, dalloc(&c->z_buf[1], nv));
      FB_CUDA(c, dalloc(&c->wt_buf[1], nv));
      // dead / unbound vertices keep whatever z holds (their weight is 0, the value never matters, but it must be finite)
      FB_CUDA(c, cudaMemcpy(c->z_buf[1], c->z, sizeof(float) * nv, cudaMemcpyDeviceToDevice));
      FB_CUDA(c, cudaMemcpy(c->wt_buf[1], c->wt, sizeof(float) * nv, cudaMemcpyDeviceToDevice));
      for (int b = 0; b < 2; ++b) FB_CUDA(c, cudaEventCreateWithFlags(&c->ev_zfree[b], cudaEventDisableTiming));
    }
    FB_CUDA(c, dalloc(&c->x_stage[0], (size_t)c->S * c->maxV));
    FB_CUDA(c, dalloc(&c->x_stage[1], (size_t)c->S * c->maxV));
  }
  return FB_OK;
}

// FB_PIPE_TRACE=<steps>: timing events around the stages of the pipelined step, printed once after
// <steps> steps (diagnosis of the overlap; the events themselves cost a few us of host time per step).
struct PipeTrace {
  int n = 0, cap = 0;
  std::vector<cudaEvent_t> ev;  // [cap][10]: h2d0 h2d1 epi0 epi1 asm0 asm1 solve1 d2h1 pfcopy0 pfcopy1
  cudaEvent_t at(int k, int j) { return ev[(size_t)k * 10 + j]; }
};
static PipeTrace* pipe_trace() {
  static PipeTrace* t = nullptr;
  static bool init = false;
  if (!init) {
    init = true;
    const char* e = getenv("FB_PIPE_TRACE");
    if (e && atoi(e) > 0) {
      t = new PipeTrace();
      t->cap = atoi(e);
      t->ev.resize((size_t)t->cap * 10);
      for (auto& x : t->ev) cudaEventCreate(&x);
    }
  }
  return t;
}
static void pipe_trace_mark(PipeTrace* t, int j, cudaStream_t st) {
  if (t && t->n < t->cap) cudaEventRecord(t->at(t->n, j), st);
}
static void pipe_trace_step_done(PipeTrace* t) {
  if (!t || t->n >= t->cap) return;
  if (++t->n < t->cap) return;
  cudaDeviceSynchronize();
  static const char* name[10] = {"h2d0", "h2d1", "epi0", "epi1", "asm0", "asm1", "solve1", "d2h1", "pfcopy0", "pfcopy1"};
  for (int k = 1; k < t->cap; ++k) {
    fprintf(stderr, "[pipe-trace] step %2d:", k);
    for (int j = 0; j < 10; ++j) {
      float ms = -1.f;
      if (cudaEventElapsedTime(&ms, t->at(1, 2), t->at(k, j)) != cudaSuccess) { cudaGetLastError(); ms = -1.f; }
      fprintf(stderr, " %s %7.1f", name[j], ms * 1e3f);
    }
    fprintf(stderr, "\n");
  }
}

// Host image -> slot on the copy stream, ordered after the last kernel that read the slot.
static int pipeline_upload(fb_ctx* c, int slot, const uint8_t* const* images, const int32_t* pool_idx,
                           const float* poses, bool allow_landing) {
  const size_t fsz = (size_t)c->W * c->H;
  FB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_free[slot], 0));
  for (int s = 0; s < c->S; ++s) {
    int rc = fb_frame_pose_set(c, s, slot, poses + 7 * s);
    if (rc) return rc;
    if (!images && (pool_idx[s] < 0 || pool_idx[s] >= c->pool_n)) FB_FAIL(c, FB_E_ARG, "fb_hotpath_step: bad pool index");
  }
  // Device-pool frames of a step usually sit at a constant stride (entries of one replica): then they
  // go as ONE pitched device-to-device copy instead of S launches (measured, 8 streams: 93 -> 89 us
  // per step).  Host frames keep one linear copy each: a pitched H2D copy and a cudaMemcpyBatchAsync
  // of the same 8 frames were both slower on this box (122 -> 135 / 134 us per step).
  uint8_t* dst0 = c->imgs + (size_t)slot * fsz;
  const size_t dpitch = (size_t)c->n_slots * fsz;
  ptrdiff_t stride = 0;
  // Host frames that sit back to back in (pinned) memory -- a multi-camera capture buffer -- go up as
  // ONE linear transfer into a landing buffer (the slots of the S streams are not adjacent in `imgs`,
  // and 8 separate 300 kB copies cost ~12 us each: the e2e leg was bound by them); the epipolar
  // kernel reads the comparison frames from there.
  c->slot_landing[slot] = -1;
  if (images && c->S > 1 && allow_landing) {  // poseframes must live in their slot: features refer to them for many frames
    bool contiguous = true;
    for (int s = 0; s + 1 < c->S; ++s) contiguous = contiguous && images[s + 1] - images[s] == (ptrdiff_t)fsz;
    if (contiguous) {
      const int lb = slot & 1;
      if (allow_landing) pipe_trace_mark(pipe_trace(), 0, c->copy_stream);
      FB_CUDA(c, cudaMemcpyAsync(c->incoming + (size_t)lb * c->S * fsz, images[0], (size_t)c->S * fsz, cudaMemcpyHostToDevice, c->copy_stream));
      if (allow_landing) pipe_trace_mark(pipe_trace(), 1, c->copy_stream);
      c->slot_landing[slot] = lb;
      FB_CUDA(c, cudaEventRecord(c->ev_ready[slot], c->copy_stream));
      FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_ready[slot], 0));
      return FB_OK;
    }
  }
  bool uniform = c->S > 1 && !images;
  for (int s = 0; s + 1 < c->S && uniform; ++s) {
    const ptrdiff_t d = (ptrdiff_t)(pool_idx[s + 1] - pool_idx[s]) * (ptrdiff_t)fsz;
    if (s == 0) stride = d;
    uniform = d == stride && d >= (ptrdiff_t)fsz && d < ((ptrdiff_t)1 << 30);
  }
  if (uniform) {
    const uint8_t* src0 = c->pool + (size_t)pool_idx[0] * fsz;
    FB_CUDA(c, cudaMemcpy2DAsync(dst0, dpitch, src0, (size_t)stride, fsz, (size_t)c->S, cudaMemcpyDeviceToDevice, c->copy_stream));
  } else {
    for (int s = 0; s < c->S; ++s) {
      uint8_t* dst = dst0 + (size_t)s * dpitch;
      if (images) FB_CUDA(c, cudaMemcpyAsync(dst, images[s], fsz, cudaMemcpyHostToDevice, c->copy_stream));
      else FB_CUDA(c, cudaMemcpyAsync(dst, c->pool + (size_t)pool_idx[s] * fsz, fsz, cudaMemcpyDeviceToDevice, c->copy_stream));
    }
  }
  FB_CUDA(c, cudaEventRecord(c->ev_ready[slot], c->copy_stream));
  FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_ready[slot], 0));
  return FB_OK;
}

// The stages of one frame on c->stream (blocking / resident modes).
static int hotpath_step_inline(fb_ctx* c, const fb_step_desc* d) {
  int rc;
  std::vector<int32_t> slots(c->S);
  if (d->new_poseframe && d->ref_from_slot > 0) {
    const int from = d->ref_from_slot - 1;
    if ((rc = check_slot(c, 0, from)) != 0) return rc;
    const size_t fsz = (size_t)c->W * c->H, pitch = (size_t)c->n_slots * fsz;
    for (int s = 0; s < c->S; ++s)
      if ((rc = fb_frame_pose_set(c, s, d->ref_slot, d->ref_poses + 7 * s)) != 0) return rc;
    FB_CUDA(c, cudaMemcpy2DAsync(c->imgs + (size_t)d->ref_slot * fsz, pitch, c->imgs + (size_t)from * fsz, pitch, fsz, (size_t)c->S,
                                 cudaMemcpyDeviceToDevice, c->stream));
  } else if (d->new_poseframe) {
    for (int s = 0; s < c->S; ++s) {
      rc = d->ref_images ? fb_frame_set(c, s, d->ref_slot, d->ref_images[s], c->W, d->ref_poses + 7 * s)
                         : fb_frame_from_pool(c, s, d->ref_slot, d->ref_pool_idx[s], d->ref_poses + 7 * s);
      if (rc) return rc;
    }
  }
  for (int s = 0; s < c->S; ++s) {
    rc = d->cmp_images ? fb_frame_set(c, s, d->cmp_slot, d->cmp_images[s], c->W, d->cmp_poses + 7 * s)
                       : fb_frame_from_pool(c, s, d->cmp_slot, d->cmp_pool_idx[s], d->cmp_poses + 7 * s);
    if (rc) return rc;
  }
  if (d->new_poseframe) {
    std::fill(slots.begin(), slots.end(), d->ref_slot);
    rc = fb_features_reinit(c, slots.data(), d->mu0, d->var0);
    if (rc) return rc;
  }
  std::fill(slots.begin(), slots.end(), d->cmp_slot);
  rc = fb_idepth_update(c, slots.data());
  if (rc) return rc;
  rc = fb_graph_data_from_features(c, d->adaptive_weights);
  if (rc) return rc;
  rc = fb_nltgv2_solve(c, d->iters, &d->rparams, d->variant);
  if (rc) return rc;
  if (d->x_out) return fb_graph_x_get_all(c, d->x_out);
  return FB_OK;
}

// Streaming mode: a software pipeline over consecutive frames on three streams.
//   copy stream : frame k+1 into its slot (H2D from pinned host memory, or D2D from the device pool),
//                 ordered after the kernels that last read the slot
//   c->stream   : feature re-init + epipolar update of frame k+1
//   solve stream: data-term assembly + NLTGV2 solve + D2H of frame k
// The filter update of frame k+1 does not depend on the solve of frame k; the assembly (which
// snapshots mu into z) is the only hand-over.  Measured on B200, 8 streams: 149 us per frame batch
// against 176 us with a single compute stream.
static int hotpath_step_pipelined(fb_ctx* c, const fb_step_desc* d) {
  int rc = pipeline_init(c);
  if (rc) return rc;
  PipeTrace* tr = pipe_trace();
  if ((rc = check_slot(c, 0, d->cmp_slot)) != 0) return rc;
  if (d->new_poseframe && (rc = check_slot(c, 0, d->ref_slot)) != 0) return rc;
  c->pipe_dirty = true;
  std::vector<int32_t> slots(c->S);
  if (d->new_poseframe && d->ref_from_slot > 0) {
    // the poseframe is a frame that is already on the device: one pitched device copy on the main
    // stream (ordered after every kernel that read either slot), no upload
    const int from = d->ref_from_slot - 1;
    if ((rc = check_slot(c, 0, from)) != 0) return rc;
    const size_t fsz = (size_t)c->W * c->H, pitch = (size_t)c->n_slots * fsz;
    for (int s = 0; s < c->S; ++s)
      if ((rc = fb_frame_pose_set(c, s, d->ref_slot, d->ref_poses + 7 * s)) != 0) return rc;
    const bool landed = c->slot_landing[from] >= 0;
    const uint8_t* src = landed ? c->incoming + (size_t)c->slot_landing[from] * c->S * fsz : c->imgs + (size_t)from * fsz;
    pipe_trace_mark(tr, 8, c->stream);
    FB_CUDA(c, cudaMemcpy2DAsync(c->imgs + (size_t)d->ref_slot * fsz, pitch, src, landed ? fsz : pitch, fsz, (size_t)c->S,
                                 cudaMemcpyDeviceToDevice, c->stream));
    pipe_trace_mark(tr, 9, c->stream);
    FB_CUDA(c, cudaEventRecord(c->ev_free[from], c->stream));  // the next upload into `from` waits for this read
  } else if (d->new_poseframe) {
    rc = pipeline_upload(c, d->ref_slot, d->ref_images, d->ref_pool_idx, d->ref_poses, false);
    if (rc) return rc;
  }
  rc = pipeline_upload(c, d->cmp_slot, d->cmp_images, d->cmp_pool_idx, d->cmp_poses, true);
  if (rc) return rc;
  if (d->new_poseframe) {
    std::fill(slots.begin(), slots.end(), d->ref_slot);
    rc = fb_features_reinit(c, slots.data(), d->mu0, d->var0);
    if (rc) return rc;
  }
  std::fill(slots.begin(), slots.end(), d->cmp_slot);
  if (c->slot_landing[d->cmp_slot] >= 0) c->epi_cmp_frames = c->incoming + (size_t)c->slot_landing[d->cmp_slot] * c->S * c->W * c->H;
  pipe_trace_mark(tr, 2, c->stream);
  rc = fb_idepth_update(c, slots.data());
  if (rc) return rc;
  pipe_trace_mark(tr, 3, c->stream);
  // the frames read by this update may be overwritten once the epipolar kernel has run
  FB_CUDA(c, cudaEventRecord(c->ev_free[d->cmp_slot], c->stream));
  if (d->new_poseframe) {
    // the OTHER poseframe slots are no longer referenced by any feature after the re-init.  Only slots
    // that held a poseframe: re-recording the event of the other COMPARISON slot here made the next
    // frame's upload wait for this frame's epipolar update (52 us of exposed H2D every epoch).
    for (int k = 0; k < c->n_slots; ++k)
      if (k != d->cmp_slot && k != d->ref_slot && c->slot_is_ref[k]) {
        FB_CUDA(c, cudaEventRecord(c->ev_free[k], c->stream));
        c->slot_is_ref[k] = 0;
      }
    c->slot_is_ref[d->ref_slot] = 1;
  }
  // Second stage on its own (high-priority) stream: the solve + D2H of frame k overlap the epipolar
  // update AND the assembly of frame k+1 (FB_PIPE_SINGLE_STAGE=1 disables).  The assembly snapshots mu
  // into the data term z / wt, which is double-buffered: it runs on the main stream right behind the
  // epipolar update into the buffer the running solve does not read (on the solve stream its 8 us sat
  // in the cycle that bounds the step: assembly + solve).
  cudaStream_t main_stream = c->stream;
  const bool dbuf = c->solve_stream && c->z_buf[1];
  const int zb = (int)(c->n_pipe_steps & 1);
  if (dbuf) {
    c->z = c->z_buf[zb];
    c->wt = c->wt_buf[zb];
    if (c->zfree_valid[zb]) FB_CUDA(c, cudaStreamWaitEvent(main_stream, c->ev_zfree[zb], 0));  // the solve of frame k-2 read this buffer
  } else if (c->solve_stream) {
    // ev_free[cmp_slot] was recorded right after the epipolar kernel: the same point the second stage waits for
    FB_CUDA(c, cudaStreamWaitEvent(c->solve_stream, c->ev_free[d->cmp_slot], 0));
    c->stream = c->solve_stream;
  }
  pipe_trace_mark(tr, 4, c->stream);
  rc = fb_graph_data_from_features(c, d->adaptive_weights);
  pipe_trace_mark(tr, 5, c->stream);
  if (!rc && dbuf) {
    if (cudaEventRecord(c->ev_asm, main_stream) != cudaSuccess || cudaStreamWaitEvent(c->solve_stream, c->ev_asm, 0) != cudaSuccess) rc = FB_E_CUDA;
    c->stream = c->solve_stream;
  } else if (!rc && c->solve_stream) {
    // the next frame's filter update may touch the feature table once the assembly has read it
    if (cudaEventRecord(c->ev_asm, c->stream) != cudaSuccess || cudaStreamWaitEvent(main_stream, c->ev_asm, 0) != cudaSuccess) rc = FB_E_CUDA;
  }
  if (!rc) rc = fb_nltgv2_solve(c, d->iters, &d->rparams, d->variant);
  pipe_trace_mark(tr, 6, c->stream);
  if (!rc && d->x_out) {
    // The read-back runs on a stream of its own from a device snapshot of x.  Queued on the solve
    // stream, or reading x itself (the next solve then has to wait for the copy), its ~12 us of DMA
    // latency sat on the critical path of every step (88 us per step against 78 without read-back).
    const size_t bytes = sizeof(float) * (size_t)c->S * c->maxV;
    const int n = c->n_pipelined;
    if (c->out_stream) {
      float* snap = c->x_stage[n & 1];
      if ((n >= 2 && cudaStreamWaitEvent(c->stream, c->ev_result[(n - 2) & 3], 0) != cudaSuccess) ||  // the copy that last read this snapshot
          cudaMemcpyAsync(snap, c->x, bytes, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess ||
          cudaEventRecord(c->ev_solved, c->stream) != cudaSuccess || cudaStreamWaitEvent(c->out_stream, c->ev_solved, 0) != cudaSuccess ||
          cudaMemcpyAsync(d->x_out, snap, bytes, cudaMemcpyDeviceToHost, c->out_stream) != cudaSuccess ||
          cudaEventRecord(c->ev_result[n & 3], c->out_stream) != cudaSuccess)
        rc = FB_E_CUDA;
    } else if (cudaMemcpyAsync(d->x_out, c->x, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
               cudaEventRecord(c->ev_result[n & 3], c->stream) != cudaSuccess) {
      rc = FB_E_CUDA;
    }
    cudaStream_t os = c->out_stream ? c->out_stream : c->stream;
    c->n_pipelined++;
    pipe_trace_mark(tr, 7, os);
  }
  if (!rc && dbuf) {
    if (cudaEventRecord(c->ev_zfree[zb], c->stream) != cudaSuccess) rc = FB_E_CUDA;
    c->zfree_valid[zb] = true;
  }
  c->n_pipe_steps++;
  c->stream = main_stream;
  pipe_trace_step_done(tr);
  if (rc == FB_E_CUDA) c->err = "fb_hotpath_step: CUDA error in the pipelined step";
  return rc;
}

extern "C" int fb_hotpath_step(fb_ctx* c, const fb_step_desc* d) {
  CHECK_CTX_NODRAIN(c);
  if (!d || !d->cmp_poses || (!d->cmp_images && !d->cmp_pool_idx))
    FB_FAIL(c, FB_E_ARG, "fb_hotpath_step: null descriptor field");
  if (d->new_poseframe && (!d->ref_poses || (d->ref_from_slot <= 0 && !d->ref_images && !d->ref_pool_idx)))
    FB_FAIL(c, FB_E_ARG, "fb_hotpath_step: poseframe inputs missing");
  if (d->pipelined) {
    c->pipe_hold = true;  // the building blocks called inside must not wait for the work in flight
    const int rc = hotpath_step_pipelined(c, d);
    c->pipe_hold = false;
    return rc;
  }
  pipeline_drain(c);
  return hotpath_step_inline(c, d);
}

// Makes c->stream wait (on the device, not the host) for everything the pipelined steps have
// enqueued on the auxiliary streams, so an event recorded on c->stream afterwards closes the region.
extern "C" int fb_pipeline_join(fb_ctx* c) {
  CHECK_CTX_NODRAIN(c);
  if (!c->copy_stream) return FB_OK;
  FB_CUDA(c, cudaEventRecord(c->ev_join, c->copy_stream));
  FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  if (c->solve_stream) {
    FB_CUDA(c, cudaEventRecord(c->ev_join2, c->solve_stream));
    FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join2, 0));
  }
  if (c->out_stream && c->n_pipelined > 0) FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_result[(c->n_pipelined - 1) & 3], 0));
  return FB_OK;
}

extern "C" int fb_results_wait(fb_ctx* c, int lag) {
  CHECK_CTX_NODRAIN(c);
  if (lag < 0 || lag > 3) FB_FAIL(c, FB_E_ARG, "fb_results_wait: lag must be in [0,3]");
  if (c->n_pipelined - 1 - lag < 0) return FB_OK;
  FB_CUDA(c, cudaEventSynchronize(c->ev_result[(c->n_pipelined - 1 - lag) & 3]));
  if (grid_watchdog_fired(c)) FB_FAIL(c, FB_E_STATE, "grid-resident solver: mailbox exchange timed out (watchdog)");
  return FB_OK;
}

// ------------------------------------------------------------------------------------ triangulation
extern "C" int fb_delaunay(int n, const float* pts, int32_t* tris, int32_t* n_tris, int32_t* edges,
                           int32_t* n_edges) {
  if (n < 0 || (n > 0 && !pts) || !tris || !n_tris || !edges || !n_edges) return FB_E_ARG;
  *n_tris = 0;
  *n_edges = 0;
  fbdel::Triangulator T;
  std::vector<int> t, e;
  if (!T.run(n, pts, t, e)) return FB_E_ARG;
  std::copy(t.begin(), t.end(), tris);
  std::copy(e.begin(), e.end(), edges);
  *n_tris = (int32_t)(t.size() / 3);
  *n_edges = (int32_t)(e.size() / 2);
  return FB_OK;
}

// ------------------------------------------------------------------------------------ interpolation
extern "C" int fb_mesh_set(fb_ctx* c, int s, int T, const int32_t* tri) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (T < 0 || T > c->maxT) FB_FAIL(c, FB_E_NOMEM, "fb_mesh_set: T exceeds capacity (2*max_vertices)");
  if (T > 0 && !tri) FB_FAIL(c, FB_E_ARG, "fb_mesh_set: null triangles");
  const int V = c->hV[s];
  for (int k = 0; k < 3 * T; ++k)
    if (tri[k] < 0 || tri[k] >= V) FB_FAIL(c, FB_E_ARG, "fb_mesh_set: vertex id out of range");
  if (T) FB_CUDA(c, cudaMemcpyAsync(c->tri + (size_t)s * c->maxT * 3, tri, sizeof(int32_t) * 3 * T, cudaMemcpyHostToDevice, c->stream));
  c->hT[s] = T;
  FB_CUDA(c, cudaMemcpyAsync(c->nT + s, &c->hT[s], sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FB_OK;
}

extern "C" int fb_interpolate(fb_ctx* c, int s, const fb_tri_filter_params* filter,
                              float* idepthmap, uint8_t* tri_valid) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  const int T = c->hT[s];
  const size_t npx = (size_t)c->W * c->H;
  const size_t vb = (size_t)s * c->maxV;
  int32_t* owner = c->owner + (size_t)s * npx;
  // The unfiltered map is persistent state: the next fb_update reads it as the prediction for new
  // vertices / features.  A filtered request (getFilteredInverseDepthMap) must not replace it, so it
  // is rendered into a scratch map.
  float* map = c->idmap + (size_t)s * npx;
  if (filter) {
    if (!c->idmap_scratch) FB_CUDA(c, dalloc(&c->idmap_scratch, npx));
    map = c->idmap_scratch;
  }
  uint8_t* valid = c->tri_valid + (size_t)s * c->maxT;
  const int32_t* tri = c->tri + (size_t)s * c->maxT * 3;
  {
    ProfScope ps(c, FB_PROF_INTERP);
    if (T) {
      fb_tri_filter_params fp;
      fb_default_tri_filter_params(&fp);
      float cos_thresh = 0.f;
      if (filter) {
        fp = *filter;
        cos_thresh = (float)cos((double)fp.oblique_normal_thresh);
      }
      k_tri_validity<<<fb_div_up(T, 256), 256, 0, c->stream>>>(c->W, c->d_K + 9 * s, c->vpos + vb, c->x + vb, T, tri, fp, cos_thresh, filter ? 1 : 0, valid);
      k_raster_claim<<<fb_div_up(T * 32, 256), 256, 0, c->stream>>>(c->W, c->H, c->vpos + vb, T, tri, valid, owner);
      c->launches += 2;
    }
    k_raster_shade<<<fb_div_up((int)npx, 256), 256, 0, c->stream>>>(c->W, c->H, c->vpos + vb, c->x + vb, tri, owner, map);
    c->launches++;
    FB_CUDA(c, cudaGetLastError());
  }
  if (idepthmap) FB_CUDA(c, cudaMemcpyAsync(idepthmap, map, sizeof(float) * npx, cudaMemcpyDeviceToHost, c->stream));
  if (tri_valid && T) FB_CUDA(c, cudaMemcpyAsync(tri_valid, valid, T, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FB_OK;
}

// ------------------------------------------------------------------------------------ update pipeline
extern "C" void fb_default_update_params(fb_update_params* p) {
  // detection win 16 / min_grad 5 / idepth_var_max 0.01: /root/reference/cfg/flame_nodelet.yaml:69-71,85
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
  fb_default_nltgv2_params(&p->rparams);
  p->triangulator = 0;
  // /root/reference/cfg/flame_nodelet.yaml:67,70,83,90-92
  p->rescale_data = 0;
  p->min_height = -1e14f;
  p->max_height = 1e14f;
  p->check_sticky_obstacles = 0;
  p->min_error = 100.0f;
  p->do_letterbox = 0;
}

#include "flame_update.cuh"

static void update_mark_host_graph(fb_ctx* c, int s) {
  if (c->upd) c->upd->st[s].dev_graph = false;
}
static void update_invalidate_graphs(fb_ctx* c) {
  if (!c->upd) return;
  for (auto& st : c->upd->st)
    for (int k = 0; k < 2; ++k)
      if (st.frame_graph[k]) {
        cudaGraphExecDestroy(st.frame_graph[k]);
        st.frame_graph[k] = nullptr;
      }
}
static bool any_device_graph(const fb_ctx* c) {
  if (!c->upd) return false;
  for (int s = 0; s < c->S; ++s)
    if (c->upd->st[s].dev_graph && c->hV[s] > 0) return true;
  return false;
}

extern "C" int fb_set_update_params(fb_ctx* c, const fb_update_params* p) {
  CHECK_CTX(c);
  if (!p || p->detection_win_size < 4 || p->detection_win_size > 64 || p->iters < 0 || p->detection_border < 1 ||
      p->triangulator < 0 || p->triangulator > 1)
    FB_FAIL(c, FB_E_ARG, "fb_set_update_params: bad parameters (win in [4,64], border >= 1, triangulator 0|1)");
  if (p->triangulator == 0 && c->maxV > 65535)
    FB_FAIL(c, FB_E_ARG, "fb_set_update_params: the device triangulation packs vertex ranks in 16 bits (max_vertices <= 65535); set triangulator = 1");
  if (p->check_sticky_obstacles != 0)
    FB_FAIL(c, FB_E_ARG, "fb_set_update_params: check_sticky_obstacles is not implemented (only 0 is accepted)");
  if (p->min_error != 100.0f)
    FB_FAIL(c, FB_E_ARG, "fb_set_update_params: the detector's photometric-error gate is not implemented (min_error must stay at its default 100)");
  if (!(p->min_height <= p->max_height))
    FB_FAIL(c, FB_E_ARG, "fb_set_update_params: min_height > max_height");
  int rc = update_alloc(c);
  if (rc) return rc;
  c->upd->up = *p;
  update_invalidate_graphs(c);
  return FB_OK;
}

extern "C" int fb_update(fb_ctx* c, int s, double time, int img_id, const float pose[7],
                         const uint8_t* gray, int pitch, int is_poseframe) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (!pose || !gray) FB_FAIL(c, FB_E_ARG, "fb_update: null input");
  if (c->n_slots < 3) FB_FAIL(c, FB_E_STATE, "fb_update: needs n_slots >= 3 (poseframe ring + current frame)");
  return fb_update_impl(c, s, time, img_id, pose, gray, pitch, is_poseframe);
}

extern "C" int fb_get_mesh_sizes(fb_ctx* c, int s, int32_t* V, int32_t* T, int32_t* E) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  const bool have = c->upd && c->upd->st[s].have_graph;
  if (V) *V = have ? c->hV[s] : 0;
  if (T) *T = have ? c->hT[s] : 0;
  if (E) *E = have ? c->hE[s] : 0;
  return FB_OK;
}

extern "C" int fb_get_mesh(fb_ctx* c, int s, const fb_tri_filter_params* filter, float* vtx_xy,
                           float* idepth, float* normals, int32_t* tris, uint8_t* tri_valid,
                           int32_t* edges) {
  CHECK_CTX_RO(c);
  CHECK_STREAM(c, s);
  if (!c->upd || !c->upd->st[s].have_graph) FB_FAIL(c, FB_E_STATE, "fb_get_mesh: no mesh yet");
  UpdateStream& S = c->upd->st[s];
  const int V = c->hV[s], T = c->hT[s];
  const size_t vb = (size_t)s * c->maxV;
  cudaStream_t st = c->stream;
  std::vector<float> x(V), w1(V), w2(V);
  std::vector<float2> pos(V);
  FB_CUDA(c, cudaMemcpyAsync(x.data(), c->x + vb, sizeof(float) * V, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(w1.data(), c->w1 + vb, sizeof(float) * V, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(w2.data(), c->w2 + vb, sizeof(float) * V, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(pos.data(), c->vpos + vb, sizeof(float2) * V, cudaMemcpyDeviceToHost, st));
  // the mesh lives on the device (built there by fb_update, or uploaded by fb_graph_set / fb_mesh_set)
  if (tris && T) FB_CUDA(c, cudaMemcpyAsync(tris, c->tri + (size_t)s * c->maxT * 3, sizeof(int32_t) * 3 * T, cudaMemcpyDeviceToHost, st));
  if (edges && c->hE[s]) FB_CUDA(c, cudaMemcpyAsync(edges, c->eij + (size_t)s * c->maxE, sizeof(int32_t) * 2 * c->hE[s], cudaMemcpyDeviceToHost, st));
  if (tri_valid && T) {
    fb_tri_filter_params fp;
    fb_default_tri_filter_params(&fp);
    float cos_thresh = 0.f;
    if (filter) {
      fp = *filter;
      cos_thresh = (float)cos((double)fp.oblique_normal_thresh);
    }
    uint8_t* valid = c->tri_valid + (size_t)s * c->maxT;
    k_tri_validity<<<fb_div_up(T, 256), 256, 0, st>>>(c->W, c->d_K + 9 * s, c->vpos + vb, c->x + vb, T, c->tri + (size_t)s * c->maxT * 3, fp, cos_thresh, 1, valid);
    c->launches++;
    FB_CUDA(c, cudaMemcpyAsync(tri_valid, valid, T, cudaMemcpyDeviceToHost, st));
  }
  FB_CUDA(c, cudaStreamSynchronize(st));
  const float* K = &c->h_K[9 * s];
  for (int v = 0; v < V; ++v) {
    if (vtx_xy) { vtx_xy[2 * v] = pos[v].x; vtx_xy[2 * v + 1] = pos[v].y; }
    if (idepth) idepth[v] = x[v];
    if (normals) {
      // idepth(u,v) = w1 u + w2 v + c0 is the plane n.X = d seen through K: n/d = K^T (w1, w2, c0)
      const float c0 = x[v] - w1[v] * pos[v].x - w2[v] * pos[v].y;
      float nx = K[0] * w1[v], ny = K[4] * w2[v], nz = K[2] * w1[v] + K[5] * w2[v] + c0;
      const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
      if (nn > 0.f) { nx /= nn; ny /= nn; nz /= nn; }
      // point the normal toward the camera (negative z in the RDF optical frame)
      if (nz > 0.f) { nx = -nx; ny = -ny; nz = -nz; }
      normals[3 * v] = nx; normals[3 * v + 1] = ny; normals[3 * v + 2] = nz;
    }
  }
  (void)S;
  return FB_OK;
}

extern "C" int fb_get_idepthmap(fb_ctx* c, int s, const fb_tri_filter_params* filter, float* out) {
  CHECK_CTX_RO(c);
  CHECK_STREAM(c, s);
  if (!out) FB_FAIL(c, FB_E_ARG, "fb_get_idepthmap: null output");
  if (!c->upd || !c->upd->st[s].have_graph) {
    const float qnan = nanf("");
    std::fill(out, out + (size_t)c->W * c->H, qnan);
    return FB_OK;
  }
  const size_t npx = (size_t)c->W * c->H;
  if (!filter) {  // rendered by the last fb_update; no need to rasterise again
    FB_CUDA(c, cudaMemcpyAsync(out, c->idmap + (size_t)s * npx, sizeof(float) * npx, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FB_OK;
  }
  UpdateStream& S = c->upd->st[s];
  const bool same = S.spec_on && memcmp(&S.spec_filter, filter, sizeof(*filter)) == 0;
  if (same && S.spec_epoch == c->mut_epoch && c->idmap_f) {
    // the last fb_update rendered exactly this map beside the unfiltered one, and nothing has touched the
    // context since (read-only getters aside)
    FB_CUDA(c, cudaMemcpyAsync(out, c->idmap_f + (size_t)s * npx, sizeof(float) * npx, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    S.stats["filtered_maps_reused"] += 1.0;
    return FB_OK;
  }
  const int rc = fb_interpolate(c, s, filter, out, nullptr);
  if (rc == FB_OK && !same && !getenv("FB_NO_SPEC_MAP")) {
    // remember the filter: from the next frame on fb_update renders this map in its own raster pass
    if (!c->idmap_f) {
      FB_CUDA(c, dalloc(&c->idmap_f, (size_t)c->S * npx));
      FB_CUDA(c, dalloc(&c->owner2, (size_t)c->S * npx));
      FB_CUDA(c, cudaMemsetAsync(c->owner2, 0x7f, sizeof(int32_t) * (size_t)c->S * npx, c->stream));
    }
    S.spec_on = true;
    S.spec_filter = *filter;
    S.spec_epoch = ~0ull;
    update_invalidate_graphs(c);  // the captured frames do not hold the second map yet
  }
  return rc;
}

extern "C" int fb_update_run(fb_ctx* c, int s, int k0, int k1, const uint8_t* frames, size_t frame_stride,
                             const float* poses, int poseframe_every, const fb_tri_filter_params* filter, float* out_map) {
  if (!c) return FB_E_ARG;
  if (!frames || !poses || !out_map || k1 < k0 || poseframe_every < 1) FB_FAIL(c, FB_E_ARG, "fb_update_run: bad argument");
  int n = 0;
  for (int k = k0; k < k1; ++k) {
    int rc = fb_update(c, s, (double)k / 30.0, k, poses + 7 * (size_t)k, frames + (size_t)k * frame_stride, c->W, k % poseframe_every == 0);
    if (rc < 0) return rc;
    if (rc == 1) {
      rc = fb_get_idepthmap(c, s, filter, out_map);
      if (rc < 0) return rc;
      ++n;
    }
  }
  return n;
}

extern "C" int fb_get_raw_idepths(fb_ctx* c, int s, int32_t* N, float* xy, float* mu, float* var) {
  CHECK_CTX_RO(c);
  CHECK_STREAM(c, s);
  if (!N) FB_FAIL(c, FB_E_ARG, "fb_get_raw_idepths: null count");
  *N = 0;
  if (!c->upd) return FB_OK;
  UpdateState* U = c->upd;
  const size_t fb = (size_t)s * c->maxF;
  std::vector<float2> u(c->maxF);
  std::vector<float> m(c->maxF), v(c->maxF);
  std::vector<int32_t> valid(c->maxF);
  cudaStream_t st = c->stream;
  FB_CUDA(c, cudaMemcpyAsync(u.data(), U->f_ucur + fb, sizeof(float2) * c->maxF, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(m.data(), U->f_mucur + fb, sizeof(float) * c->maxF, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(v.data(), U->f_varcur + fb, sizeof(float) * c->maxF, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(valid.data(), U->f_valid + fb, sizeof(int32_t) * c->maxF, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  int n = 0;
  for (int f = 0; f < c->maxF; ++f)
    if (valid[f]) {
      if (xy) { xy[2 * n] = u[f].x; xy[2 * n + 1] = u[f].y; }
      if (mu) mu[n] = m[f];
      if (var) var[n] = v[f];
      ++n;
    }
  *N = n;
  return FB_OK;
}

extern "C" int fb_get_stat(fb_ctx* c, int s, const char* key, double* value) {
  CHECK_CTX_RO(c);
  CHECK_STREAM(c, s);
  if (!key || !value || !c->upd) FB_FAIL(c, FB_E_ARG, "fb_get_stat: bad argument");
  auto& m = c->upd->st[s].stats;
  auto it = m.find(key);
  if (it == m.end()) FB_FAIL(c, FB_E_ARG, std::string("fb_get_stat: unknown key ") + key);
  *value = it->second;
  return FB_OK;
}

extern "C" int fb_update_poseframe_poses(fb_ctx* c, int s, int n, const int32_t* ids, const float* poses) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (n < 0 || (n > 0 && (!ids || !poses))) FB_FAIL(c, FB_E_ARG, "fb_update_poseframe_poses: bad argument");
  if (!c->upd) return FB_OK;
  UpdateStream& S = c->upd->st[s];
  for (int k = 0; k < n; ++k)
    for (size_t slot = 0; slot < S.pf_img_id.size(); ++slot)
      if (S.pf_img_id[slot] == ids[k])
        memcpy(&c->h_pose[((size_t)s * c->n_slots + slot) * 7], poses + 7 * k, sizeof(float) * 7);
  return FB_OK;
}

extern "C" int fb_prune_poseframes(fb_ctx* c, int s, int n, const int32_t* keep) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (n < 0 || (n > 0 && !keep)) FB_FAIL(c, FB_E_ARG, "fb_prune_poseframes: bad argument");
  if (!c->upd) return FB_OK;
  UpdateStream& S = c->upd->st[s];
  const size_t fb = (size_t)s * c->maxF;
  for (size_t slot = 0; slot < S.pf_img_id.size(); ++slot) {
    if (S.pf_img_id[slot] < 0) continue;
    bool kept = false;
    for (int k = 0; k < n; ++k) kept = kept || keep[k] == S.pf_img_id[slot];
    if (!kept) {
      k_kill_ref_slot<<<fb_div_up(c->maxF, 256), 256, 0, c->stream>>>(c->maxF, (int)slot, c->f_alive + fb, c->f_ref + fb);
      c->launches++;
      S.pf_img_id[slot] = -1;
    }
  }
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FB_OK;
}

extern "C" int fb_get_feature_pool(fb_ctx* c, int s, float* u_ref, int32_t* ref_slot, float* mu,
                                   float* var, int32_t* dropouts, int32_t* alive) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  const size_t fb = (size_t)s * c->maxF;
  const int N = c->maxF;
  cudaStream_t st = c->stream;
  if (u_ref) FB_CUDA(c, cudaMemcpyAsync(u_ref, c->f_uref + fb, sizeof(float2) * N, cudaMemcpyDeviceToHost, st));
  if (ref_slot) FB_CUDA(c, cudaMemcpyAsync(ref_slot, c->f_ref + fb, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
  if (mu) FB_CUDA(c, cudaMemcpyAsync(mu, c->f_mu + fb, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
  if (var) FB_CUDA(c, cudaMemcpyAsync(var, c->f_var + fb, sizeof(float) * N, cudaMemcpyDeviceToHost, st));
  if (dropouts) FB_CUDA(c, cudaMemcpyAsync(dropouts, c->f_drop + fb, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
  if (alive) FB_CUDA(c, cudaMemcpyAsync(alive, c->f_alive + fb, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  return FB_OK;
}

// Device triangulation of an arbitrary point set through the update pipeline's kernels
// (k_ds_prepare .. k_ds_emit): the points stand in for the projected features of `stream`.
extern "C" int fb_delaunay_device(fb_ctx* c, int s, int n, const float* pts, int32_t* tris, int32_t* n_tris,
                                  int32_t* edges, int32_t* n_edges) {
  CHECK_CTX(c);
  CHECK_STREAM(c, s);
  if (n < 0 || (n > 0 && !pts) || !tris || !n_tris || !edges || !n_edges) FB_FAIL(c, FB_E_ARG, "fb_delaunay_device: bad argument");
  if (n > c->maxF || n > c->maxV) FB_FAIL(c, FB_E_NOMEM, "fb_delaunay_device: n exceeds max_features / max_vertices");
  if (c->maxV > 65535) FB_FAIL(c, FB_E_ARG, "fb_delaunay_device: max_vertices <= 65535");
  int rc = update_alloc(c);
  if (rc) return rc;
  UpdateState* U = c->upd;
  DelGpu& D = U->del;
  const size_t fb = (size_t)s * c->maxF, vb = (size_t)s * c->maxV, eb = (size_t)s * c->maxE;
  cudaStream_t st = c->stream;
  *n_tris = *n_edges = 0;
  std::vector<int32_t> valid(c->maxF, 0);
  std::fill(valid.begin(), valid.begin() + n, 1);
  FB_CUDA(c, cudaMemcpyAsync(U->f_valid + fb, valid.data(), sizeof(int32_t) * c->maxF, cudaMemcpyHostToDevice, st));
  if (n) FB_CUDA(c, cudaMemcpyAsync(U->f_ucur + fb, pts, sizeof(float2) * n, cudaMemcpyHostToDevice, st));
  FB_CUDA(c, cudaMemsetAsync(U->f_varcur + fb, 0, sizeof(float) * c->maxF, st));
  DsgSelect q{U->f_ucur + fb, U->f_varcur + fb, U->f_valid + fb, 1.0f, c->maxF, c->maxV, c->W, c->H, 0, nullptr, nullptr, nullptr, 0.f, 0.f};
  k_ds_prepare<<<1, DSG_THREADS, 0, st>>>(q, D, s, c->vfeat + vb, c->vpos + vb, D.f2v + fb, c->nV + s);
  k_ds_stars<<<dsg_stars_grid(c->device, c->maxV), DSG_WARPS * 32, 0, st>>>(D, s, c->maxV);
  k_ds_scan<<<1, DSG_THREADS, 0, st>>>(D, s, c->maxV, c->maxE, c->maxT, c->row + (size_t)s * (c->maxV + 1), c->nE + s, c->nT + s);
  k_ds_emit<<<fb_div_up(c->maxV, 4), 128, 0, st>>>(D, s, c->maxV, c->maxE, c->maxT, c->vpos + vb, c->eij + eb, c->ec + eb,
                                                    c->tri + (size_t)s * c->maxT * 3);
  c->launches += 4;
  FB_CUDA(c, cudaGetLastError());
  int32_t meta[DSG_META];
  FB_CUDA(c, cudaMemcpyAsync(meta, D.meta + (size_t)s * DSG_META, sizeof(meta), cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  // this entry point leaves no usable graph behind
  c->hV[s] = c->hE[s] = c->hT[s] = 0;
  int32_t zero[DSG_META] = {0};
  FB_CUDA(c, cudaMemcpyAsync(D.meta + (size_t)s * DSG_META, zero, sizeof(zero), cudaMemcpyHostToDevice, st));
  FB_CUDA(c, cudaMemsetAsync(c->nV + s, 0, sizeof(int32_t), st));
  FB_CUDA(c, cudaMemsetAsync(c->nE + s, 0, sizeof(int32_t), st));
  FB_CUDA(c, cudaMemsetAsync(c->nT + s, 0, sizeof(int32_t), st));
  if (meta[DSG_ERR]) {
    char msg[96];
    snprintf(msg, sizeof(msg), "fb_delaunay_device: failed (flags 0x%x)", meta[DSG_ERR]);
    FB_CUDA(c, cudaStreamSynchronize(st));
    FB_FAIL(c, FB_E_STATE, msg);
  }
  const int T = meta[DSG_NT], E = meta[DSG_NE];
  if (T) FB_CUDA(c, cudaMemcpyAsync(tris, c->tri + (size_t)s * c->maxT * 3, sizeof(int32_t) * 3 * T, cudaMemcpyDeviceToHost, st));
  if (E) FB_CUDA(c, cudaMemcpyAsync(edges, c->eij + eb, sizeof(int32_t) * 2 * E, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  *n_tris = T;
  *n_edges = E;
  return T > 0 ? FB_OK : FB_E_ARG;  // degenerate input: same contract as fb_delaunay
}

// ------------------------------------------------------------------------------------ frame creation / detection
extern "C" int fb_frame_gradient(fb_ctx* c, int s, int slot, float* mag) {
  CHECK_CTX(c);
  int rc = check_slot(c, s, slot);
  if (rc) return rc;
  if (!mag) FB_FAIL(c, FB_E_ARG, "fb_frame_gradient: null output");
  const size_t npx = (size_t)c->W * c->H;
  float* d = nullptr;
  FB_CUDA(c, dalloc(&d, npx));
  k_gradient_mag<<<fb_div_up((int)npx, 256), 256, 0, c->stream>>>(c->W, c->H, c->imgs + ((size_t)s * c->n_slots + slot) * npx, d);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(mag, d, sizeof(float) * npx, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  FB_CUDA(c, e);
  return FB_OK;
}

extern "C" int fb_frame_pyr_down(fb_ctx* c, int s, int slot, uint8_t* out) {
  CHECK_CTX(c);
  int rc = check_slot(c, s, slot);
  if (rc) return rc;
  if (!out) FB_FAIL(c, FB_E_ARG, "fb_frame_pyr_down: null output");
  const size_t npx = (size_t)c->W * c->H, n2 = (size_t)(c->W / 2) * (c->H / 2);
  uint8_t* d = nullptr;
  FB_CUDA(c, dalloc(&d, n2));
  k_pyr_down<<<fb_div_up((int)n2, 256), 256, 0, c->stream>>>(c->W, c->H, c->imgs + ((size_t)s * c->n_slots + slot) * npx, d);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(out, d, n2, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  FB_CUDA(c, e);
  return FB_OK;
}

extern "C" int fb_detect(fb_ctx* c, int s, int slot, int win, int border, float min_grad_mag,
                         const uint8_t* occupied, float* det_xy, int32_t* det_ok, int32_t* n_det) {
  CHECK_CTX(c);
  int rc = check_slot(c, s, slot);
  if (rc) return rc;
  if (win < 4 || win > 64 || border < 0 || !det_xy || !det_ok) FB_FAIL(c, FB_E_ARG, "fb_detect: bad argument");
  const int cells = (c->W / win) * (c->H / win);
  const size_t npx = (size_t)c->W * c->H;
  uint8_t* d_occ = nullptr; float2* d_xy = nullptr; int32_t *d_ok = nullptr, *d_rank = nullptr, *d_cnt = nullptr;
  FB_CUDA(c, dalloc(&d_occ, cells)); FB_CUDA(c, dalloc(&d_xy, cells)); FB_CUDA(c, dalloc(&d_ok, cells));
  FB_CUDA(c, dalloc(&d_rank, cells)); FB_CUDA(c, dalloc(&d_cnt, 1));
  cudaStream_t st = c->stream;
  cudaError_t e = occupied ? cudaMemcpyAsync(d_occ, occupied, cells, cudaMemcpyHostToDevice, st) : cudaMemsetAsync(d_occ, 0, cells, st);
  k_detect_features<<<fb_div_up(cells * 32, 256), 256, 0, st>>>(c->W, c->H, win, border, min_grad_mag, c->imgs + ((size_t)s * c->n_slots + slot) * npx, d_occ, d_xy, d_ok);
  k_scan_flags<<<1, 1024, 0, st>>>(cells, d_ok, d_rank, d_cnt);
  c->launches += 2;
  int32_t cnt = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(det_xy, d_xy, sizeof(float2) * cells, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(det_ok, d_ok, sizeof(int32_t) * cells, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_occ); cudaFree(d_xy); cudaFree(d_ok); cudaFree(d_rank); cudaFree(d_cnt);
  FB_CUDA(c, e);
  if (n_det) *n_det = cnt;
  return FB_OK;
}

// ------------------------------------------------------------------------------------ profiling
extern "C" int fb_profile_enable(fb_ctx* c, int enable) {
  CHECK_CTX(c);
  // enable: 0 = off, 1 = every section, 2 + section = that section only (two event records per call of
  // a section cost ~5 us of host time: a tight loop that only wants the solver's duration asks for it alone)
  c->prof = enable != 0;
  c->prof_mask = enable >= 2 ? (1u << (enable - 2)) : 0xffffffffu;
  return FB_OK;
}

extern "C" int fb_profile_reset(fb_ctx* c) {
  CHECK_CTX(c);
  for (int k = 0; k < FB_PROF_NUM; ++k) {
    prof_fold(c, k);
    c->sec[k].total_ms = 0.0;
    c->sec[k].calls = 0;
    c->sec[k].launches = 0;
  }
  return FB_OK;
}

extern "C" int fb_profile_get(fb_ctx* c, int section, float* total_ms, int64_t* calls, int64_t* launches) {
  CHECK_CTX(c);
  if (section < 0 || section >= FB_PROF_NUM) FB_FAIL(c, FB_E_ARG, "fb_profile_get: bad section");
  prof_fold(c, section);
  if (total_ms) *total_ms = (float)c->sec[section].total_ms;
  if (calls) *calls = c->sec[section].calls;
  if (launches) *launches = c->sec[section].launches;
  return FB_OK;
}

extern "C" int64_t fb_launch_count(const fb_ctx* c) { return c ? c->launches : 0; }
