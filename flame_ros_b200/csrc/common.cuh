// common.cuh -- shared declarations for libflame_b200 (sm_100a).
//
// Float discipline (DESIGN.md "Numerics"): this translation unit is compiled with --fmad=false, so
// a fused multiply-add happens only where fmaf() is written.  Every kernel spells its arithmetic in
// the order documented in DESIGN.md so results are reproducible run to run and comparable with the
// test oracle bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <chrono>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/flame_b200.h"

#define FB_MAX_WIN 15
#define FB_MAX_SEARCH 256
#define FB_GEO_STRIDE 16  // floats per (stream, slot) geometry record: A[9] b[3] e[3] pad

struct ProfSection {
  std::vector<cudaEvent_t> ev;  // begin/end pairs not yet folded into total_ms
  double total_ms = 0.0;
  int64_t calls = 0;
  int64_t launches = 0;
};

struct fb_ctx {
  int device = 0;
  int S = 0, W = 0, H = 0, n_slots = 0, maxF = 0, maxV = 0, maxE = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // pipelined fb_hotpath_step: uploads on a copy stream, per-slot ready/free events, result events
  cudaStream_t copy_stream = nullptr;
  cudaStream_t out_stream = nullptr;     // D2H of the vertex idepths of frame k (off the solve stream's critical path)
  cudaEvent_t ev_solved = nullptr;
  // pipelined step: the data term is double-buffered so that the assembly of frame k+1 (main stream)
  // overlaps the solve of frame k (solve stream); z / wt point at the buffer of the latest assembly
  float* z_buf[2] = {nullptr, nullptr};
  float* wt_buf[2] = {nullptr, nullptr};
  cudaEvent_t ev_zfree[2] = {nullptr, nullptr};  // recorded after the solve that read buffer b
  bool zfree_valid[2] = {false, false};
  int64_t n_pipe_steps = 0;
  float* x_stage[2] = {nullptr, nullptr}; // device snapshots of x the result copies read (the next solve is free to write x)
  cudaStream_t solve_stream = nullptr;   // assembly + solver of frame k run here while `stream` already
                                         // processes the epipolar update of frame k+1
  cudaEvent_t ev_epi = nullptr, ev_asm = nullptr, ev_join = nullptr, ev_join2 = nullptr;
  // pinned staging ring for the small per-frame parameter uploads (poses, slot indices): pageable
  // cudaMemcpyAsync would block the host until the stream reaches the copy and stall the pipeline
  uint8_t* stage = nullptr;
  size_t stage_slot_bytes = 0;
  int stage_next = 0, stage_last = 0;
  std::vector<cudaEvent_t> stage_ev;   // completion of the last copy issued from each slot
  bool pipe_hold = false;                // inside a pipelined step: internal calls must not drain
  bool pipe_dirty = false;               // pipelined work may be in flight on the auxiliary streams
  std::vector<cudaEvent_t> ev_ready, ev_free;
  // landing buffers of the pipelined step: S contiguous frames per buffer, so a step whose host
  // frames are contiguous goes up as ONE linear transfer (8 separate 300 kB copies cost ~12 us each)
  uint8_t* incoming = nullptr;            // [2][S][H*W]
  std::vector<int> slot_landing;          // per slot: -1 = the frame is in its slot, else landing buffer index
  std::vector<char> slot_is_ref;          // per slot: holds a poseframe that features may still refer to
  const uint8_t* epi_cmp_frames = nullptr;  // consumed by the next fb_idepth_update
  cudaEvent_t ev_result[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t n_pipelined = 0;
  std::string err;

  // ---- graph (per stream s: vertex base s*maxV, edge base s*maxE, incidence base s*2*maxE)
  float4* vbar = nullptr;   // (xb, w1b, w2b, 0)         [S*maxV]
  float* x = nullptr;       // planar primal state        [S*maxV] each
  float* w1 = nullptr;
  float* w2 = nullptr;
  float* z = nullptr;       // data term
  float* wt = nullptr;      // data weight
  float4* ec = nullptr;     // (alpha, beta, dx, dy)      [S*maxE]
  int2* eij = nullptr;      // (i, j) stream-local        [S*maxE]
  float4* q4 = nullptr;     // (q1, q2, q3, 0)            [S*maxE]
  int32_t* row = nullptr;   // CSR row pointers           [S*(maxV+1)]
  int32_t* inc = nullptr;   // (edge<<1)|role             [S*2*maxE]
  int32_t* epos = nullptr;  // position of edge e in its TARGET's CSR row [S*maxE] (variant 5's slot address)
  int32_t* vnin = nullptr;  // in-degree of every vertex  [S*maxV]
  int32_t* nV = nullptr;    // device counts              [S]
  int32_t* nE = nullptr;
  std::vector<int32_t> hV, hE;       // host mirrors of the counts
  int32_t* vfeat = nullptr; // vertex -> feature index    [S*maxV]
  float2* vpos = nullptr;   // vertex pixel positions     [S*maxV]
  double* costs = nullptr;  // [S*2]

  // ---- persistent-cluster solver images (variant 2), rebuilt by fb_graph_set
  struct ClusterPlan* plan = nullptr;
  struct GridPlan* gplan = nullptr;   // grid-resident solver tables (variant 3)
  struct UpdateState* upd = nullptr;  // fb_update pipeline state (flame_update.cuh)
  struct TilePlan* tplan = nullptr;   // device-planned tile-resident solver (variant 5, nltgv2_tile.cuh)

  // ---- frames
  uint8_t* imgs = nullptr;  // [S][n_slots][H][W]
  std::vector<float> h_pose;  // [S][n_slots][7]
  std::vector<float> h_K;     // [S][9]
  float* d_pose = nullptr;    // [S][n_slots][7]
  float* d_K = nullptr;       // [S][9]
  int32_t* d_cmp = nullptr;   // [S]
  float* d_geo = nullptr;     // [S][n_slots][FB_GEO_STRIDE]
  uint8_t* pool = nullptr;    // device frame pool
  int pool_n = 0;

  // ---- features (per stream base s*maxF)
  float2* f_uref = nullptr;
  float* f_mu = nullptr;
  float* f_var = nullptr;
  int32_t* f_drop = nullptr;
  int32_t* f_alive = nullptr;
  int32_t* f_ref = nullptr;
  int32_t* f_status = nullptr;
  float2* f_ucmp = nullptr;
  int32_t* nF = nullptr;       // [S]
  std::vector<int32_t> hF;
  int32_t* counters = nullptr; // [S][FB_NUM_COUNTERS]
  fb_epi_params epi;

  // ---- mesh / interpolation
  int32_t* tri = nullptr;      // [S][maxT][3]
  int maxT = 0;
  std::vector<int32_t> hT;
  int32_t* nT = nullptr;
  uint8_t* tri_valid = nullptr;  // [S*maxT]
  int32_t* owner = nullptr;      // [S][H*W] owning triangle per pixel
  float* idmap = nullptr;        // [S][H*W]

  // ---- CUDA-graph cache for the streaming solver: one executable per value of `only` (index
  // only + 1; fb_update on a batch context solves one stream at a time, round robin)
  std::vector<cudaGraphExec_t> solve_exec;   // [2 * (S + 1)]: per `only` and per data-term buffer
  std::vector<const float*> solve_z;
  std::vector<int> solve_iters;
  std::vector<fb_nltgv2_params> solve_params;
  // ---- plan-free resident solver (variant 4, nltgv2_coop.cuh)
  float4* coop_contrib = nullptr;  // [S*2*maxE] K^T q per edge end
  int* coop_err = nullptr;         // mapped host flag
  int coop_cluster = -1;           // cluster size: -1 not probed, 0 unavailable
  int coop_ept = 0, coop_vpt = 0;  // edges / vertices per thread of the instantiation in use
  float* idmap_scratch = nullptr;  // [H*W] filtered maps of getter calls (never the prediction source)
  // fb_update renders the filtered map the getters last asked for beside the unfiltered one (flame_update.cuh)
  float* idmap_f = nullptr;        // [S*H*W]
  int32_t* owner2 = nullptr;       // [S*H*W] second ownership map, kept clean by the shading kernel
  uint64_t mut_epoch = 0;          // counts the entry points that may change what a getter returns
  int last_variant = 0;
  int last_cluster = 0;  // cluster size of the last variant-2 launch / parts per stream of variant 3
  bool grid_disabled = false;  // variant 3 launch refused once by the device: auto stops trying it
  int last_threads = 0;    // variant 3: threads per CTA of the last launch
  int last_transport = 0;  // variant 3: 1 = cluster (DSMEM st.async), 2 = L2 mailboxes
  int cluster_min = 1;  // FB_CLUSTER_MIN env: lower bound on the cluster size of variant 2

  // ---- profiling
  bool prof = false;
  unsigned prof_mask = 0xffffffffu;  // sections that record events while prof is on (bit = section)
  std::vector<cudaEvent_t> prof_free;  // recycled timing events (creating one costs ~2 us of host time)
  ProfSection sec[FB_PROF_NUM];
  int64_t launches = 0;
};

#define FB_CUDA(ctx, call)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                     \
      return FB_E_CUDA;                                                                     \
    }                                                                                       \
  } while (0)

#define FB_FAIL(ctx, code, msg) \
  do {                          \
    (ctx)->err = (msg);         \
    return (code);              \
  } while (0)

static inline int fb_div_up(int a, int b) { return (a + b - 1) / b; }

template <typename T>
static cudaError_t dalloc(T** p, size_t n) {
  return cudaMalloc((void**)p, sizeof(T) * (n ? n : 1));
}

// Brackets a section with CUDA events (when profiling is enabled) and counts calls / launches.
struct ProfScope {
  static cudaEvent_t take(fb_ctx* c) {
    cudaEvent_t e = nullptr;
    if (!c->prof_free.empty()) {
      e = c->prof_free.back();
      c->prof_free.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
  fb_ctx* c;
  int sec;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int64_t l0;
  ProfScope(fb_ctx* ctx, int section) : c(ctx), sec(section), l0(ctx->launches) {
    if (c->prof && ((c->prof_mask >> section) & 1u)) {
      e0 = take(c);
      e1 = take(c);
      cudaEventRecord(e0, c->stream);
    }
  }
  ~ProfScope() {
    ProfSection& s = c->sec[sec];
    s.calls++;
    s.launches += c->launches - l0;
    if (e0) {
      cudaEventRecord(e1, c->stream);
      s.ev.push_back(e0);
      s.ev.push_back(e1);
    }
  }
};

#ifdef DSG_CLOCKS  // -DDSG_CLOCKS: thread 0 prints the cycles spent in every phase (debug builds only)
#define DSG_CLK_DECL long long clk_[12]; int nclk_ = 0;
#define DSG_CLK { if (threadIdx.x == 0 && nclk_ < 12) clk_[nclk_++] = clock64(); }
#define DSG_CLK_PRINT(name) { if (threadIdx.x == 0) { printf("%s cycles:", name); for (int i_ = 1; i_ < nclk_; ++i_) printf(" %lld", clk_[i_] - clk_[i_ - 1]); printf("\n"); } }
#else
#define DSG_CLK_DECL
#define DSG_CLK
#define DSG_CLK_PRINT(name)
#endif
