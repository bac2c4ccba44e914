// flame/params.h -- flame::Params with exactly the fields the reference frontends assign
// (/root/reference/src/flame_nodelet.cc:172-263; defaults /root/reference/cfg/flame_nodelet.yaml).
#pragma once

namespace flame {

struct FeatureParams {      // params.fparams (/root/reference/src/flame_nodelet.cc:227,238)
  float min_grad_mag = 5.0f;
  int win_size = 5;
};
struct MeasModelParams {    // params.zparams (/root/reference/src/flame_nodelet.cc:237,245)
  int win_size = 5;
  float epipolar_line_var = 4.0f;
};
struct RegularizerParams {  // params.rparams (/root/reference/src/flame_nodelet.cc:256-259)
  float data_factor = 0.15f;
  float step_x = 0.001f;
  float step_q = 125.0f;
  float theta = 0.25f;
  float x_min = 0.0f;   // not set by the reference frontends; box on inverse depth
  float x_max = 10.0f;
};

struct Params {
  // output / debug (/root/reference/src/flame_nodelet.cc:172-219)
  bool debug_quiet = false;
  float scene_color_scale = 1.0f;
  bool do_oblique_triangle_filter = true;
  float oblique_normal_thresh = 1.57f;
  float oblique_idepth_diff_factor = 0.35f;
  float oblique_idepth_diff_abs = 0.1f;
  bool do_edge_length_filter = true;
  float edge_length_thresh = 0.333f;
  bool do_idepth_triangle_filter = true;
  float min_triangle_idepth = 0.01f;
  bool debug_draw_wireframe = false;
  bool debug_draw_features = false;
  bool debug_draw_detections = false;
  bool debug_draw_matches = false;
  bool debug_draw_normals = false;
  bool debug_draw_idepthmap = false;
  bool debug_draw_text_overlay = false;
  bool debug_flip_images = false;
  // threading (/root/reference/src/flame_nodelet.cc:221-222): accepted, unused -- work runs on the GPU
  int omp_num_threads = 4;
  int omp_chunk_size = 1024;
  // features (/root/reference/src/flame_nodelet.cc:225-245)
  bool do_letterbox = false;
  float min_grad_mag = 5.0f;
  float min_error = 100.0f;
  int detection_win_size = 16;
  int max_dropouts = 5;
  FeatureParams fparams;
  MeasModelParams zparams;
  // regularisation (/root/reference/src/flame_nodelet.cc:248-263)
  bool do_nltgv2 = true;
  bool adaptive_data_weights = false;
  bool rescale_data = false;
  bool init_with_prediction = true;
  float idepth_var_max_graph = 0.01f;
  RegularizerParams rparams;
  float min_height = -1e14f;
  float max_height = 1e14f;
  bool check_sticky_obstacles = false;
  // not exposed by the reference frontends
  int nltgv2_iters = 50;          // primal-dual iterations per frame
  int num_poseframes = 7;         // poseframe ring size
  int max_features = 8192;
  int device = 0;                 // CUDA device index
};

}  // namespace flame
