// Engine state behind the opaque gt_handle.
#pragma once
#include "common.cuh"

struct MaxpoolOp { View in, out; };  // 5x5 stride-1 max pool (SPPF chain)
struct UpsampleOp { View in, out; }; // 2x nearest (only used when not fused in a conv epilogue)

enum OpType { OP_CONV = 1, OP_MAXPOOL = 2 };
struct PlanOp {
  OpType type;
  int conv = -1;      // index into conv_ops
  MaxpoolOp pool;
};

// ORB per-level geometry
struct OrbLevel {
  int w, h;
  size_t off;         // byte offset of this level inside one frame's pyramid slab
  float scale;        // 1.2^level (float, as OpenCV's getScale)
  int quota_cur, quota_ref;
  int cand_cap;       // FAST candidate capacity
  size_t cand_off;    // element offset inside one frame's candidate slab
};

#define GT_ORB_LEVELS 8
#define GT_MAX_KP 8192      // per-frame keypoint capacity (>= max_features * ref_multiplier)

struct OrbSet {             // one frame's final features (device pointers into slabs)
  float* kp = nullptr;      // [GT_MAX_KP][6] x,y,size,angle,response,octave
  uint8_t* desc = nullptr;  // [GT_MAX_KP][32]
  int* count = nullptr;     // [1]
};

// conv kernel variants: 0 pixel-major, 1 swapped, 2 swapped + halo, 3 pixel-major two CTAs / SM, 4 pixel-major + halo, 5 = 3 + 4
#define GT_CONV_VARIANTS 7

struct gt_engine {
  gt_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int tune_mode = 0;                    // GT_TUNE=1: time every conv variant at weight load and print table rows (development); 0: the shipped table / rule decides
  int swap_mode = -1;                   // conv kernel per layer: -1 autotune (time the variants at weight load), 0 pixel-major (conv_tc.cu), 1 swapped (conv_sw.cu), 2 swapped + halo staging where it applies, 3 pixel-major at two CTAs per SM where it applies; GT_SWAP env
  int plan_variant = 0;                 // variant conv_tc_plan builds right now (0 / 1)
  int pdl = 1;                          // programmatic dependent launch between conv layers (GT_PDL=0 disables)
  int pair_mode = 1;                    // 1: the autotuner also times variant 6 (CTA-pair swapped kernel, cout % 256 == 0); GT_PAIR=0 leaves it out
  int halo_mode = 0;                    // conv A-operand staging: 0 per-tap boxes, 1 halo boxes + shifted descriptors (GT_HALO=1 enables: fewer L2->SM bytes, but the layers are tensor-issue bound)
  std::vector<void*> dev_allocs;
  std::vector<void*> host_allocs;

  // geometry
  int net_h = 0, net_w = 0, new_h = 0, new_w = 0, pad_top = 0, pad_left = 0;
  float gain = 1.f;
  int work_h = 0, work_w = 0;
  int A = 0, no = 0;                    // anchors, raw head row width (64 + nc (+1))
  int ncp = 0;                          // class-logit row width (nc rounded up to 4 floats)
  int lvl_h[3], lvl_w[3], lvl_off[3];

  // staging + stage 1
  uint8_t* frames_dev = nullptr;        // [B][H][W][3] staging buffer 0 (host inputs)
  uint8_t* frames_dev2 = nullptr;       // staging buffer 1: gt_prefetch_frames copies batch i+1 while batch i computes
  cudaStream_t copy_stream = nullptr;
  cudaStream_t aux_stream = nullptr;    // low priority: the mask-independent half of ORB runs here next to the detector
  cudaStream_t aux2_stream = nullptr;   // the image pyramid, beside FAST on level 0 (orb_front_split)
  cudaEvent_t ev_aux_a = nullptr, ev_aux_b = nullptr;
  cudaEvent_t ev_pre = nullptr, ev_front = nullptr;
  int overlap = 2;                      // GT_OVERLAP: 2 ORB front on the aux stream beside decode + NMS; 3 as 2, but the image pyramid already beside the conv stack (fills the tails between layers); 1 the whole front beside the detector (no gain: the conv CTAs own the SMs); 0 serial
  int l2promo_128 = 0;                  // GT_L2PROMO=128: 128-byte L2 promotion for every activation tensor map (the round-1 setting)
  int nms_fused = 1;                    // GT_NMS_FUSED=0: always the multi-launch sort / gather / mask / sweep path
  int chain_mode = 1;                   // GT_CHAIN=0: model.1 and model.2.cv1 as two launches instead of one chained kernel
  int match_mode = 2;                   // GT_MATCH: 2 Hamming 2-NN as E4M3 tcgen05 GEMM (match_tc.cu), 1 the same with fp16 operands, 0 POPC kernel
  int mask_sparse = 1;                  // GT_MASK_SPARSE=0: dense mask pyramid (7 full-plane launches) instead of the box-driven sparse one
  long long silu_tanh_px = 1;           // GT_SILU_TANH_PX: layers with at least this many output pixels per image use the one-MUFU SiLU (default 1 = every SiLU layer: measured raw-head error unchanged at 5.8e-3, conv stack -3 %; 0 = none)
  int front_split = 0;                  // GT_FRONT_SPLIT=1: FAST on level 0 runs beside the seven pyramid launches (two aux streams) -- measured neutral (5.013 vs 5.020 ms / step), off by default
  int conv_smem_kb = 227;               // dynamic smem budget of the conv kernels (200 with GT_OVERLAP=1 to leave room for ORB blocks); GT_CONV_SMEM_KB
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};   // H2D of staging buffer k finished (copy stream)
  cudaEvent_t ev_consumed[2] = {nullptr, nullptr}; // the preprocess kernel that read staging buffer k finished
  const void* prefetched_src[2] = {nullptr, nullptr};
  int prefetch_next = 0;
  const uint8_t* deferred_src = nullptr;   // gt_prefetch_frames_deferred: started by the next gt_extract_batch
  int deferred_B = 0;
  int input_format = 0;                 // GT_INPUT_BGR24 | GT_INPUT_NV12 (gt_set_input_format)
  uint8_t* frames_bgr = nullptr;        // NV12 ingest: device BGR24 frames produced by nv12_to_bgr_kernel (allocated on first use)
  bool lb_fast = true;                  // exact-1/2 letterbox with 16-pixel-aligned rows: the vector kernel writes the network input (pre_fast: and the gray working image)
  bool pre_fast = true;                 // exact-1/2 letterbox + 1/2 working image with 16-pixel-aligned rows: the fused vector kernel; else the table-driven general kernels
  int* lb_tab[8] = {};                  // letterbox resize tables (x0, x1, a0, a1, y0, y1, b0, b1), see detector.cu
  int* gw_tab[8] = {};                  // gray working-image resize tables
  int lb_mode = 0, gw_mode = 0;         // 0 identity, 1 exact 2x2 decimation, 2 bilinear
  bf16* net_s2d = nullptr;              // [B][net_h/4][net_w/4][64] 4x4 space-to-depth letterboxed RGB0 (exact u8 values, 16-bit)
  const uint8_t* cur_frames = nullptr;  // device pointer of the frames of the last gt_preprocess
  // CLAHE front end (cfg.clahe; clahe.cu): full-resolution gray plane, per-tile histograms and LUTs, interpolation tables
  uint8_t* gray_full = nullptr;         // [B][H][W]
  uint8_t* gray_eq = nullptr;           // [B][H][W] equalised plane (only when the working image is smaller than the frame)
  int* clahe_hist = nullptr;            // [B][64][256]
  uint8_t* clahe_lut = nullptr;         // [B][64][256]
  int clahe_tw = 0, clahe_th = 0, clahe_clip = 0;
  float clahe_scale = 0.f;
  int* clahe_xi[2] = {}; float* clahe_xa[2] = {};   // per column: tile indices (left, right), weights (xa, 1 - xa)
  int* clahe_yi[2] = {}; float* clahe_ya[2] = {};   // per row

  // detector
  bool weights_loaded = false;
  bool tuned = false;
  std::vector<gt_conv_desc> conv_descs;             // canonical list
  std::vector<ConvOp> conv_ops;                     // fused tcgen05 ops (the variant in use)
  std::vector<ConvOp> conv_var[GT_CONV_VARIANTS];   // the other variants of each op (autotune), same indexing as conv_ops; empty when a variant is forced
  std::vector<char> conv_var_ok[GT_CONV_VARIANTS];  // 1 where the variant applies to the op
  std::vector<ConvSig> conv_sig;                    // GT_TUNE=1: signature of each op (table key)
  std::vector<PlanOp> plan;
  int conv0_op = -1;                                // index of layer 0 in conv_ops (custom weight packing)
  View feat_views[23];
  float* raw_box = nullptr;                         // [B][A][64] DFL logits (f32 rows written by the cv2.x.2 convs)
  float* raw_cls = nullptr;                         // [B][A][ncp] class logits (cv3.x.2)
  float* raw_ang = nullptr;                         // [B][A][4] OBB angle logit in column 0 (cv4.x.2)
  // decode + NMS workspaces
  int cand_cap = 0;                                 // candidates per image
  float* cand_box = nullptr;                        // [B][cand_cap][5] x1,y1,x2,y2 (or x,y,w,h) + angle
  unsigned long long* cand_key = nullptr;           // [B][cand_cap_pow2] sort keys
  float* cand_conf = nullptr;
  int* cand_cls = nullptr;
  int* cand_anchor = nullptr;
  int* cand_count = nullptr;                        // [B]
  int* nonfinite_dev = nullptr;                     // [1] cumulative count of anchors whose head row held inf / NaN (16-bit overflow guard)
  int* nonfinite_host = nullptr;                    // pinned mirror, refreshed by every decode (valid after the call's sync / gt_wait)
  uint32_t cls_filter[3] = {0, 0, 0};               // gt_set_class_filter: allow-list for classes 0..95
  bool cls_filter_on = false;
  unsigned long long* nms_mask = nullptr;
  int nms_cap = 0;
  float* det_out = nullptr;                         // [B][max_det][7]
  int* det_count = nullptr;                         // [B]
  int* det_keep = nullptr;                          // [B][max_det]
  float* pred_tmp = nullptr; size_t pred_tmp_bytes = 0;

  // stage 3
  OrbLevel lv[GT_ORB_LEVELS];
  size_t pyr_bytes = 0;                             // one frame's pyramid slab
  size_t cand_total = 0;
  uint8_t* pyr = nullptr;                           // [B+1][pyr_bytes]  (slot B = reference)
  uint8_t* pyr_mask = nullptr;                      // same geometry
  unsigned int* fast_cand = nullptr;                // [B+1][cand_total] packed (y<<16|x)
  uint8_t* fast_score = nullptr;                    // [B+1][cand_total]
  int* fast_count = nullptr;                        // [B+1][8]
  unsigned int* sel_xy = nullptr;                   // [B+1][8][sel_cap]
  float* sel_resp = nullptr;
  int* sel_count = nullptr;                         // [B+1][8]
  int sel_cap = 0;
  float* kp_all = nullptr;                          // [B+1][GT_MAX_KP][6]
  uint8_t* desc_all = nullptr;                      // [B+1][GT_MAX_KP][32]
  int* kp_count = nullptr;                          // [B+1]
  int* lvl_kp_off = nullptr;                        // [B+1][9]
  float* boxes_dev = nullptr;                       // [B+1][max_det][4]
  int* nboxes_dev = nullptr;                        // [B+1]
  float* det_xywh_dev = nullptr;                    // [B][max_det][4] detections as xywh (warp input)
  int* det_nbox_dev = nullptr;                      // [B]
  float* box_stage[2] = {nullptr, nullptr};         // mapped pinned staging for host-side mask boxes (see upload_boxes)
  int* nbox_stage[2] = {nullptr, nullptr};
  int box_stage_next = 0;
  bool have_ref = false;
  OrbLevel* lv_dev = nullptr;                       // device copy of lv[]
  int* rs_tab[GT_ORB_LEVELS][4] = {};               // per-level resize tables: xofs, xc1, yofs, yc1
  // matching / RANSAC
  uint8_t* desc_x = nullptr;                        // [B+1][64 groups][k-blocks][128][128 B] descriptors expanded to MMA operand tiles (match_tc.cu)
  float* desc_c = nullptr;                          // [B+1][GT_MAX_KP] popc * 8192 + row (float), huge beyond the count
  // registration (f-3): L2 matcher work buffers, grown on demand by match_l2_run / gt_match_l2 and freed by gt_destroy
  int reg_cap = 0, reg_io_cap = 0;
  uint8_t* reg_xq = nullptr; uint8_t* reg_xt = nullptr;   // fp16 operand images [cap / 128][2][128][128 B]
  float* reg_cq = nullptr; float* reg_ct = nullptr;         // |d|^2 per row
  int* reg_cand = nullptr; int* reg_n = nullptr;            // [cap][4] tensor-core candidates; the two counts
  float* reg_q32 = nullptr; float* reg_t32 = nullptr;       // device copies of host descriptors [io_cap][128]
  int* reg_idx = nullptr; float* reg_dist = nullptr;        // [io_cap][2]
  int* match_idx = nullptr;                         // [B][GT_MAX_KP][2]
  int* match_dist = nullptr;                        // [B][GT_MAX_KP][2]
  float* pairs = nullptr;                           // [B][GT_MAX_KP][4] cur x,y, ref x,y (working res)
  int* pair_count = nullptr;                        // [B]
  float* npairs = nullptr;                          // [B][GT_MAX_KP][4] Hartley-normalised pairs
  float* norms = nullptr;                           // [B][8] normalisation parameters
  float* hyp_score = nullptr;                       // [B][max_iter]
  int* hyp_hist = nullptr;                          // [B][256] histogram of the subset scores (preemptive RANSAC scoring)
  double* H_dev = nullptr;                          // [B][9]
  int* H_status = nullptr;                          // [B]
  int* H_stats = nullptr;                           // [B][4]
  float* boxes_stab_dev = nullptr;                  // [B][max_det][4]
  double* warp_minv = nullptr;                      // [B][9] inverse transforms of gt_warp_frames
  uint8_t* warp_out = nullptr;                      // [B][H][W][3] device staging of gt_warp_frames for host destinations (allocated on first use)

  // timing
  cudaEvent_t ev_sets[2][8];            // two sets of stage-timing events: gt_extract_batch_async alternates between them
  cudaEvent_t* ev = ev_sets[0];         // the set of the call being enqueued
  cudaEvent_t ev_done[2] = {nullptr, nullptr};   // all work and read-backs of ticket k are complete
  int async_ticket = 0;
  bool ticket_pending[2] = {false, false};   // ticket issued by gt_extract_batch_async and not yet passed to gt_wait
  bool post_event_done = false;         // nms_run recorded ev[4] right behind the one-launch NMS (before its host-side look at the overflow flags)
  float stage_ms[4] = {0, 0, 0, 0};
  float conv_ms = 0;
  double conv_flops = 0;

  int dev_alloc(void** p, size_t bytes);
  int host_alloc(void** p, size_t bytes);
};

// detector.cu
int detector_build(gt_engine* e);
int detector_load_weights(gt_engine* e, const float* const* w, const float* const* b, int n);
int detector_autotune(gt_engine* e, cudaStream_t st);
int detector_fill_pad(gt_engine* e, cudaStream_t st);
int detector_build_general_preprocess(gt_engine* e);
int detector_preprocess_nv12(gt_engine* e, const uint8_t* nv12_dev, int B, cudaStream_t st);   // fused (default geometry only)
int detector_nv12_to_bgr(gt_engine* e, const uint8_t* nv12_dev, uint8_t* bgr_dev, int B, cudaStream_t st);
int detector_preprocess(gt_engine* e, const uint8_t* frames_dev, int B, cudaStream_t st);
int detector_forward(gt_engine* e, int B, cudaStream_t st);
int detector_postprocess(gt_engine* e, int B, float conf, float iou, int agnostic, uint32_t classes_mask, cudaStream_t st);
int nms_run(gt_engine* e, const float* pred_dev, int B, int A, int nc, int rotated, float conf, float iou, int agnostic,
            uint32_t classes_mask, int max_det, bool scale_to_frame, cudaStream_t st);

// warp.cu
int warp_frames_run(gt_engine* e, const uint8_t* src_dev, uint8_t* dst_dev, const double* H_host, int B, cudaStream_t st);

// clahe.cu
int clahe_build(gt_engine* e);
int clahe_run(gt_engine* e, const uint8_t* frames_dev, int B, cudaStream_t st);

// orb.cu
int orb_build(gt_engine* e);
int orb_run(gt_engine* e, int slot0, int nslots, bool as_reference, bool build_mask, cudaStream_t st);
int orb_front(gt_engine* e, int slot0, int nslots, cudaStream_t st);   // image pyramid + FAST: independent of the mask
int orb_pyramid(gt_engine* e, int slot0, int nslots, cudaStream_t st);
int orb_fast(gt_engine* e, int slot0, int nslots, cudaStream_t st, int level_lo, int level_hi, bool reset);
int orb_front_split(gt_engine* e, int slot0, int nslots, cudaStream_t st, cudaStream_t st2, cudaEvent_t ev_a, cudaEvent_t ev_b);
int orb_mask(gt_engine* e, int slot0, int nslots, bool build_mask, cudaStream_t st);   // vehicle mask + its pyramid: independent of orb_front
int orb_back(gt_engine* e, int slot0, int nslots, bool as_reference, cudaStream_t st);
// match_ransac.cu
int stab_build(gt_engine* e);
int stab_match_and_fit(gt_engine* e, int B, cudaStream_t st);
int match_run(gt_engine* e, const uint8_t* q, const int* nq_dev, int nq_max, const uint8_t* t, const int* nt_dev, int nt_max,
              int* out_idx, int* out_dist, int batch, size_t q_stride, size_t out_stride, cudaStream_t st);
// match_tc.cu
int match_tc_build(gt_engine* e);
int match_tc_run(gt_engine* e, int q_slot0, int q_step, int nq_cap, int t_slot0, int t_step, int batch, cudaStream_t st);
int match_l2_run(gt_engine* e, const float* q_dev, int nq, const float* t_dev, int nt, int* out_idx_dev, float* out_dist_dev, cudaStream_t st);
int homography_run(gt_engine* e, const float* pairs, const int* counts, int B, int pair_stride, float thr, int max_iter,
                   double* out_H, int* out_status, int* out_stats, float ratio, bool full_res, const int* kp_count, cudaStream_t st);
int warp_boxes_run(gt_engine* e, const double* H_dev, const int* status_dev, const float* in, float* out, const int* counts, int B,
                   int stride, cudaStream_t st);
