/*
 * geotrax_b200.h -- C-ABI of the B200-native geo-trax extraction hot path.
 *
 * Drop-in boundary for the per-frame loop of /root/reference/geotrax/extract.py:134-214.  The reference reaches
 * this arithmetic through two Python objects:
 *     ultralytics.YOLO(...).track(frame, **cfg)            extract.py:153, 217-236
 *     stabilo.Stabilizer(**cfg).set_ref_frame/stabilize/   extract.py:139, 177-187; utils/registration.py:59-85
 *         transform_cur_boxes/get_cur_trans_matrix
 * The Python shims in geo-trax_b200/ keep those two surfaces and call the entry points below through ctypes.
 *
 * Conventions: extern "C"; every call returns 0 on success or a negative gt_status; no exceptions cross the
 * boundary; gt_last_error() gives the text.  Pointers are plain caller-owned buffers.  Unless a parameter is
 * documented "device", it may be either a host pointer or a CUDA device pointer (detected with
 * cudaPointerGetAttributes); host inputs are staged through the library's own pinned ring.  `stream` is a
 * cudaStream_t passed as void* (NULL = the handle's own stream).  One handle per GPU; a handle is not
 * thread-safe; calls on one handle are serialised on its stream.  The library owns all device workspaces.
 */
#ifndef GEOTRAX_B200_H_
#define GEOTRAX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GT_ABI_VERSION 1

typedef struct gt_engine* gt_handle;

typedef enum {
  GT_OK = 0,
  GT_ERR_INVALID = -1,  /* bad argument / unsupported configuration            */
  GT_ERR_CUDA = -2,     /* a CUDA runtime / driver call failed                 */
  GT_ERR_STATE = -3,    /* call order violated (e.g. detect before weights)    */
  GT_ERR_NOMEM = -4
} gt_status;

enum { GT_TASK_DETECT = 0, GT_TASK_OBB = 1 };
enum { GT_ACT_BF16 = 0, GT_ACT_FP16 = 1 };

/* Mirrors the YAML keys the reference splats into the two objects:
 *   ultralytics: block  /root/reference/geotrax/cfg/default.yaml:229-250  (imgsz, conf, iou, max_det, classes, agnostic_nms)
 *   stabilo: block      /root/reference/geotrax/cfg/default.yaml:103-145                                               */
typedef struct gt_config {
  int32_t abi_version;         /* GT_ABI_VERSION */
  /* frame geometry */
  int32_t frame_h, frame_w;    /* source frame, u8 BGR HWC (2160 x 3840) */
  int32_t max_batch;           /* frames per call (<= 32) */
  /* detector */
  int32_t imgsz;               /* 1920 */
  int32_t nc;                  /* number of classes (4) */
  int32_t task;                /* GT_TASK_DETECT | GT_TASK_OBB */
  int32_t max_det;             /* 1000 */
  int32_t max_nms;             /* 30000 */
  /* stabilizer */
  float   downsample_ratio;    /* 0.5 (only 0.5 and 1.0 are implemented) */
  int32_t max_features;        /* 2000 */
  float   ref_multiplier;      /* 2.0 */
  int32_t mask_use;            /* 1 */
  float   mask_margin_ratio;   /* 0.15 */
  float   filter_ratio;        /* 0.9  (Lowe) */
  float   ransac_threshold;    /* 2.0 px, in the working (downsampled) image */
  int32_t ransac_max_iter;     /* 5000 hypotheses */
  int32_t query_is_current;    /* 1: knnMatch(query=current, train=reference) */
  int32_t ransac_full_res;     /* 0: threshold applies at working resolution, H conjugated afterwards */
  uint32_t seed;               /* RANSAC sampling seed (deterministic) */
  int32_t act_dtype;           /* 16-bit storage format of activations + weights: GT_ACT_BF16 | GT_ACT_FP16 (f32 accumulate) */
  int32_t clahe;               /* 1: CLAHE (clip 2.0, 8x8 tiles, as cv2.createCLAHE) on the gray frame before the working-image resize
                                  -- stabilo's `clahe: true` (/root/reference/geotrax/cfg/stable.yaml:115); 0 = default preset */
  int32_t reserved[6];
} gt_config;

void gt_default_config(gt_config* cfg);

int gt_create(const gt_config* cfg, int device, gt_handle* out);
int gt_destroy(gt_handle h);
const char* gt_last_error(gt_handle h);     /* h may be NULL: last creation error */
int gt_abi_version(void);

/* ---- detector weights -----------------------------------------------------------------------------------------
 * Host passes BN-folded convolutions in the canonical order returned by gt_conv_count()/gt_conv_info():
 * weight[i] is f32 [cout][cin][k][k] (PyTorch layout), bias[i] f32 [cout].  The library converts to act_dtype,
 * reorders to [cout][k*k][cin_pad] and uploads.  Replaces ultralytics' AutoBackend + Model.fuse() (extract.py:222). */
typedef struct gt_conv_desc {
  char    name[48];            /* ultralytics state_dict prefix, e.g. "model.2.m.0.cv1" */
  int32_t cin, cout, k, stride;
  int32_t act;                 /* 1 = SiLU, 0 = linear (final head convs) */
} gt_conv_desc;

int gt_conv_count(gt_handle h);
int gt_conv_info(gt_handle h, int idx, gt_conv_desc* out);
int gt_load_weights(gt_handle h, const float* const* weights, const float* const* biases, int n_convs);

/* ---- stage 1: letterbox / normalise (+ half-res gray for stage 3) ---------------------------------------------
 * frames: u8 [B][frame_h][frame_w][3] BGR.  Replaces LetterBox + BasePredictor.preprocess (extract.py:153) and
 * stabilo's BGR2GRAY + resize front end (extract.py:177,181).  Results stay in the handle's workspaces.        */
int gt_preprocess(gt_handle h, const uint8_t* frames, int B, void* stream);
/* optional ingest pipelining (replaces nothing in the reference, whose cv2.VideoCapture loop is synchronous, extract.py:146):
 * starts the H2D copy of a (pinned) host batch on the library's copy stream; the next gt_preprocess / gt_extract_batch
 * called with the same pointer consumes it.  Two staging buffers: copy of batch i+1 overlaps compute of batch i.      */
int gt_prefetch_frames(gt_handle h, const uint8_t* frames, int B);
/* same, started by the next gt_extract_batch right after it has queued its own small inputs (mask boxes) */
int gt_prefetch_frames_deferred(gt_handle h, const uint8_t* frames, int B);
/* Decoder-format ingest (what NVDEC / any H.26x decoder emits, half the bytes of BGR24 over PCIe).  After
 * gt_set_input_format(h, GT_INPUT_NV12) every `frames` argument of gt_preprocess / gt_prefetch_frames* / gt_extract_batch is
 * u8 NV12: [B][frame_h * 3 / 2][frame_w] (Y plane, then interleaved U,V at half resolution); frame_h and frame_w must be even.
 * The frames are converted on the device to the BGR24 the rest of the path (and the reference, extract.py:146 reader.read())
 * works on, with OpenCV's cvtColor(COLOR_YUV2BGR_NV12) integer arithmetic (ITU-R BT.601 limited range, 20-bit fixed point).   */
enum { GT_INPUT_BGR24 = 0, GT_INPUT_NV12 = 1 };
int gt_set_input_format(gt_handle h, int format);
/* ---- NVDEC ingest: H.264 / HEVC Annex-B elementary stream -> dense NV12 frames in HBM -------------------------
 * Replaces `reader.read()` (/root/reference/geotrax/extract.py:146, 248: FFmpeg software decode + swscale on the host) for callers
 * that hold the bitstream: the frames never cross PCIe uncompressed.  The decoder belongs to an engine (same GPU, same frame size: the
 * stream's display size must equal frame_w x frame_h, 8-bit 4:2:0).  libnvcuvid.so.1 (driver library) is dlopen()ed; where it is
 * missing gt_nvdec_available() returns 0 and gt_decoder_create fails -- there is no software decoder behind this.
 *   feed : any number of bytes of the stream (NULL / 0 = end of stream); every picture completed by them is decoded and appended, in
 *          display order, to a ring of `capacity_frames` (>= 2 * max_batch) dense NV12 frames [H * 3 / 2][W] in device memory
 *   take : pointer to up to max_frames consecutive decoded frames + their number; valid until the ring wraps onto them (i.e. until
 *          capacity_frames further frames have been decoded).  After gt_set_input_format(h, GT_INPUT_NV12) that pointer is a `frames`
 *          argument of gt_preprocess / gt_extract_batch.                                                                              */
typedef struct gt_decoder* gt_decoder_handle;
enum { GT_CODEC_H264 = 0, GT_CODEC_HEVC = 1 };
int gt_nvdec_available(void);
int gt_decoder_create(gt_handle h, int codec, int capacity_frames, gt_decoder_handle* out);
int gt_decoder_destroy(gt_decoder_handle d);
int gt_decoder_feed(gt_decoder_handle d, const uint8_t* data, size_t size);
int gt_decoder_pending(gt_decoder_handle d);
int gt_decoder_take(gt_decoder_handle d, int max_frames, const uint8_t** dev_nv12, int32_t* n_frames);
const char* gt_decoder_last_error(gt_decoder_handle d);
/* debug/parity read-back: letterboxed planar RGB u8 [B][3][net_h][net_w] (the 1/255 scale is folded into layer 0's
 * f32 weights, so the network input is exact) and u8 gray [B][work_h][work_w] */
int gt_get_net_input(gt_handle h, int B, uint8_t* out_u8, int32_t* net_h, int32_t* net_w);
int gt_get_gray(gt_handle h, int B, uint8_t* out, int32_t* work_h, int32_t* work_w);

/* ---- stage 2: YOLOv8s forward + DFL decode + confidence filter + NMS ------------------------------------------
 * Consumes the tensor left by gt_preprocess.  out_boxes: f32 [B][max_det][6] = x1,y1,x2,y2,conf,cls in source
 * frame pixels (task OBB: [B][max_det][7] = x,y,w,h,r,conf,cls).  out_keep: optional int32 [B][max_det] anchor index
 * of each kept box.  classes_mask: bit c set = class c allowed (0 = all).  Replaces model.track(...)'s
 * inference + non_max_suppression + scale_boxes (extract.py:153).                                             */
int gt_detect(gt_handle h, int B, float conf, float iou, int agnostic, uint32_t classes_mask,
              float* out_boxes, int32_t* out_counts, int32_t* out_keep, void* stream);
/* classes >= 32 (gt_create accepts nc <= 80; `classes: int | list[int]`, /root/reference/geotrax/cfg/default.yaml:243): a sticky
 * allow-list used by gt_detect / gt_nms / gt_extract_batch whenever their per-call `classes_mask` is 0.  classes == NULL switches
 * the filter off (every class passes); n == 0 with a non-NULL pointer filters every class out (ultralytics' `classes=[]`).       */
int gt_set_class_filter(gt_handle h, const int32_t* classes, int n);
/* 16-bit overflow guard: cumulative number of anchors whose head row held inf / NaN since gt_create (fp16 activations that left
 * the format's range reach the head as non-finite values and are dropped from the candidates).  Valid after the call that ran the
 * detector has synchronised (gt_detect / gt_extract_batch return, or gt_wait).  0 on a healthy engine.                           */
int gt_get_health(gt_handle h, int64_t* nonfinite_rows);
/* candidates per frame that passed the confidence / class filter in the last decode (before NMS; capped by max_nms inside NMS) */
int gt_get_candidate_counts(gt_handle h, int B, int32_t* out_counts);
/* raw head tensor f32 [B][A][no] (A = anchors, no = 64+nc(+1)), anchor-major; parity gate (1) */
int gt_get_raw_head(gt_handle h, int B, float* out, int32_t* A, int32_t* no);
/* any intermediate feature map by ultralytics layer index (0..21): act_dtype NHWC as uint16 bit patterns */
int gt_get_feature(gt_handle h, int layer, int B, uint16_t* out, int32_t* C, int32_t* H, int32_t* W);

/* stand-alone NMS on caller-supplied decoded predictions (for bit-exact index parity):
 * pred f32 [B][A][4+nc(+1)] xywh(+angle last) + class probabilities, letterboxed pixels.                       */
int gt_nms(gt_handle h, const float* pred, int B, int A, int nc, int rotated, float conf, float iou, int agnostic,
           uint32_t classes_mask, int max_det, float* out_rows /*[B][max_det][6|7] letterboxed*/, int32_t* out_counts,
           int32_t* out_keep, void* stream);

/* stand-alone convolution (unit parity of the tcgen05 implicit-GEMM kernel against torch.conv2d):
  * x act_dtype NHWC [B][H][W][cin] as uint16 bits, w f32 [cout][cin][k][k], bias f32, optional residual NHWC;
 * out act_dtype NHWC [B][Ho][Wo][cout] (out_f32 != 0: f32).                                                       */
int gt_conv2d(gt_handle h, const uint16_t* x, int B, int H, int W, int cin, const float* w, const float* bias,
              int cout, int k, int stride, int act, const uint16_t* residual, void* out, int out_f32, void* stream);

/* ---- stage 3: Stabilo-style frame-to-reference homography -------------------------------------------------------
 * Works on the half-res gray left by gt_preprocess.  boxes: f32 xywh in source-frame pixels (the vehicle mask),
 * nboxes[b] per frame, box_stride boxes reserved per frame.                                                    */
int gt_set_reference(gt_handle h, int frame_slot, const float* boxes, int nboxes, void* stream);
/* out_H f64 [B][9] row-major current->reference in source-frame pixels, h33 = 1;
 * out_status int32 [B]: 0 ok, 1 no homography;  out_stats int32 [B][4] = kp_ref, kp_cur, matches, inliers.       */
int gt_stabilize(gt_handle h, int B, const float* boxes, const int32_t* nboxes, int box_stride,
                 double* out_H, int32_t* out_status, int32_t* out_stats, void* stream);
/* boxes xywh (n,4) warped in place: 4 corners -> H -> axis-aligned envelope -> xywh (extract.py:183)            */
int gt_warp_boxes(gt_handle h, const double* H, float* boxes, int n, void* stream);

/* frames u8 BGR [B][frame_h][frame_w][3] warped into the reference frame: out = cv2.warpPerspective(frame, H, (w, h)) with INTER_LINEAR and
 * a constant 0 border, bit for bit -- what the reference's visualisation does with every frame and its transform
 * (/root/reference/geotrax/visualize.py:285-289).  H f64 [B][9] row-major (host), frames / out host or device.                          */
int gt_warp_frames(gt_handle h, const uint8_t* frames, const double* H, int B, uint8_t* out, void* stream);

/* ORB stage read-backs for stage-wise parity against OpenCV.  which: 0 = current batch slot b, 1 = reference.   */
int gt_orb_level_info(gt_handle h, int level, int32_t* w, int32_t* hgt, int32_t* quota_cur, int32_t* quota_ref);
int gt_get_pyramid_level(gt_handle h, int which, int b, int level, uint8_t* out_img, uint8_t* out_mask);
/* keypoints f32 [n][6] = x, y (level-0 pixel units), size, angle(deg), response, octave; descriptors u8 [n][32]  */
int gt_get_keypoints(gt_handle h, int which, int b, int max_n, float* out_kp, uint8_t* out_desc, int32_t* n);
/* FAST candidates of one level after 3x3 NMS + border filter (the vehicle mask is applied at selection): (y << 16 | x), score */
int gt_orb_get_candidates(gt_handle h, int which, int b, int level, int max_n, uint32_t* out_xy, uint8_t* out_score, int32_t* n);
/* run ORB alone on a caller-supplied gray image [B][work_h][work_w] (+ optional masks), into current slots      */
int gt_orb_detect(gt_handle h, const uint8_t* gray, const uint8_t* mask, int B, int as_reference, void* stream);
/* stand-alone 2-NN Hamming matcher + ratio test: query [nq][32], train [nt][32] -> per query best/second index/dist */
int gt_match(gt_handle h, const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* out_idx /*[nq][2]*/,
             int32_t* out_dist /*[nq][2]*/, void* stream);
/* f-3 (registration, /root/reference/geotrax/utils/registration.py:57-93 -> stabilo BFMatcher(NORM_L2).knnMatch(k=2) on RootSIFT
 * descriptors): brute-force L2 2-NN of float descriptors, query f32 [nq][dim], train f32 [nt][dim], dim == 128, host or device
 * pointers; out_idx i32 [nq][2] (-1 where the train set has fewer rows), out_dist f32 [nq][2] = L2 distance.  Tensor-core candidate
 * pass (fp16 operands, four candidates per query) + exact fp32 re-rank; equal distances -> lower train index.               */
#define GT_MATCH_L2_MAX 262144
int gt_match_l2(gt_handle h, const float* query, int nq, const float* train, int nt, int dim, int32_t* out_idx /*[nq][2]*/,
                float* out_dist /*[nq][2]*/, void* stream);
/* stand-alone robust homography: pts f32 [n][2] each (n <= max_batch * 8192) -> H f64[9], inlier count              */
int gt_find_homography(gt_handle h, const float* src, const float* dst, int n, float thr, int max_iter,
                       double* out_H, int32_t* out_inliers, void* stream);

/* ---- fused per-batch call used by the shims and bench.py --------------------------------------------------------
 * preprocess -> detect -> stabilize -> warp the detections.  out_boxes_stab f32 [B][max_det][4] xywh.
 * mask_boxes (xywh, source-frame pixels, [B][mask_stride][4], mask_nboxes[B]) are the vehicle boxes masked out of the
 * ORB input -- in the reference they are the tracker's boxes of the same frame (extract.py:166,181); NULL = use this
 * call's own detections.  first_is_reference != 0: frame 0 of this batch becomes the reference.                 */
int gt_extract_batch(gt_handle h, const uint8_t* frames, int B, int first_is_reference, float conf, float iou,
                     int agnostic, uint32_t classes_mask, const float* mask_boxes, const int32_t* mask_nboxes,
                     int mask_stride, float* out_boxes, int32_t* out_counts,
                     float* out_boxes_stab, double* out_H, int32_t* out_status, int32_t* out_stats, void* stream);

/* Pipelined form of gt_extract_batch: enqueues the kernels and the read-backs (the out_* buffers must be pinned host memory for the
 * read-backs to be asynchronous) and returns a ticket (0 / 1) immediately; gt_wait(ticket) blocks until that batch's outputs are
 * valid.  At most two tickets in flight (a third call first waits for the ticket it reuses).  Lets the caller enqueue batch i+1
 * before reading batch i, so the GPU does not idle
 * between batches.                                                                                                  */
int gt_extract_batch_async(gt_handle h, const uint8_t* frames, int B, int first_is_reference, float conf, float iou,
                           int agnostic, uint32_t classes_mask, const float* mask_boxes, const int32_t* mask_nboxes,
                           int mask_stride, float* out_boxes, int32_t* out_counts,
                           float* out_boxes_stab, double* out_H, int32_t* out_status, int32_t* out_stats, void* stream,
                           int32_t* ticket);
int gt_wait(gt_handle h, int ticket);

/* ---- instrumentation ---------------------------------------------------------------------------------------------
 * ms[0..3] = preprocess, inference, postprocess(decode+NMS), stabilize of the last call (fills Results.speed,
 * extract.py:155-156).  gt_launch_count: kernels launched by this handle since creation.                         */
int gt_stage_times(gt_handle h, float* ms4);
int64_t gt_launch_count(gt_handle h);
/* time (ms, CUDA events on the handle's stream) the conv stack took in the last gt_detect, and its FLOPs        */
int gt_conv_stack_stats(gt_handle h, float* ms, double* flops);
/* fused conv launches per forward, and how many of them the load-time autotune assigned to the swapped-operand kernel */
int gt_conv_kernel_info(gt_handle h, int32_t* n_ops, int32_t* n_swapped);
/* how many of the fused conv launches run as CTA pairs (cta_group::2 swapped kernel, cout a multiple of 256); < 0 on a null handle */
int gt_conv_pair_count(gt_handle h);

#ifdef __cplusplus
}
#endif
#endif /* GEOTRAX_B200_H_ */
