// C-ABI entry points (include/geotrax_b200.h) and engine lifecycle.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include "engine.cuh"

static std::string g_create_error;

void gt_set_error(gt_engine* e, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
}

int gt_engine::dev_alloc(void** p, size_t bytes) {
  cudaError_t r = cudaMalloc(p, bytes ? bytes : 16);
  if (r != cudaSuccess) {
    gt_set_error(this, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(r));
    return GT_ERR_NOMEM;
  }
  dev_allocs.push_back(*p);
  return GT_OK;
}
int gt_engine::host_alloc(void** p, size_t bytes) {
  cudaError_t r = cudaMallocHost(p, bytes ? bytes : 16);
  if (r != cudaSuccess) {
    gt_set_error(this, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(r));
    return GT_ERR_NOMEM;
  }
  host_allocs.push_back(*p);
  return GT_OK;
}

extern "C" {

int gt_abi_version(void) { return GT_ABI_VERSION; }

void gt_default_config(gt_config* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = GT_ABI_VERSION;
  c->frame_h = 2160; c->frame_w = 3840; c->max_batch = 16;
  c->imgsz = 1920; c->nc = 4; c->task = GT_TASK_DETECT; c->max_det = 1000; c->max_nms = 30000;
  c->downsample_ratio = 0.5f; c->max_features = 2000; c->ref_multiplier = 2.0f; c->mask_use = 1; c->mask_margin_ratio = 0.15f;
  c->filter_ratio = 0.9f; c->ransac_threshold = 2.0f; c->ransac_max_iter = 5000; c->query_is_current = 1; c->ransac_full_res = 0;
  c->seed = 0x9E3779B9u;
  c->act_dtype = GT_ACT_FP16;
  c->clahe = 0;
}

const char* gt_last_error(gt_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int round_half_even(double v) { return (int)std::nearbyint(v); }

int gt_create(const gt_config* cfg, int device, gt_handle* out) {
  if (!cfg || !out) { gt_set_error(nullptr, "gt_create: null argument"); return GT_ERR_INVALID; }
  *out = nullptr;
  if (cfg->abi_version != GT_ABI_VERSION) { gt_set_error(nullptr, "gt_create: ABI version %d != %d", cfg->abi_version, GT_ABI_VERSION); return GT_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    gt_set_error(nullptr, "gt_create: no CUDA device visible (this library has no CPU fallback)");
    return GT_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { gt_set_error(nullptr, "gt_create: device %d out of range (%d visible)", device, ndev); return GT_ERR_INVALID; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { gt_set_error(nullptr, "gt_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return GT_ERR_CUDA; }
  gt_engine* e = new gt_engine();
  e->cfg = *cfg;
  e->device = device;
  if (const char* hm = getenv("GT_HALO")) e->halo_mode = atoi(hm);
  if (const char* pm = getenv("GT_PAIR")) e->pair_mode = atoi(pm);
  if (const char* pd = getenv("GT_PDL")) e->pdl = atoi(pd);
  if (const char* sm = getenv("GT_SWAP")) e->swap_mode = atoi(sm);
  if (const char* tm = getenv("GT_TUNE")) e->tune_mode = atoi(tm);
  if (const char* ov = getenv("GT_OVERLAP")) e->overlap = atoi(ov);
  if (const char* fs = getenv("GT_FRONT_SPLIT")) e->front_split = atoi(fs);
  if (const char* sp = getenv("GT_SILU_TANH_PX")) e->silu_tanh_px = atoll(sp);
  if (const char* ms = getenv("GT_MASK_SPARSE")) e->mask_sparse = atoi(ms);
  if (const char* cm = getenv("GT_CHAIN")) e->chain_mode = atoi(cm);
  if (const char* nf = getenv("GT_NMS_FUSED")) e->nms_fused = atoi(nf);
  if (const char* lp = getenv("GT_L2PROMO")) e->l2promo_128 = atoi(lp) == 128;
  if (const char* mm = getenv("GT_MATCH")) e->match_mode = std::min(2, std::max(0, atoi(mm)));
  if (const char* kb = getenv("GT_CONV_SMEM_KB")) e->conv_smem_kb = std::min(227, std::max(96, atoi(kb)));
  if (e->overlap == 1 && !getenv("GT_CONV_SMEM_KB")) e->conv_smem_kb = 200;   // room for ORB blocks beside the conv CTAs
  auto fail = [&](int rc) { g_create_error = e->err; gt_destroy(e); return rc; };
#define CR(expr) do { int _rc = (expr); if (_rc != GT_OK) return fail(_rc); } while (0)
#define CRC(call) do { cudaError_t _er = (call); if (_er != cudaSuccess) { gt_set_error(e, "%s -> %s", #call, cudaGetErrorString(_er)); return fail(GT_ERR_CUDA); } } while (0)
  CRC(cudaSetDevice(device));
  {   // the handle's own stream is HIGH priority: its small kernels (decode, NMS, mask pyramid) must get SM slots ahead of the thousands of
      // queued FAST blocks of the low-priority aux stream that run beside them (GT_TIMELINE: the mask pyramid took 471 us instead of 177)
    int lo = 0, hi = 0;
    CRC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CRC(cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, hi));
  }
  for (int k = 0; k < 2; ++k) {
    for (int i = 0; i < 8; ++i) CRC(cudaEventCreate(&e->ev_sets[k][i]));
    CRC(cudaEventCreateWithFlags(&e->ev_done[k], cudaEventDisableTiming));
  }
  e->ev = e->ev_sets[0];
  const gt_config& c = e->cfg;
  if (c.act_dtype != GT_ACT_BF16 && c.act_dtype != GT_ACT_FP16) { gt_set_error(e, "gt_create: bad act_dtype %d", c.act_dtype); return fail(GT_ERR_INVALID); }
  if (c.max_batch < 1 || c.max_batch > 32 || c.nc < 1 || c.nc > 80 || c.max_det < 1 || c.max_det > 4096) {
    gt_set_error(e, "gt_create: max_batch/nc/max_det out of range"); return fail(GT_ERR_INVALID);
  }
  // letterbox geometry (LetterBox auto=True, stride 32)
  const double r = std::min((double)c.imgsz / c.frame_h, (double)c.imgsz / c.frame_w);
  e->gain = (float)r;
  e->new_w = round_half_even(c.frame_w * r);
  e->new_h = round_half_even(c.frame_h * r);
  const double dw = ((c.imgsz - e->new_w) % 32) / 2.0, dh = ((c.imgsz - e->new_h) % 32) / 2.0;
  e->pad_top = round_half_even(dh - 0.1); e->pad_left = round_half_even(dw - 0.1);
  const int pad_bot = round_half_even(dh + 0.1), pad_right = round_half_even(dw + 0.1);
  e->net_h = e->new_h + e->pad_top + pad_bot;
  e->net_w = e->new_w + e->pad_left + pad_right;
  if ((e->net_h % 32) || (e->net_w % 32) || e->new_w < 32 || e->new_h < 32) {
    gt_set_error(e, "gt_create: letterbox geometry %dx%d -> %dx%d (imgsz %d) is not a multiple of the stride", c.frame_w, c.frame_h, e->net_w, e->net_h, c.imgsz);
    return fail(GT_ERR_INVALID);
  }
  if (!(c.downsample_ratio > 0.f && c.downsample_ratio <= 1.0f)) { gt_set_error(e, "gt_create: downsample_ratio must be in (0, 1]"); return fail(GT_ERR_INVALID); }
  e->work_w = (int)(c.frame_w * c.downsample_ratio);
  e->work_h = (int)(c.frame_h * c.downsample_ratio);
  if (e->work_w < 128 || e->work_h < 128) { gt_set_error(e, "gt_create: working image %dx%d too small for the 8-level ORB pyramid", e->work_w, e->work_h); return fail(GT_ERR_INVALID); }
  // the fused vector kernel covers the default preset's geometry (exact 1/2 letterbox, 1/2 working image, 16-pixel-aligned rows)
  e->lb_fast = r == 0.5 && (c.frame_w % 16) == 0 && (c.frame_h % 4) == 0 && (e->pad_left % 8) == 0 && (e->pad_top % 2) == 0;
  e->pre_fast = e->lb_fast && c.downsample_ratio == 0.5f && e->work_w * 2 == c.frame_w && e->work_h * 2 == c.frame_h;
  const int B = c.max_batch;
  CR(e->dev_alloc((void**)&e->frames_dev, (size_t)B * c.frame_h * c.frame_w * 3));
  CR(e->dev_alloc((void**)&e->frames_dev2, (size_t)B * c.frame_h * c.frame_w * 3));
  CRC(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    CRC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CRC(cudaStreamCreateWithPriority(&e->aux_stream, cudaStreamNonBlocking, lo));
    CRC(cudaStreamCreateWithPriority(&e->aux2_stream, cudaStreamNonBlocking, lo));
    CRC(cudaEventCreateWithFlags(&e->ev_aux_a, cudaEventDisableTiming));
    CRC(cudaEventCreateWithFlags(&e->ev_aux_b, cudaEventDisableTiming));
    CRC(cudaEventCreateWithFlags(&e->ev_pre, cudaEventDisableTiming));
    CRC(cudaEventCreateWithFlags(&e->ev_front, cudaEventDisableTiming));
  }
  for (int i = 0; i < 2; ++i) {
    CRC(cudaEventCreateWithFlags(&e->ev_copied[i], cudaEventDisableTiming));
    CRC(cudaEventCreateWithFlags(&e->ev_consumed[i], cudaEventDisableTiming));
  }
  CR(conv_tc_init(e));
  CR(orb_build(e));      // allocates the pyramid slabs (level 0 = stage-1 gray output)
  CR(detector_build(e));
  CR(stab_build(e));
  // constant letterbox border
  if (!e->pre_fast || e->cfg.clahe) CR(detector_build_general_preprocess(e));
  CR(clahe_build(e));
  CR(detector_fill_pad(e, e->stream));
  CRC(cudaStreamSynchronize(e->stream));
#undef CR
#undef CRC
  *out = e;
  return GT_OK;
}

int gt_destroy(gt_handle e) {
  if (!e) return GT_OK;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (void* p : e->dev_allocs) cudaFree(p);
  for (void* p : {(void*)e->reg_xq, (void*)e->reg_xt, (void*)e->reg_cq, (void*)e->reg_ct, (void*)e->reg_cand, (void*)e->reg_n, (void*)e->reg_q32,
                  (void*)e->reg_t32, (void*)e->reg_idx, (void*)e->reg_dist})
    if (p) cudaFree(p);
  for (void* p : e->host_allocs) cudaFreeHost(p);
  for (int k = 0; k < 2; ++k) {
    for (int i = 0; i < 8; ++i) if (e->ev_sets[k][i]) cudaEventDestroy(e->ev_sets[k][i]);
    if (e->ev_done[k]) cudaEventDestroy(e->ev_done[k]);
  }
  for (int i = 0; i < 2; ++i) {
    if (e->ev_copied[i]) cudaEventDestroy(e->ev_copied[i]);
    if (e->ev_consumed[i]) cudaEventDestroy(e->ev_consumed[i]);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
  if (e->aux2_stream) cudaStreamDestroy(e->aux2_stream);
  if (e->ev_aux_a) cudaEventDestroy(e->ev_aux_a);
  if (e->ev_aux_b) cudaEventDestroy(e->ev_aux_b);
  if (e->ev_pre) cudaEventDestroy(e->ev_pre);
  if (e->ev_front) cudaEventDestroy(e->ev_front);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return GT_OK;
}

}  // extern "C"

// ---- helpers -----------------------------------------------------------------------------------------------------------
static bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
static bool is_pinned_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}
static cudaStream_t pick_stream(gt_engine* e, void* s) { return s ? (cudaStream_t)s : e->stream; }

// copy a caller buffer (host or device) into a device workspace
static int to_device(gt_engine* e, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!bytes) return GT_OK;
  GT_CUDA(e, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
  if (!is_device_ptr(src) && !is_pinned_ptr(src)) GT_CUDA(e, cudaStreamSynchronize(st));  // pageable source: keep caller semantics simple
  return GT_OK;
}
static int to_caller(gt_engine* e, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!bytes || !dst) return GT_OK;
  GT_CUDA(e, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
  return GT_OK;
}

#define ENTER(e)                                  \
  if (!(e)) return GT_ERR_INVALID;                \
  if (cudaSetDevice((e)->device) != cudaSuccess) { gt_set_error((e), "cudaSetDevice failed"); return GT_ERR_CUDA; }

extern "C" {

int gt_conv_count(gt_handle e) { return e ? (int)e->conv_descs.size() : GT_ERR_INVALID; }
int gt_conv_info(gt_handle e, int idx, gt_conv_desc* out) {
  if (!e || !out || idx < 0 || idx >= (int)e->conv_descs.size()) return GT_ERR_INVALID;
  *out = e->conv_descs[idx];
  return GT_OK;
}
int gt_load_weights(gt_handle e, const float* const* weights, const float* const* biases, int n_convs) {
  ENTER(e);
  GT_TRY(detector_load_weights(e, weights, biases, n_convs));
  GT_TRY(detector_autotune(e, e->stream));
  GT_CUDA(e, cudaStreamSynchronize(e->stream));
  return GT_OK;
}

static size_t frame_bytes(const gt_engine* e) {
  const size_t px = (size_t)e->cfg.frame_h * e->cfg.frame_w;
  return e->input_format == GT_INPUT_NV12 ? px * 3 / 2 : px * 3;
}

int gt_set_input_format(gt_handle e, int format) {
  ENTER(e);
  GT_CHECK(e, format == GT_INPUT_BGR24 || format == GT_INPUT_NV12, "gt_set_input_format: unknown format %d", format);
  GT_CHECK(e, format == GT_INPUT_BGR24 || ((e->cfg.frame_h % 2) == 0 && (e->cfg.frame_w % 2) == 0), "gt_set_input_format: NV12 needs even frame dimensions");
  if (format == GT_INPUT_NV12 && !e->frames_bgr)
    GT_TRY(e->dev_alloc((void**)&e->frames_bgr, (size_t)e->cfg.max_batch * e->cfg.frame_h * e->cfg.frame_w * 3));
  e->input_format = format;
  return GT_OK;
}

// Starts the host->device copy of a batch on the copy stream into the idle staging buffer; a later gt_preprocess /
// gt_extract_batch with the same `frames` pointer consumes it instead of copying.  Overlaps PCIe ingest with compute.
int gt_prefetch_frames(gt_handle e, const uint8_t* frames, int B) {
  ENTER(e);
  GT_CHECK(e, frames && B >= 1 && B <= e->cfg.max_batch, "gt_prefetch_frames: bad batch %d", B);
  if (is_device_ptr(frames)) return GT_OK;   // nothing to stage
  int k = e->prefetch_next;
  if (e->prefetched_src[k] != nullptr && e->prefetched_src[k ^ 1] == nullptr) k ^= 1;   // never overwrite an unconsumed prefetch
  e->prefetch_next = k ^ 1;
  uint8_t* dst = k ? e->frames_dev2 : e->frames_dev;
  GT_CUDA(e, cudaStreamWaitEvent(e->copy_stream, e->ev_consumed[k], 0));   // the kernel that last read this buffer is done
  // one copy per frame: small uploads of the running batch (mask boxes) are not stuck behind one 400 MB transfer
  const size_t fb = frame_bytes(e);
  for (int b = 0; b < B; ++b) GT_CUDA(e, cudaMemcpyAsync(dst + b * fb, frames + b * fb, fb, cudaMemcpyHostToDevice, e->copy_stream));
  GT_CUDA(e, cudaEventRecord(e->ev_copied[k], e->copy_stream));
  e->prefetched_src[k] = frames;
  return GT_OK;
}

// Same, but the copy is started by the next gt_extract_batch right after it has queued its own small inputs (mask boxes),
// so those are not stuck behind 400 MB of frames on the copy engine.
int gt_prefetch_frames_deferred(gt_handle e, const uint8_t* frames, int B) {
  ENTER(e);
  GT_CHECK(e, frames && B >= 1 && B <= e->cfg.max_batch, "gt_prefetch_frames_deferred: bad batch %d", B);
  e->deferred_src = frames;
  e->deferred_B = B;
  return GT_OK;
}

int gt_preprocess(gt_handle e, const uint8_t* frames, int B, void* stream) {
  ENTER(e);
  GT_CHECK(e, frames && B >= 1 && B <= e->cfg.max_batch, "gt_preprocess: bad batch %d", B);
  cudaStream_t st = pick_stream(e, stream);
  const uint8_t* src = frames;
  int used = -1;
  GT_CUDA(e, cudaEventRecord(e->ev[0], st));
  if (!is_device_ptr(frames)) {
    for (int k = 0; k < 2; ++k)
      if (e->prefetched_src[k] == frames) used = k;
    if (used >= 0) {   // already on its way (gt_prefetch_frames): just order this stream after the copy
      GT_CUDA(e, cudaStreamWaitEvent(st, e->ev_copied[used], 0));
      e->prefetched_src[used] = nullptr;
    } else {
      used = e->prefetch_next;
      if (e->prefetched_src[used] != nullptr && e->prefetched_src[used ^ 1] == nullptr) used ^= 1;   // keep a pending prefetch intact
      if (e->prefetched_src[used] != nullptr) {   // both staging buffers hold pending prefetches: the older one is dropped
        GT_CUDA(e, cudaStreamWaitEvent(st, e->ev_copied[used], 0));
        e->prefetched_src[used] = nullptr;
      }
      e->prefetch_next = used ^ 1;
      GT_CUDA(e, cudaStreamWaitEvent(st, e->ev_consumed[used], 0));
      const size_t bytes = (size_t)B * frame_bytes(e);
      GT_TRY(to_device(e, used ? e->frames_dev2 : e->frames_dev, frames, bytes, st));
    }
    src = used ? e->frames_dev2 : e->frames_dev;
  }
  if (e->input_format == GT_INPUT_NV12 && e->pre_fast && (e->cfg.frame_w % 32) == 0 && !e->cfg.clahe) {   // default geometry: fused NV12 -> letterbox + gray
    GT_TRY(detector_preprocess_nv12(e, src, B, st));
    if (used >= 0) GT_CUDA(e, cudaEventRecord(e->ev_consumed[used], st));
    GT_CUDA(e, cudaEventRecord(e->ev[1], st));
    return GT_OK;
  }
  if (e->input_format == GT_INPUT_NV12) {   // decoder-format ingest, any geometry: NV12 -> BGR24 on the device, then the BGR path
    GT_TRY(detector_nv12_to_bgr(e, src, e->frames_bgr, B, st));
    if (used >= 0) { GT_CUDA(e, cudaEventRecord(e->ev_consumed[used], st)); used = -1; }   // the staging buffer is free already
    src = e->frames_bgr;
  }
  GT_TRY(detector_preprocess(e, src, B, st));
  if (e->cfg.clahe) GT_TRY(clahe_run(e, src, B, st));   // the working image is CLAHE(gray) instead of gray (stable preset)
  if (used >= 0) GT_CUDA(e, cudaEventRecord(e->ev_consumed[used], st));
  GT_CUDA(e, cudaEventRecord(e->ev[1], st));
  return GT_OK;
}

int gt_get_net_input(gt_handle e, int B, uint8_t* out, int32_t* net_h, int32_t* net_w) {
  ENTER(e);
  if (net_h) *net_h = e->net_h;
  if (net_w) *net_w = e->net_w;
  if (out) {  // debug read-back: space-to-depth 16-bit tensor -> planar RGB u8
    GT_CHECK(e, B >= 1 && B <= e->cfg.max_batch, "gt_get_net_input: bad batch %d", B);
    GT_CUDA(e, cudaStreamSynchronize(e->stream));
    const int sh = e->net_h / 4, sw = e->net_w / 4;
    std::vector<uint16_t> h((size_t)B * sh * sw * 64);
    GT_CUDA(e, cudaMemcpy(h.data(), e->net_s2d, h.size() * 2, cudaMemcpyDefault));
    const bool fp16 = e->cfg.act_dtype == GT_ACT_FP16;
    const size_t plane = (size_t)e->net_h * e->net_w;
    for (int b = 0; b < B; ++b)
      for (int sy = 0; sy < sh; ++sy)
        for (int sx = 0; sx < sw; ++sx) {
          const uint16_t* px = &h[(((size_t)b * sh + sy) * sw + sx) * 64];
          for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c)
              for (int ch = 0; ch < 3; ++ch) {
                const uint16_t bits = px[(r * 4 + c) * 4 + ch];
                float f;
                if (fp16) { __half hv; memcpy(&hv, &bits, 2); f = __half2float(hv); }
                else { uint32_t u = (uint32_t)bits << 16; memcpy(&f, &u, 4); }
                out[((size_t)b * 3 + ch) * plane + (size_t)(4 * sy + r) * e->net_w + 4 * sx + c] = (uint8_t)(f + 0.5f);
              }
        }
  }
  return GT_OK;
}
int gt_get_gray(gt_handle e, int B, uint8_t* out, int32_t* work_h, int32_t* work_w) {
  ENTER(e);
  if (work_h) *work_h = e->work_h;
  if (work_w) *work_w = e->work_w;
  if (out) {
    GT_CUDA(e, cudaStreamSynchronize(e->stream));
    GT_CUDA(e, cudaMemcpy2D(out, (size_t)e->work_h * e->work_w, e->pyr, e->pyr_bytes, (size_t)e->work_h * e->work_w, B, cudaMemcpyDefault));
  }
  return GT_OK;
}

static int detect_impl(gt_engine* e, int B, float conf, float iou, int agnostic, uint32_t classes_mask, cudaStream_t st) {
  GT_CUDA(e, cudaEventRecord(e->ev[2], st));
  GT_TRY(detector_forward(e, B, st));
  GT_CUDA(e, cudaEventRecord(e->ev[3], st));
  GT_TRY(detector_postprocess(e, B, conf, iou, agnostic, classes_mask, st));
  if (!e->post_event_done) GT_CUDA(e, cudaEventRecord(e->ev[4], st));   // (else nms_run recorded it right behind its last kernel)
  e->post_event_done = false;
  return GT_OK;
}

static int copy_dets(gt_engine* e, int B, float* out_boxes, int32_t* out_counts, int32_t* out_keep, cudaStream_t st) {
  const int row = e->cfg.task == GT_TASK_OBB ? 7 : 6;
  GT_TRY(to_caller(e, out_boxes, e->det_out, (size_t)B * e->cfg.max_det * row * sizeof(float), st));
  GT_TRY(to_caller(e, out_counts, e->det_count, (size_t)B * sizeof(int), st));
  GT_TRY(to_caller(e, out_keep, e->det_keep, (size_t)B * e->cfg.max_det * sizeof(int), st));
  return GT_OK;
}

static void update_times(gt_engine* e) {
  float t;
  if (cudaEventElapsedTime(&t, e->ev[0], e->ev[1]) == cudaSuccess) e->stage_ms[0] = t;
  if (cudaEventElapsedTime(&t, e->ev[2], e->ev[3]) == cudaSuccess) { e->stage_ms[1] = t; e->conv_ms = t; }
  if (cudaEventElapsedTime(&t, e->ev[3], e->ev[4]) == cudaSuccess) e->stage_ms[2] = t;
  if (cudaEventElapsedTime(&t, e->ev[5], e->ev[6]) == cudaSuccess) e->stage_ms[3] = t;
  cudaGetLastError();
}

int gt_detect(gt_handle e, int B, float conf, float iou, int agnostic, uint32_t classes_mask, float* out_boxes, int32_t* out_counts,
              int32_t* out_keep, void* stream) {
  ENTER(e);
  GT_CHECK(e, B >= 1 && B <= e->cfg.max_batch, "gt_detect: bad batch %d", B);
  cudaStream_t st = pick_stream(e, stream);
  GT_TRY(detect_impl(e, B, conf, iou, agnostic, classes_mask, st));
  GT_TRY(copy_dets(e, B, out_boxes, out_counts, out_keep, st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  update_times(e);
  return GT_OK;
}

int gt_set_class_filter(gt_handle e, const int32_t* classes, int n) {
  ENTER(e);
  GT_CHECK(e, n >= 0 && (classes || n == 0), "gt_set_class_filter: bad arguments");
  e->cls_filter[0] = e->cls_filter[1] = e->cls_filter[2] = 0u;
  e->cls_filter_on = classes != nullptr;
  for (int i = 0; i < n; ++i) {
    GT_CHECK(e, classes[i] >= 0 && classes[i] < 96, "gt_set_class_filter: class id %d out of range", (int)classes[i]);
    e->cls_filter[classes[i] >> 5] |= 1u << (classes[i] & 31);
  }
  return GT_OK;
}

int gt_get_health(gt_handle e, int64_t* nonfinite_rows) {
  if (!e || !nonfinite_rows) return GT_ERR_INVALID;
  *nonfinite_rows = e->nonfinite_host ? (int64_t)*(volatile int*)e->nonfinite_host : 0;
  return GT_OK;
}

int gt_get_candidate_counts(gt_handle e, int B, int32_t* out_counts) {
  ENTER(e);
  GT_CHECK(e, out_counts && B >= 1 && B <= e->cfg.max_batch, "gt_get_candidate_counts: bad arguments");
  GT_CUDA(e, cudaStreamSynchronize(e->stream));
  GT_CUDA(e, cudaMemcpy(out_counts, e->cand_count, sizeof(int) * B, cudaMemcpyDefault));
  return GT_OK;
}

int gt_get_raw_head(gt_handle e, int B, float* out, int32_t* A, int32_t* no) {
  ENTER(e);
  if (A) *A = e->A;
  if (no) *no = e->no;
  if (out) {
    GT_CUDA(e, cudaStreamSynchronize(e->stream));
    const size_t rows = (size_t)B * e->A;   // dense [no] rows assembled from the three device arrays
    GT_CUDA(e, cudaMemcpy2D(out, (size_t)e->no * 4, e->raw_box, 64 * 4, 64 * 4, rows, cudaMemcpyDefault));
    GT_CUDA(e, cudaMemcpy2D(out + 64, (size_t)e->no * 4, e->raw_cls, (size_t)e->ncp * 4, (size_t)e->cfg.nc * 4, rows, cudaMemcpyDefault));
    if (e->cfg.task == GT_TASK_OBB)
      GT_CUDA(e, cudaMemcpy2D(out + 64 + e->cfg.nc, (size_t)e->no * 4, e->raw_ang, 4 * 4, 4, rows, cudaMemcpyDefault));
  }
  return GT_OK;
}

int gt_get_feature(gt_handle e, int layer, int B, uint16_t* out, int32_t* C, int32_t* H, int32_t* W) {
  ENTER(e);
  GT_CHECK(e, layer >= 0 && layer < 23 && e->feat_views[layer].ptr, "gt_get_feature: layer %d has no buffer", layer);
  const View& v = e->feat_views[layer];
  if (C) *C = v.C;
  if (H) *H = v.H;
  if (W) *W = v.W;
  if (out) {
    GT_CUDA(e, cudaStreamSynchronize(e->stream));
    GT_CUDA(e, cudaMemcpy2D(out, (size_t)v.C * 2, v.ptr + v.coff, (size_t)v.ctot * 2, (size_t)v.C * 2, (size_t)B * v.H * v.W, cudaMemcpyDefault));
  }
  return GT_OK;
}

int gt_nms(gt_handle e, const float* pred, int B, int A, int nc, int rotated, float conf, float iou, int agnostic, uint32_t classes_mask,
           int max_det, float* out_rows, int32_t* out_counts, int32_t* out_keep, void* stream) {
  ENTER(e);
  GT_CHECK(e, pred && B >= 1 && B <= e->cfg.max_batch && A >= 1 && A <= e->A && nc >= 1 && nc <= 80, "gt_nms: bad shape B=%d A=%d nc=%d", B, A, nc);
  cudaStream_t st = pick_stream(e, stream);
  const int row_in = 4 + nc + (rotated ? 1 : 0);
  const size_t bytes = (size_t)B * A * row_in * sizeof(float);
  const float* src = pred;
  if (!is_device_ptr(pred)) {
    if (bytes > e->pred_tmp_bytes) {
      GT_TRY(e->dev_alloc((void**)&e->pred_tmp, bytes));
      e->pred_tmp_bytes = bytes;
    }
    GT_TRY(to_device(e, e->pred_tmp, pred, bytes, st));
    src = e->pred_tmp;
  }
  GT_TRY(nms_run(e, src, B, A, nc, rotated, conf, iou, agnostic, classes_mask, max_det, false, st));
  const int row = rotated ? 7 : 6;
  GT_TRY(to_caller(e, out_rows, e->det_out, (size_t)B * max_det * row * sizeof(float), st));
  GT_TRY(to_caller(e, out_counts, e->det_count, (size_t)B * sizeof(int), st));
  GT_TRY(to_caller(e, out_keep, e->det_keep, (size_t)B * max_det * sizeof(int), st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  return GT_OK;
}

int gt_conv2d(gt_handle e, const uint16_t* x, int B, int H, int W, int cin, const float* w, const float* bias, int cout, int k, int stride,
              int act, const uint16_t* residual, void* out, int out_f32, void* stream) {
  ENTER(e);
  GT_CHECK(e, x && w && out && B >= 1 && cin >= 8 && (cin % 8) == 0 && cout >= 1, "gt_conv2d: bad arguments");
  GT_CHECK(e, out_f32 || (cout % 8) == 0, "gt_conv2d: bf16 output needs cout %% 8 == 0");
  cudaStream_t st = pick_stream(e, stream);
  const int pad = k / 2, Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  bf16 *dx = nullptr, *dres = nullptr;
  void* dout = nullptr;
  const size_t mark = e->dev_allocs.size();
  auto cleanup = [&]() {
    cudaStreamSynchronize(st);
    while (e->dev_allocs.size() > mark) { cudaFree(e->dev_allocs.back()); e->dev_allocs.pop_back(); }
  };
  int rc = GT_OK;
  do {
    if ((rc = e->dev_alloc((void**)&dx, (size_t)B * H * W * cin * 2)) != GT_OK) break;
    if ((rc = e->dev_alloc(&dout, (size_t)B * Ho * Wo * cout * (out_f32 ? 4 : 2))) != GT_OK) break;
    if ((rc = to_device(e, dx, x, (size_t)B * H * W * cin * 2, st)) != GT_OK) break;
    View in; in.ptr = dx; in.C = cin; in.ctot = cin; in.coff = 0; in.H = H; in.W = W;
    View ov; ov.ptr = (bf16*)dout; ov.C = cout; ov.ctot = cout; ov.coff = 0; ov.H = Ho; ov.W = Wo;
    View rv = ov;
    if (residual) {
      if ((rc = e->dev_alloc((void**)&dres, (size_t)B * Ho * Wo * cout * 2)) != GT_OK) break;
      if ((rc = to_device(e, dres, residual, (size_t)B * Ho * Wo * cout * 2, st)) != GT_OK) break;
      rv.ptr = dres;
    }
    ConvOp op;
    ConvPlanArgs pa;
    pa.in = in; pa.Bmax = B; pa.cin = cin; pa.cout = cout; pa.k = k; pa.stride = stride; pa.act = act;
    if (out_f32) { pa.out_f32 = (float*)dout; pa.out_img_stride = (long long)Ho * Wo; pa.out_ctot_f32 = cout; pa.out_coff_f32 = 0; }
    else pa.out = &ov;
    pa.res = residual ? &rv : nullptr;
    e->plan_variant = e->swap_mode < 0 ? 2 : e->swap_mode;   // unit parity of the swapped kernel (+ halo staging where it applies) by default; GT_SWAP=0: the pixel-major one, 1: swapped without halo
    rc = conv_tc_plan(e, &op, pa);
    e->plan_variant = 0;
    if (rc != GT_OK) break;
    const float* ws[1] = {w};
    const float* bs[1] = {bias};
    int couts[1] = {cout};
    if ((rc = conv_tc_pack_weights(e, &op, ws, bs, couts, 1)) != GT_OK) break;
    if ((rc = conv_tc_launch(e, &op, B, st)) != GT_OK) break;
    cudaError_t er = cudaStreamSynchronize(st);
    if (er != cudaSuccess) { gt_set_error(e, "gt_conv2d kernel failed: %s", cudaGetErrorString(er)); rc = GT_ERR_CUDA; break; }
    er = cudaMemcpy(out, dout, (size_t)B * Ho * Wo * cout * (out_f32 ? 4 : 2), cudaMemcpyDefault);
    if (er != cudaSuccess) { gt_set_error(e, "gt_conv2d copy-out failed: %s", cudaGetErrorString(er)); rc = GT_ERR_CUDA; break; }
  } while (0);
  cleanup();
  return rc;
}

// ---- stage 3 ------------------------------------------------------------------------------------------------------------
// src / n live in mapped pinned host memory (cudaMallocHost, UVA: the host pointer is valid on the device)
__global__ void boxes_from_host_kernel(const float* __restrict__ src, const int* __restrict__ n, float* __restrict__ dst, int* __restrict__ ndst, int md) {
  const int b = blockIdx.x, cnt = n[b];
  const float4* s4 = reinterpret_cast<const float4*>(src) + (size_t)b * md;
  float4* d4 = reinterpret_cast<float4*>(dst) + (size_t)b * md;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) d4[i] = s4[i];
  if (threadIdx.x == 0) ndst[b] = cnt;
}

static int upload_boxes(gt_engine* e, int slot0, int B, const float* boxes, const int32_t* nboxes, int box_stride, cudaStream_t st) {
  const int md = e->cfg.max_det;
  if (!boxes || !nboxes) {
    GT_CUDA(e, cudaMemsetAsync(e->nboxes_dev + slot0, 0, sizeof(int) * B, st));
    return GT_OK;
  }
  GT_CHECK(e, box_stride >= 0 && box_stride <= md, "boxes: stride %d exceeds max_det %d", box_stride, md);
  // Host boxes never go through the H2D copy engine: it is busy with the next batch's 400 MB of frames (gt_prefetch_frames), and
  // a 256 KB upload queued behind that copy delayed the whole step by milliseconds.  They are written into a mapped pinned
  // staging buffer by the CPU (only nboxes[b] rows per frame) and pulled over PCIe by a small kernel (zero-copy reads).
  cudaPointerAttributes at;
  const bool on_device = cudaPointerGetAttributes(&at, boxes) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
  cudaPointerAttributes an;
  const bool n_on_device = cudaPointerGetAttributes(&an, nboxes) == cudaSuccess && (an.type == cudaMemoryTypeDevice || an.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  if (!on_device && !n_on_device && B <= e->cfg.max_batch + 1) {
    if (!e->box_stage[0]) {
      for (int k = 0; k < 2; ++k) {
        GT_TRY(e->host_alloc((void**)&e->box_stage[k], (size_t)(e->cfg.max_batch + 1) * md * 16));
        GT_TRY(e->host_alloc((void**)&e->nbox_stage[k], (size_t)(e->cfg.max_batch + 1) * sizeof(int)));
      }
    }
    const int k = e->box_stage_next;
    e->box_stage_next ^= 1;
    for (int b = 0; b < B; ++b) {
      const int n = std::min(std::max((int)nboxes[b], 0), box_stride);
      memcpy(e->box_stage[k] + (size_t)b * md * 4, boxes + (size_t)b * box_stride * 4, (size_t)n * 16);
      e->nbox_stage[k][b] = n;
    }
    boxes_from_host_kernel<<<B, 128, 0, st>>>(e->box_stage[k], e->nbox_stage[k], e->boxes_dev + (size_t)slot0 * md * 4, e->nboxes_dev + slot0, md);
    e->launches++;
    GT_CUDA(e, cudaGetLastError());
    return GT_OK;
  }
  GT_CUDA(e, cudaMemcpy2DAsync(e->boxes_dev + (size_t)slot0 * md * 4, (size_t)md * 16, boxes, (size_t)box_stride * 16, (size_t)box_stride * 16, B,
                               cudaMemcpyDefault, st));
  GT_CUDA(e, cudaMemcpyAsync(e->nboxes_dev + slot0, nboxes, sizeof(int) * B, cudaMemcpyDefault, st));
  return GT_OK;
}

int gt_set_reference(gt_handle e, int frame_slot, const float* boxes, int nboxes, void* stream) {
  ENTER(e);
  GT_CHECK(e, frame_slot >= 0 && frame_slot < e->cfg.max_batch, "gt_set_reference: slot %d out of range", frame_slot);
  cudaStream_t st = pick_stream(e, stream);
  const int R = e->cfg.max_batch;  // reference slot
  // copy gray level 0 of the chosen frame into the reference slab
  GT_CUDA(e, cudaMemcpyAsync(e->pyr + (size_t)R * e->pyr_bytes, e->pyr + (size_t)frame_slot * e->pyr_bytes, (size_t)e->work_h * e->work_w,
                             cudaMemcpyDeviceToDevice, st));
  int32_t nb = nboxes;
  GT_TRY(upload_boxes(e, R, 1, nboxes > 0 ? boxes : nullptr, nboxes > 0 ? &nb : nullptr, nboxes, st));
  GT_CUDA(e, cudaStreamSynchronize(st));  // nb is a stack variable
  GT_TRY(orb_run(e, R, 1, true, true, st));
  e->have_ref = true;
  return GT_OK;
}

// GT_TIMELINE=1 (diagnostic): timestamps inside one synchronous gt_extract_batch, printed relative to the end of the preprocess kernel
static cudaEvent_t g_tl[12];
static bool g_tl_on = false, g_tl_init = false;
static void tl_mark(int i, cudaStream_t st) {
  if (!g_tl_init) { g_tl_init = true; g_tl_on = getenv("GT_TIMELINE") != nullptr; if (g_tl_on) for (auto& ev : g_tl) cudaEventCreate(&ev); }
  if (g_tl_on) cudaEventRecord(g_tl[i], st);
}
static void tl_print() {
  if (!g_tl_on) return;
  static const char* name[12] = {"preprocess done", "conv stack done", "decode+NMS done", "mask pyramid done", "front (aux) done", "waited for front", "select+describe done",
                                 "match+RANSAC done", "", "", "", ""};
  fprintf(stderr, "[gt timeline]");
  for (int i = 1; i < 8; ++i) { float ms = 0.f; if (cudaEventElapsedTime(&ms, g_tl[0], g_tl[i]) == cudaSuccess) fprintf(stderr, " %s %.0f us |", name[i], ms * 1e3f); }
  fprintf(stderr, "\n");
  cudaGetLastError();
}

static int stabilize_impl(gt_engine* e, int B, cudaStream_t st, bool front_on_aux = false) {
  GT_CHECK(e, e->have_ref, "gt_stabilize: no reference frame set");
  GT_CUDA(e, cudaEventRecord(e->ev[5], st));
  // the vehicle-mask pyramid needs only the boxes: it is enqueued BEFORE this stream waits for the aux stream's image pyramid + FAST,
  // so it overlaps the tail of the FAST kernel (both are latency / instruction bound and share the SMs well)
  if (front_on_aux) {
    GT_TRY(orb_mask(e, 0, B, true, st));
    tl_mark(3, st);
    GT_CUDA(e, cudaStreamWaitEvent(st, e->ev_front, 0));
    tl_mark(5, st);
  } else {
    GT_TRY(orb_front(e, 0, B, st));
    GT_TRY(orb_mask(e, 0, B, true, st));
  }
  GT_TRY(orb_back(e, 0, B, false, st));
  tl_mark(6, st);
  GT_TRY(stab_match_and_fit(e, B, st));
  tl_mark(7, st);
  GT_CUDA(e, cudaEventRecord(e->ev[6], st));
  return GT_OK;
}

int gt_stabilize(gt_handle e, int B, const float* boxes, const int32_t* nboxes, int box_stride, double* out_H, int32_t* out_status,
                 int32_t* out_stats, void* stream) {
  ENTER(e);
  GT_CHECK(e, B >= 1 && B <= e->cfg.max_batch, "gt_stabilize: bad batch %d", B);
  cudaStream_t st = pick_stream(e, stream);
  GT_TRY(upload_boxes(e, 0, B, boxes, nboxes, box_stride, st));
  GT_TRY(stabilize_impl(e, B, st));
  GT_TRY(to_caller(e, out_H, e->H_dev, (size_t)B * 9 * sizeof(double), st));
  GT_TRY(to_caller(e, out_status, e->H_status, (size_t)B * sizeof(int), st));
  GT_TRY(to_caller(e, out_stats, e->H_stats, (size_t)B * 4 * sizeof(int), st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  update_times(e);
  return GT_OK;
}

int gt_warp_boxes(gt_handle e, const double* H, float* boxes, int n, void* stream) {
  ENTER(e);
  GT_CHECK(e, H && boxes && n >= 0 && n <= e->cfg.max_det, "gt_warp_boxes: bad arguments (n=%d)", n);
  if (n == 0) return GT_OK;
  cudaStream_t st = pick_stream(e, stream);
  GT_TRY(to_device(e, e->H_dev, H, 9 * sizeof(double), st));
  GT_TRY(to_device(e, e->boxes_dev, boxes, (size_t)n * 16, st));
  GT_CUDA(e, cudaMemcpyAsync(e->nboxes_dev, &n, sizeof(int), cudaMemcpyHostToDevice, st));
  GT_TRY(warp_boxes_run(e, e->H_dev, nullptr, e->boxes_dev, e->boxes_stab_dev, e->nboxes_dev, 1, e->cfg.max_det, st));
  GT_TRY(to_caller(e, boxes, e->boxes_stab_dev, (size_t)n * 16, st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  return GT_OK;
}

int gt_warp_frames(gt_handle e, const uint8_t* frames, const double* H, int B, uint8_t* out, void* stream) {
  ENTER(e);
  GT_CHECK(e, frames && H && out && B >= 1 && B <= e->cfg.max_batch, "gt_warp_frames: bad arguments (B=%d)", B);
  cudaStream_t st = pick_stream(e, stream);
  const size_t bytes = (size_t)B * e->cfg.frame_h * e->cfg.frame_w * 3;
  const uint8_t* src = frames;
  if (!is_device_ptr(frames)) {
    GT_TRY(to_device(e, e->frames_dev, frames, bytes, st));
    src = e->frames_dev;
  }
  uint8_t* dst = out;
  if (!is_device_ptr(out)) {
    if (!e->warp_out) GT_TRY(e->dev_alloc((void**)&e->warp_out, (size_t)e->cfg.max_batch * e->cfg.frame_h * e->cfg.frame_w * 3));
    dst = e->warp_out;
  }
  GT_TRY(warp_frames_run(e, src, dst, H, B, st));
  if (dst != out) GT_CUDA(e, cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToHost, st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  return GT_OK;
}

int gt_orb_level_info(gt_handle e, int level, int32_t* w, int32_t* hgt, int32_t* quota_cur, int32_t* quota_ref) {
  if (!e || level < 0 || level >= GT_ORB_LEVELS) return GT_ERR_INVALID;
  if (w) *w = e->lv[level].w;
  if (hgt) *hgt = e->lv[level].h;
  if (quota_cur) *quota_cur = e->lv[level].quota_cur;
  if (quota_ref) *quota_ref = e->lv[level].quota_ref;
  return GT_OK;
}

int gt_get_pyramid_level(gt_handle e, int which, int b, int level, uint8_t* out_img, uint8_t* out_mask) {
  ENTER(e);
  GT_CHECK(e, level >= 0 && level < GT_ORB_LEVELS && b >= 0 && b < e->cfg.max_batch, "gt_get_pyramid_level: bad index");
  const int slot = which ? e->cfg.max_batch : b;
  const OrbLevel& L = e->lv[level];
  GT_CUDA(e, cudaStreamSynchronize(e->stream));
  if (out_img) GT_CUDA(e, cudaMemcpy(out_img, e->pyr + (size_t)slot * e->pyr_bytes + L.off, (size_t)L.w * L.h, cudaMemcpyDefault));
  if (out_mask) GT_CUDA(e, cudaMemcpy(out_mask, e->pyr_mask + (size_t)slot * e->pyr_bytes + L.off, (size_t)L.w * L.h, cudaMemcpyDefault));
  return GT_OK;
}

int gt_get_keypoints(gt_handle e, int which, int b, int max_n, float* out_kp, uint8_t* out_desc, int32_t* n) {
  ENTER(e);
  GT_CHECK(e, b >= 0 && b < e->cfg.max_batch && n, "gt_get_keypoints: bad index");
  const int slot = which ? e->cfg.max_batch : b;
  GT_CUDA(e, cudaStreamSynchronize(e->stream));
  int cnt = 0;
  GT_CUDA(e, cudaMemcpy(&cnt, e->kp_count + slot, sizeof(int), cudaMemcpyDefault));
  cnt = std::min(cnt, GT_MAX_KP);
  *n = cnt;
  const int m = std::min(cnt, max_n);
  if (out_kp && m) GT_CUDA(e, cudaMemcpy(out_kp, e->kp_all + (size_t)slot * GT_MAX_KP * 6, (size_t)m * 6 * sizeof(float), cudaMemcpyDefault));
  if (out_desc && m) GT_CUDA(e, cudaMemcpy(out_desc, e->desc_all + (size_t)slot * GT_MAX_KP * 32, (size_t)m * 32, cudaMemcpyDefault));
  return GT_OK;
}

int gt_orb_get_candidates(gt_handle e, int which, int b, int level, int max_n, uint32_t* out_xy, uint8_t* out_score, int32_t* n) {
  ENTER(e);
  GT_CHECK(e, level >= 0 && level < GT_ORB_LEVELS && b >= 0 && b < e->cfg.max_batch && n, "gt_orb_get_candidates: bad index");
  const int slot = which ? e->cfg.max_batch : b;
  GT_CUDA(e, cudaStreamSynchronize(e->stream));
  int cnt = 0;
  GT_CUDA(e, cudaMemcpy(&cnt, e->fast_count + slot * GT_ORB_LEVELS + level, sizeof(int), cudaMemcpyDefault));
  cnt = std::min(cnt, e->lv[level].cand_cap);
  *n = cnt;
  const int m = std::min(cnt, max_n);
  const size_t off = (size_t)slot * e->cand_total + e->lv[level].cand_off;
  if (out_xy && m) GT_CUDA(e, cudaMemcpy(out_xy, e->fast_cand + off, (size_t)m * 4, cudaMemcpyDefault));
  if (out_score && m) GT_CUDA(e, cudaMemcpy(out_score, e->fast_score + off, (size_t)m, cudaMemcpyDefault));
  return GT_OK;
}

int gt_orb_detect(gt_handle e, const uint8_t* gray, const uint8_t* mask, int B, int as_reference, void* stream) {
  ENTER(e);
  GT_CHECK(e, gray && B >= 1 && B <= e->cfg.max_batch && (!as_reference || B == 1), "gt_orb_detect: bad arguments");
  cudaStream_t st = pick_stream(e, stream);
  const int slot0 = as_reference ? e->cfg.max_batch : 0;
  const size_t img = (size_t)e->work_h * e->work_w;
  GT_CUDA(e, cudaMemcpy2DAsync(e->pyr + (size_t)slot0 * e->pyr_bytes, e->pyr_bytes, gray, img, img, B, cudaMemcpyDefault, st));
  if (mask) GT_CUDA(e, cudaMemcpy2DAsync(e->pyr_mask + (size_t)slot0 * e->pyr_bytes, e->pyr_bytes, mask, img, img, B, cudaMemcpyDefault, st));
  else GT_CUDA(e, cudaMemset2DAsync(e->pyr_mask + (size_t)slot0 * e->pyr_bytes, e->pyr_bytes, 255, img, B, st));
  GT_TRY(orb_run(e, slot0, B, as_reference != 0, false, st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  if (as_reference) e->have_ref = true;
  return GT_OK;
}

int gt_match(gt_handle e, const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* out_idx, int32_t* out_dist, void* stream) {
  ENTER(e);
  GT_CHECK(e, query && train && nq >= 1 && nt >= 1 && nq <= GT_MAX_KP && nt <= GT_MAX_KP && out_idx && out_dist, "gt_match: bad arguments");
  cudaStream_t st = pick_stream(e, stream);
  uint8_t* dq = e->desc_all;                                        // slot 0
  uint8_t* dt = e->desc_all + (size_t)e->cfg.max_batch * GT_MAX_KP * 32;  // reference slot
  GT_TRY(to_device(e, dq, query, (size_t)nq * 32, st));
  GT_TRY(to_device(e, dt, train, (size_t)nt * 32, st));
  int h_n[2] = {nq, nt};
  GT_CUDA(e, cudaMemcpyAsync(e->kp_count, &h_n[0], sizeof(int), cudaMemcpyHostToDevice, st));
  GT_CUDA(e, cudaMemcpyAsync(e->kp_count + e->cfg.max_batch, &h_n[1], sizeof(int), cudaMemcpyHostToDevice, st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  GT_TRY(match_run(e, dq, e->kp_count, nq, dt, e->kp_count + e->cfg.max_batch, nt, e->match_idx, e->match_dist, 1, 0, 0, st));
  GT_TRY(to_caller(e, out_idx, e->match_idx, (size_t)nq * 2 * sizeof(int), st));
  GT_TRY(to_caller(e, out_dist, e->match_dist, (size_t)nq * 2 * sizeof(int), st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  return GT_OK;
}

int gt_match_l2(gt_handle e, const float* query, int nq, const float* train, int nt, int dim, int32_t* out_idx, float* out_dist, void* stream) {
  ENTER(e);
  GT_CHECK(e, query && train && out_idx && out_dist && nq >= 1 && nt >= 1 && nq <= GT_MATCH_L2_MAX && nt <= GT_MATCH_L2_MAX,
           "gt_match_l2: bad arguments (nq=%d, nt=%d, limit %d)", nq, nt, GT_MATCH_L2_MAX);
  GT_CHECK(e, dim == 128, "gt_match_l2: only 128-element descriptors (SIFT / RootSIFT) are supported, got %d", dim);
  cudaStream_t st = pick_stream(e, stream);
  const int need = std::max(nq, nt);
  if (need > e->reg_io_cap) {
    for (void* p : {(void*)e->reg_q32, (void*)e->reg_t32, (void*)e->reg_idx, (void*)e->reg_dist}) if (p) cudaFree(p);
    e->reg_q32 = e->reg_t32 = nullptr; e->reg_idx = nullptr; e->reg_dist = nullptr; e->reg_io_cap = 0;
    GT_CUDA(e, cudaMalloc((void**)&e->reg_q32, (size_t)need * 128 * sizeof(float)));
    GT_CUDA(e, cudaMalloc((void**)&e->reg_t32, (size_t)need * 128 * sizeof(float)));
    GT_CUDA(e, cudaMalloc((void**)&e->reg_idx, (size_t)need * 2 * sizeof(int)));
    GT_CUDA(e, cudaMalloc((void**)&e->reg_dist, (size_t)need * 2 * sizeof(float)));
    e->reg_io_cap = need;
  }
  GT_TRY(to_device(e, e->reg_q32, query, (size_t)nq * 128 * sizeof(float), st));
  GT_TRY(to_device(e, e->reg_t32, train, (size_t)nt * 128 * sizeof(float), st));
  GT_TRY(match_l2_run(e, e->reg_q32, nq, e->reg_t32, nt, e->reg_idx, e->reg_dist, st));
  GT_TRY(to_caller(e, out_idx, e->reg_idx, (size_t)nq * 2 * sizeof(int), st));
  GT_TRY(to_caller(e, out_dist, e->reg_dist, (size_t)nq * 2 * sizeof(float), st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  return GT_OK;
}

int gt_find_homography(gt_handle e, const float* src, const float* dst, int n, float thr, int max_iter, double* out_H, int32_t* out_inliers,
                       void* stream) {
  ENTER(e);
  const int pair_cap = e->cfg.max_batch * GT_MAX_KP;   // a single pair set may use the pair buffers of the whole batch
  GT_CHECK(e, src && dst && n >= 0 && n <= pair_cap && out_H && max_iter >= 1 && max_iter <= e->cfg.ransac_max_iter,
           "gt_find_homography: bad arguments (n=%d of at most %d, max_iter=%d of at most %d)", n, pair_cap, max_iter, e->cfg.ransac_max_iter);
  cudaStream_t st = pick_stream(e, stream);
  std::vector<float> pr((size_t)std::max(n, 1) * 4);
  for (int i = 0; i < n; ++i) { pr[i * 4] = src[i * 2]; pr[i * 4 + 1] = src[i * 2 + 1]; pr[i * 4 + 2] = dst[i * 2]; pr[i * 4 + 3] = dst[i * 2 + 1]; }
  GT_CUDA(e, cudaMemcpyAsync(e->pairs, pr.data(), (size_t)n * 16, cudaMemcpyHostToDevice, st));
  GT_CUDA(e, cudaMemcpyAsync(e->pair_count, &n, sizeof(int), cudaMemcpyHostToDevice, st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  GT_TRY(homography_run(e, e->pairs, e->pair_count, 1, pair_cap, thr, max_iter, e->H_dev, e->H_status, e->H_stats, 1.0f, true, nullptr, st));
  int stats[4], status = 0;
  GT_CUDA(e, cudaMemcpyAsync(out_H, e->H_dev, 9 * sizeof(double), cudaMemcpyDefault, st));
  GT_CUDA(e, cudaMemcpyAsync(stats, e->H_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
  GT_CUDA(e, cudaMemcpyAsync(&status, e->H_status, sizeof(int), cudaMemcpyDeviceToHost, st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  if (out_inliers) *out_inliers = stats[3];
  return status == 0 ? GT_OK : 1;
}

// boxes (x1,y1,x2,y2,...) -> xywh mask boxes for the stabilizer, on device
__global__ void dets_to_xywh_kernel(const float* __restrict__ det, const int* __restrict__ cnt, int row, int max_det, float* __restrict__ xywh,
                                    int* __restrict__ nb, int B, int obb) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = cnt[b];
  if (i == 0) nb[b] = n;
  if (i >= n) return;
  const float* d = det + ((size_t)b * max_det + i) * row;
  float* o = xywh + ((size_t)b * max_det + i) * 4;
  if (!obb) {
    o[0] = (d[0] + d[2]) * 0.5f; o[1] = (d[1] + d[3]) * 0.5f; o[2] = d[2] - d[0]; o[3] = d[3] - d[1];
  } else {  // axis-aligned envelope of the rotated box
    const float c = fabsf(cosf(d[4])), s = fabsf(sinf(d[4]));
    o[0] = d[0]; o[1] = d[1]; o[2] = d[2] * c + d[3] * s; o[3] = d[2] * s + d[3] * c;
  }
}

static int extract_batch_impl(gt_handle e, const uint8_t* frames, int B, int first_is_reference, float conf, float iou, int agnostic,
                              uint32_t classes_mask, const float* mask_boxes, const int32_t* mask_nboxes, int mask_stride, float* out_boxes,
                              int32_t* out_counts, float* out_boxes_stab, double* out_H, int32_t* out_status, int32_t* out_stats, void* stream,
                              bool sync) {
  GT_CHECK(e, frames && B >= 1 && B <= e->cfg.max_batch, "gt_extract_batch: bad batch %d", B);
  cudaStream_t st = pick_stream(e, stream);
  const int md = e->cfg.max_det;
  const int obb = e->cfg.task == GT_TASK_OBB;
  const bool ext_mask = mask_boxes && mask_nboxes;
  if (ext_mask) GT_TRY(upload_boxes(e, 0, B, mask_boxes, mask_nboxes, mask_stride, st));   // small inputs first ...
  if (e->deferred_src) {                                                                   // ... then the next batch's frames
    const uint8_t* nxt = e->deferred_src;
    e->deferred_src = nullptr;
    GT_TRY(gt_prefetch_frames(e, nxt, e->deferred_B));
  }
  GT_TRY(gt_preprocess(e, frames, B, st));
  // The mask-independent half of ORB (pyramid + FAST) needs only the gray frames.  overlap 2 (default): it runs on the aux stream beside
  // decode + NMS, whose kernels occupy 16-64 blocks; overlap 1: beside the whole detector (no gain: the conv CTAs own the SMs); 0: serial.
  const bool ov = e->overlap != 0;
  auto fork_front = [&](int part) -> int {   // part 0: pyramid + FAST, 1: pyramid only, 2: FAST only (the pyramid is already queued on the aux stream)
    if (part != 2) {
      GT_CUDA(e, cudaEventRecord(e->ev_pre, st));
      GT_CUDA(e, cudaStreamWaitEvent(e->aux_stream, e->ev_pre, 0));
    }
    if (part == 0) {
      if (e->front_split) GT_TRY(orb_front_split(e, 0, B, e->aux_stream, e->aux2_stream, e->ev_aux_a, e->ev_aux_b));
      else GT_TRY(orb_front(e, 0, B, e->aux_stream));
    } else if (part == 1) GT_TRY(orb_pyramid(e, 0, B, e->aux_stream));
    else GT_TRY(orb_fast(e, 0, B, e->aux_stream, 0, GT_ORB_LEVELS, true));
    if (part != 1) { GT_CUDA(e, cudaEventRecord(e->ev_front, e->aux_stream)); tl_mark(4, e->aux_stream); }
    return GT_OK;
  };
  tl_mark(0, st);
  if (e->overlap == 1) GT_TRY(fork_front(0));
  if (e->overlap == 3) GT_TRY(fork_front(1));
  GT_CUDA(e, cudaEventRecord(e->ev[2], st));
  GT_TRY(detector_forward(e, B, st));
  GT_CUDA(e, cudaEventRecord(e->ev[3], st));
  tl_mark(1, st);
  if (e->overlap == 2) GT_TRY(fork_front(0));
  if (e->overlap == 3) {   // FAST starts when the conv stack has finished (and, by stream order, after the pyramid)
    GT_CUDA(e, cudaEventRecord(e->ev_pre, st));
    GT_CUDA(e, cudaStreamWaitEvent(e->aux_stream, e->ev_pre, 0));
    GT_TRY(fork_front(2));
  }
  GT_TRY(detector_postprocess(e, B, conf, iou, agnostic, classes_mask, st));
  if (!e->post_event_done) GT_CUDA(e, cudaEventRecord(e->ev[4], st));   // (else nms_run recorded it right behind its last kernel)
  e->post_event_done = false;
  tl_mark(2, st);
  dim3 g((unsigned)ceil_div(md, 256), (unsigned)B);
  dets_to_xywh_kernel<<<g, 256, 0, st>>>(e->det_out, e->det_count, obb ? 7 : 6, md, e->det_xywh_dev, e->det_nbox_dev, B, obb);
  e->launches++;
  if (!ext_mask) {
    GT_CUDA(e, cudaMemcpyAsync(e->boxes_dev, e->det_xywh_dev, (size_t)B * md * 16, cudaMemcpyDeviceToDevice, st));
    GT_CUDA(e, cudaMemcpyAsync(e->nboxes_dev, e->det_nbox_dev, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
  }
  if (first_is_reference) {
    const int R = e->cfg.max_batch;
    GT_CUDA(e, cudaMemcpyAsync(e->pyr + (size_t)R * e->pyr_bytes, e->pyr, (size_t)e->work_h * e->work_w, cudaMemcpyDeviceToDevice, st));
    GT_CUDA(e, cudaMemcpyAsync(e->boxes_dev + (size_t)R * md * 4, e->boxes_dev, (size_t)md * 16, cudaMemcpyDeviceToDevice, st));
    GT_CUDA(e, cudaMemcpyAsync(e->nboxes_dev + R, e->nboxes_dev, sizeof(int), cudaMemcpyDeviceToDevice, st));
    GT_TRY(orb_run(e, R, 1, true, true, st));
    e->have_ref = true;
  }
  GT_TRY(stabilize_impl(e, B, st, ov));
  GT_TRY(warp_boxes_run(e, e->H_dev, e->H_status, e->det_xywh_dev, e->boxes_stab_dev, e->det_nbox_dev, B, md, st));
  GT_TRY(copy_dets(e, B, out_boxes, out_counts, nullptr, st));
  GT_TRY(to_caller(e, out_boxes_stab, e->boxes_stab_dev, (size_t)B * md * 16, st));
  GT_TRY(to_caller(e, out_H, e->H_dev, (size_t)B * 9 * sizeof(double), st));
  GT_TRY(to_caller(e, out_status, e->H_status, (size_t)B * sizeof(int), st));
  GT_TRY(to_caller(e, out_stats, e->H_stats, (size_t)B * 4 * sizeof(int), st));
  if (!sync) return GT_OK;
  GT_CUDA(e, cudaStreamSynchronize(st));
  update_times(e);
  tl_print();
  return GT_OK;
}

int gt_extract_batch(gt_handle e, const uint8_t* frames, int B, int first_is_reference, float conf, float iou, int agnostic,
                     uint32_t classes_mask, const float* mask_boxes, const int32_t* mask_nboxes, int mask_stride, float* out_boxes,
                     int32_t* out_counts, float* out_boxes_stab, double* out_H, int32_t* out_status, int32_t* out_stats, void* stream) {
  ENTER(e);
  return extract_batch_impl(e, frames, B, first_is_reference, conf, iou, agnostic, classes_mask, mask_boxes, mask_nboxes, mask_stride, out_boxes,
                            out_counts, out_boxes_stab, out_H, out_status, out_stats, stream, true);
}

// Pipelined form: everything (kernels and read-backs into the caller's PINNED output buffers) is enqueued and the call returns a
// ticket (0 / 1) at once; the caller may enqueue the next batch before gt_wait(ticket) -- the GPU then never idles between batches
// (the synchronous call leaves it idle for the read-back, the host wake-up and the next launches: ~0.25 ms of a 5.6 ms step).
// At most two tickets are in flight; outputs of a ticket are valid after its gt_wait.
int gt_extract_batch_async(gt_handle e, const uint8_t* frames, int B, int first_is_reference, float conf, float iou, int agnostic,
                           uint32_t classes_mask, const float* mask_boxes, const int32_t* mask_nboxes, int mask_stride, float* out_boxes,
                           int32_t* out_counts, float* out_boxes_stab, double* out_H, int32_t* out_status, int32_t* out_stats, void* stream,
                           int32_t* ticket) {
  ENTER(e);
  GT_CHECK(e, ticket != nullptr, "gt_extract_batch_async: ticket is NULL");
  const int k = e->async_ticket;
  e->async_ticket ^= 1;
  if (e->ticket_pending[k]) {   // a third batch without gt_wait on the first: its staging buffers and event set are about to be reused
    GT_CUDA(e, cudaEventSynchronize(e->ev_done[k]));
    e->ticket_pending[k] = false;
  }
  e->ev = e->ev_sets[k];
  const int rc = extract_batch_impl(e, frames, B, first_is_reference, conf, iou, agnostic, classes_mask, mask_boxes, mask_nboxes, mask_stride, out_boxes,
                                    out_counts, out_boxes_stab, out_H, out_status, out_stats, stream, false);
  if (rc == GT_OK) {
    cudaError_t er = cudaEventRecord(e->ev_done[k], pick_stream(e, stream));
    if (er != cudaSuccess) { gt_set_error(e, "gt_extract_batch_async: %s", cudaGetErrorString(er)); e->ev = e->ev_sets[0]; return GT_ERR_CUDA; }
    e->ticket_pending[k] = true;
  }
  e->ev = e->ev_sets[0];
  *ticket = k;
  return rc;
}

int gt_wait(gt_handle e, int ticket) {
  ENTER(e);
  GT_CHECK(e, ticket == 0 || ticket == 1, "gt_wait: bad ticket %d", ticket);
  GT_CUDA(e, cudaEventSynchronize(e->ev_done[ticket]));
  e->ticket_pending[ticket] = false;
  e->ev = e->ev_sets[ticket];
  update_times(e);
  e->ev = e->ev_sets[0];
  return GT_OK;
}

int gt_stage_times(gt_handle e, float* ms4) {
  if (!e || !ms4) return GT_ERR_INVALID;
  for (int i = 0; i < 4; ++i) ms4[i] = e->stage_ms[i];
  return GT_OK;
}
int64_t gt_launch_count(gt_handle e) { return e ? e->launches : -1; }
int gt_conv_kernel_info(gt_handle e, int32_t* n_ops, int32_t* n_swapped) {
  if (!e) return GT_ERR_INVALID;
  int ns = 0;
  for (const ConvOp& op : e->conv_ops) ns += op.swapped;
  if (n_ops) *n_ops = (int)e->conv_ops.size();
  if (n_swapped) *n_swapped = ns;
  return GT_OK;
}
int gt_conv_pair_count(gt_handle e) {
  if (!e) return GT_ERR_INVALID;
  int n = 0;
  for (const ConvOp& op : e->conv_ops) n += op.pair;
  return n;
}
int gt_conv_stack_stats(gt_handle e, float* ms, double* flops) {
  if (!e) return GT_ERR_INVALID;
  if (ms) *ms = e->conv_ms;
  if (flops) *flops = e->conv_flops;
  return GT_OK;
}

}  // extern "C"
