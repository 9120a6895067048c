// Stage 1 (letterbox / normalise / gray) and stage 2 (YOLOv8s graph, DFL decode, confidence filter, NMS).
//
// Reference sites (un-vendored ultralytics reached from /root/reference/geotrax/extract.py:153; restated in SURVEY.md
// section 8a-3 .. 8a-6 and Appendix A-1/A-2): LetterBox + BasePredictor.preprocess, DetectionModel.forward (fused),
// Detect/OBB head + DFL + dist2bbox/dist2rbox, non_max_suppression (torchvision.ops.nms / nms_rotated), scale_boxes.
#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include <type_traits>

#include "engine.cuh"

// =====================================================================================================================
// Stage 1: fused letterbox (exact 1/2 decimation) + BGR->RGB -> 16-bit space-to-depth NHWC tensor for layer 0, and
// BGR2GRAY + 1/2 resize -> u8 (level 0 of the ORB pyramid).  One pass over the 24.9 MB frame feeds stages 2 and 3.
//
// Network input layout: [B][net_h/4][net_w/4][64] = 4x4 space-to-depth of the letterboxed image, channel = (r * 4 + c) * 4 +
// {R, G, B, 0} for row r / column c inside the block; values are the exact u8 pixel (0..255 is exact in fp16 and bf16) -- the
// 1/255 of BasePredictor.preprocess is layer 0's epilogue scale.  With this layout the stride-2 3x3 layer-0 convolution becomes
// a stride-1 2x2 convolution over 64 channels that produces a 2x2 block of output pixels x 32 channels (N = 128) per row: a
// well-shaped GEMM (K = 256, N = 128) for the tcgen05 kernels instead of K = 27, N = 32.
// One thread: 2 letterboxed rows x 8 pixels = 4 source rows x 48 B (three 128-bit loads per row) -> two 64-byte segments.
// =====================================================================================================================
__device__ __forceinline__ uint32_t gray15(uint32_t b, uint32_t g, uint32_t r) {
  return (9798u * r + 19235u * g + 3735u * b + 16384u) >> 15;  // OpenCV BGR2GRAY, 15-bit coefficients
}

__global__ void __launch_bounds__(256) preprocess_half_kernel(const uint8_t* __restrict__ frames, bf16* __restrict__ s2d,
                                                              uint8_t* __restrict__ gray, size_t gray_frame_stride, int B,
                                                              int H, int W, int net_h, int net_w, int pad_top, int pad_left,
                                                              int new_h, int new_w, int fp16) {
  const int groups = new_w >> 3;  // 8 output pixels per thread
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_frame = (long long)(new_h >> 1) * groups;
  if (idx >= per_frame * B) return;
  const int b = (int)(idx / per_frame);
  const int rem = (int)(idx - (long long)b * per_frame);
  const int oy2 = rem / groups, og = rem - oy2 * groups;
  const uint8_t* src = frames + ((size_t)b * H + 4 * oy2) * (size_t)W * 3 + (size_t)og * 48;
  __align__(16) uint32_t line[32];  // 4 s2d pixels x 16 channels x 16 bit
  __align__(8) uint8_t gy[2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r) {     // output row parity
    uint4 r0[3], r1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      r0[i] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(2 * r) * W * 3) + i);
      r1[i] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(2 * r + 1) * W * 3) + i);
    }
    const uint8_t* a = reinterpret_cast<const uint8_t*>(r0);
    const uint8_t* c = reinterpret_cast<const uint8_t*>(r1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = j * 6;
      const uint32_t b00 = a[o], g00 = a[o + 1], r00 = a[o + 2], b01 = a[o + 3], g01 = a[o + 4], r01 = a[o + 5];
      const uint32_t b10 = c[o], g10 = c[o + 1], r10 = c[o + 2], b11 = c[o + 3], g11 = c[o + 4], r11 = c[o + 5];
      const uint32_t bb = (b00 + b01 + b10 + b11 + 2) >> 2;  // cv2.resize INTER_LINEAR at exactly 1/2
      const uint32_t gg = (g00 + g01 + g10 + g11 + 2) >> 2;
      const uint32_t rr = (r00 + r01 + r10 + r11 + 2) >> 2;
      // block j>>2 (of the two 4-wide blocks this thread touches), column j&3, row r of this thread's row pair: 2 words per pixel
      const int w0 = (j >> 2) * 16 + (r * 4 + (j & 3)) * 2;
      line[w0] = pack2_act((float)rr, (float)gg, fp16);
      line[w0 + 1] = pack2_act((float)bb, 0.f, fp16);
      const uint32_t y00 = gray15(b00, g00, r00), y01 = gray15(b01, g01, r01), y10 = gray15(b10, g10, r10), y11 = gray15(b11, g11, r11);
      gy[r][j] = (uint8_t)((y00 + y01 + y10 + y11 + 2) >> 2);  // gray first, then the 1/2 resize (stabilo order)
    }
  }
  // letterboxed rows 2*oy2 + pad_top + {0,1} = rows (rb, rb + 1) of block row Y; pixels og*8 + pad_left .. +7 = blocks X, X + 1
  const int sh = net_h >> 2, sw = net_w >> 2;
  const int ly = 2 * oy2 + pad_top, Y = ly >> 2, rb = ly & 3, X = (og * 8 + pad_left) >> 2;
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    uint4* dst = reinterpret_cast<uint4*>(s2d + (((size_t)b * sh + Y) * sw + X + blk) * 64 + rb * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = reinterpret_cast<const uint4*>(line)[blk * 4 + i];
  }
  if (gray) {
    uint8_t* g = gray + (size_t)b * gray_frame_stride + (size_t)(2 * oy2) * new_w + og * 8;
    *reinterpret_cast<uint2*>(g) = *reinterpret_cast<const uint2*>(gy[0]);
    *reinterpret_cast<uint2*>(g + new_w) = *reinterpret_cast<const uint2*>(gy[1]);
  }
}

// ---- general geometry (any frame size / imgsz / downsample_ratio): table-driven restatement of cv2.resize(INTER_LINEAR) ------
// Tables per axis (host-built, resize_tables below): source index pair (i0, i1) and 11-bit coefficients (c0, c1) of every destination
// coordinate, exactly as OpenCV's resize() computes them (oracle/prepost.py:resize_linear_u8 is the pinned restatement).
// mode 0: identity copy, 1: exact 2x2 decimation (OpenCV reroutes it to INTER_AREA), 2: bilinear.
struct ResizeTabs { const int *x0, *x1, *a0, *a1, *y0, *y1, *b0, *b1; int mode; };

__device__ __forceinline__ uint32_t lin_u8(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, int a0, int a1, int b0, int b1) {
  const int S0 = a0 * (int)p00 + a1 * (int)p01, S1 = a0 * (int)p10 + a1 * (int)p11;
  return (uint32_t)((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2);
}

__global__ void __launch_bounds__(256) letterbox_general_kernel(const uint8_t* __restrict__ frames, bf16* __restrict__ s2d, int B, int H, int W,
                                                                int net_h, int net_w, int pad_top, int pad_left, int new_h, int new_w,
                                                                ResizeTabs t, int fp16) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_frame = (long long)new_h * new_w;
  if (idx >= per_frame * B) return;
  const int b = (int)(idx / per_frame);
  const int rem = (int)(idx - (long long)b * per_frame);
  const int oy = rem / new_w, ox = rem - oy * new_w;
  const uint8_t* f = frames + (size_t)b * H * W * 3;
  uint32_t v[3];
  if (t.mode == 0) {
    const uint8_t* p = f + ((size_t)oy * W + ox) * 3;
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
  } else if (t.mode == 1) {
    const uint8_t* p = f + ((size_t)(2 * oy) * W + 2 * ox) * 3;
    const uint8_t* q = p + (size_t)W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = ((uint32_t)p[c] + p[c + 3] + q[c] + q[c + 3] + 2) >> 2;
  } else {
    const int x0 = t.x0[ox], x1 = t.x1[ox], a0 = t.a0[ox], a1 = t.a1[ox];
    const int y0 = t.y0[oy], y1 = t.y1[oy], b0 = t.b0[oy], b1 = t.b1[oy];
    const uint8_t* r0 = f + (size_t)y0 * W * 3;
    const uint8_t* r1 = f + (size_t)y1 * W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = lin_u8(r0[x0 * 3 + c], r0[x1 * 3 + c], r1[x0 * 3 + c], r1[x1 * 3 + c], a0, a1, b0, b1);
  }
  // 4x4 space-to-depth destination: block (Y, X), row rb, column cb -> halves R, G, B, 0
  const int ly = oy + pad_top, lx = ox + pad_left;
  const int sh = net_h >> 2, sw = net_w >> 2;
  uint32_t* dst = reinterpret_cast<uint32_t*>(s2d + (((size_t)b * sh + (ly >> 2)) * sw + (lx >> 2)) * 64 + (ly & 3) * 16 + (lx & 3) * 4);
  dst[0] = pack2_act((float)v[2], (float)v[1], fp16);
  dst[1] = pack2_act((float)v[0], 0.f, fp16);
}

// stabilizer working image: gray (cvtColor BGR2GRAY) first, then the resize to (work_w, work_h) -- stabilo's order
__global__ void __launch_bounds__(256) gray_work_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ gray, size_t gray_frame_stride,
                                                        int B, int H, int W, int work_h, int work_w, ResizeTabs t) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_frame = (long long)work_h * work_w;
  if (idx >= per_frame * B) return;
  const int b = (int)(idx / per_frame);
  const int rem = (int)(idx - (long long)b * per_frame);
  const int oy = rem / work_w, ox = rem - oy * work_w;
  const uint8_t* f = frames + (size_t)b * H * W * 3;
  auto g = [&](int y, int x) { const uint8_t* p = f + ((size_t)y * W + x) * 3; return gray15(p[0], p[1], p[2]); };
  uint32_t v;
  if (t.mode == 0) v = g(oy, ox);
  else if (t.mode == 1) v = (g(2 * oy, 2 * ox) + g(2 * oy, 2 * ox + 1) + g(2 * oy + 1, 2 * ox) + g(2 * oy + 1, 2 * ox + 1) + 2) >> 2;
  else {
    const int x0 = t.x0[ox], x1 = t.x1[ox], y0 = t.y0[oy], y1 = t.y1[oy];
    v = lin_u8(g(y0, x0), g(y0, x1), g(y1, x0), g(y1, x1), t.a0[ox], t.a1[ox], t.b0[oy], t.b1[oy]);
  }
  gray[(size_t)b * gray_frame_stride + (size_t)oy * work_w + ox] = (uint8_t)v;
}

// OpenCV resize() coordinate / coefficient computation for one axis (imgproc/src/resize.cpp, INTER_LINEAR branch)
static void resize_axis(int n_src, int n_dst, bool vertical, std::vector<int>& i0, std::vector<int>& i1, std::vector<int>& c0, std::vector<int>& c1) {
  i0.resize(n_dst); i1.resize(n_dst); c0.resize(n_dst); c1.resize(n_dst);
  const double scale = 1.0 / ((double)n_dst / (double)n_src);
  for (int d = 0; d < n_dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (!vertical && (s < 0 || s >= n_src - 1)) f = 0.f;    // horizontal: fraction zeroed where the index is clamped; vertical keeps it
    i0[d] = std::min(std::max(s, 0), n_src - 1);
    i1[d] = std::min(std::max(s + 1, 0), n_src - 1);
    c0[d] = (int)nearbyintf((1.f - f) * 2048.f);
    c1[d] = (int)nearbyintf(f * 2048.f);
  }
}

static int upload_tabs(gt_engine* e, int src_w, int src_h, int dst_w, int dst_h, int* dev[8], int* mode) {
  *mode = (src_w == dst_w && src_h == dst_h) ? 0 : ((src_w == 2 * dst_w && src_h == 2 * dst_h) ? 1 : 2);
  std::vector<int> v[8];
  resize_axis(src_w, dst_w, false, v[0], v[1], v[2], v[3]);
  resize_axis(src_h, dst_h, true, v[4], v[5], v[6], v[7]);
  for (int i = 0; i < 8; ++i) {
    GT_TRY(e->dev_alloc((void**)&dev[i], v[i].size() * sizeof(int)));
    GT_CUDA(e, cudaMemcpy(dev[i], v[i].data(), v[i].size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  return GT_OK;
}

int detector_build_general_preprocess(gt_engine* e) {
  GT_TRY(upload_tabs(e, e->cfg.frame_w, e->cfg.frame_h, e->new_w, e->new_h, e->lb_tab, &e->lb_mode));
  GT_TRY(upload_tabs(e, e->cfg.frame_w, e->cfg.frame_h, e->work_w, e->work_h, e->gw_tab, &e->gw_mode));
  return GT_OK;
}

// ---- NV12 -> BGR24 (decoder-format ingest): OpenCV's cvtColor(COLOR_YUV2BGR_NV12) integer arithmetic, bit-exact ----------------
// One thread = 2 x 2 luma pixels sharing one (U, V) pair: reads 2 + 2 luma bytes and 2 chroma bytes, writes 2 x 6 BGR bytes.
__global__ void __launch_bounds__(256) nv12_to_bgr_kernel(const uint8_t* __restrict__ nv12, uint8_t* __restrict__ bgr, int B, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int hw2 = W >> 1, hh2 = H >> 1;
  const long long per_frame = (long long)hw2 * hh2;
  if (idx >= per_frame * B) return;
  const int b = (int)(idx / per_frame);
  const int rem = (int)(idx - (long long)b * per_frame);
  const int y2 = rem / hw2, x2 = rem - y2 * hw2;
  const uint8_t* f = nv12 + (size_t)b * H * W * 3 / 2;
  const uint8_t* uvp = f + (size_t)H * W + (size_t)y2 * W + 2 * x2;
  const int u = (int)uvp[0] - 128, v = (int)uvp[1] - 128;
  constexpr int CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, SH = 20;
  const int ruv = (1 << (SH - 1)) + CVR * v, guv = (1 << (SH - 1)) + CVG * v + CUG * u, buv = (1 << (SH - 1)) + CUB * u;
  uint8_t* o = bgr + (size_t)b * H * W * 3;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint8_t* yp = f + (size_t)(2 * y2 + r) * W + 2 * x2;
    uint8_t* op = o + ((size_t)(2 * y2 + r) * W + 2 * x2) * 3;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int yy = max(0, (int)yp[c] - 16) * CY;
      op[c * 3 + 0] = (uint8_t)min(max((yy + buv) >> SH, 0), 255);
      op[c * 3 + 1] = (uint8_t)min(max((yy + guv) >> SH, 0), 255);
      op[c * 3 + 2] = (uint8_t)min(max((yy + ruv) >> SH, 0), 255);
    }
  }
}

// Fused NV12 ingest for the default geometry (exact 1/2 letterbox + 1/2 working image): one output pixel = one 2 x 2 luma block, which
// shares ONE chroma pair, so a thread reads 2 x 16 luma + 16 chroma bytes for 8 output pixels (1.5 B per source pixel instead of the
// 3 B of the BGR path) and no BGR frame is ever written.  Arithmetic = cvtColor(NV12 -> BGR) per source pixel, then exactly what
// preprocess_half_kernel does (2 x 2 average; gray first, then its 2 x 2 average), so the result is bit-identical to the two-pass path.
__global__ void __launch_bounds__(256) preprocess_nv12_half_kernel(const uint8_t* __restrict__ nv12, bf16* __restrict__ s2d, uint8_t* __restrict__ gray,
                                                                   size_t gray_frame_stride, int B, int H, int W, int net_h, int net_w, int pad_top,
                                                                   int pad_left, int new_h, int new_w, int fp16) {
  const int groups = new_w >> 3;  // 8 output pixels per thread
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_frame = (long long)(new_h >> 1) * groups;
  if (idx >= per_frame * B) return;
  const int b = (int)(idx / per_frame);
  const int rem = (int)(idx - (long long)b * per_frame);
  const int oy2 = rem / groups, og = rem - oy2 * groups;
  const uint8_t* f = nv12 + (size_t)b * H * W * 3 / 2;
  constexpr int CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, SH = 20;
  __align__(16) uint32_t line[32];
  __align__(8) uint8_t gy[2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r) {     // output row parity: source rows 4 oy2 + 2 r + {0, 1}, chroma row 2 oy2 + r
    const uint4 y0 = __ldg(reinterpret_cast<const uint4*>(f + (size_t)(4 * oy2 + 2 * r) * W + og * 16));
    const uint4 y1 = __ldg(reinterpret_cast<const uint4*>(f + (size_t)(4 * oy2 + 2 * r + 1) * W + og * 16));
    const uint4 uv = __ldg(reinterpret_cast<const uint4*>(f + (size_t)H * W + (size_t)(2 * oy2 + r) * W + og * 16));
    const uint8_t* a = reinterpret_cast<const uint8_t*>(&y0);
    const uint8_t* c = reinterpret_cast<const uint8_t*>(&y1);
    const uint8_t* q = reinterpret_cast<const uint8_t*>(&uv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int u = (int)q[2 * j] - 128, v = (int)q[2 * j + 1] - 128;
      const int ruv = (1 << (SH - 1)) + CVR * v, guv = (1 << (SH - 1)) + CVG * v + CUG * u, buv = (1 << (SH - 1)) + CUB * u;
      uint32_t bs = 0, gs = 0, rs = 0, ys = 0;
      const int yv[4] = {a[2 * j], a[2 * j + 1], c[2 * j], c[2 * j + 1]};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yy = max(0, yv[k] - 16) * CY;
        const uint32_t bb = (uint32_t)min(max((yy + buv) >> SH, 0), 255), gg = (uint32_t)min(max((yy + guv) >> SH, 0), 255),
                       rr = (uint32_t)min(max((yy + ruv) >> SH, 0), 255);
        bs += bb; gs += gg; rs += rr; ys += gray15(bb, gg, rr);
      }
      const int w0 = (j >> 2) * 16 + (r * 4 + (j & 3)) * 2;
      line[w0] = pack2_act((float)((rs + 2) >> 2), (float)((gs + 2) >> 2), fp16);
      line[w0 + 1] = pack2_act((float)((bs + 2) >> 2), 0.f, fp16);
      gy[r][j] = (uint8_t)((ys + 2) >> 2);
    }
  }
  const int sh = net_h >> 2, sw = net_w >> 2;
  const int ly = 2 * oy2 + pad_top, Y = ly >> 2, rb = ly & 3, X = (og * 8 + pad_left) >> 2;
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    uint4* dst = reinterpret_cast<uint4*>(s2d + (((size_t)b * sh + Y) * sw + X + blk) * 64 + rb * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = reinterpret_cast<const uint4*>(line)[blk * 4 + i];
  }
  if (gray) {
    uint8_t* g = gray + (size_t)b * gray_frame_stride + (size_t)(2 * oy2) * new_w + og * 8;
    *reinterpret_cast<uint2*>(g) = *reinterpret_cast<const uint2*>(gy[0]);
    *reinterpret_cast<uint2*>(g + new_w) = *reinterpret_cast<const uint2*>(gy[1]);
  }
}

int detector_preprocess_nv12(gt_engine* e, const uint8_t* nv12_dev, int B, cudaStream_t st) {
  const long long threads = (long long)B * (e->new_h / 2) * (e->new_w / 8);
  preprocess_nv12_half_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(nv12_dev, e->net_s2d, e->pyr, e->pyr_bytes, B, e->cfg.frame_h, e->cfg.frame_w,
                                                                                 e->net_h, e->net_w, e->pad_top, e->pad_left, e->new_h, e->new_w,
                                                                                 e->cfg.act_dtype == GT_ACT_FP16);
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  e->cur_frames = nullptr;
  return GT_OK;
}

int detector_nv12_to_bgr(gt_engine* e, const uint8_t* nv12_dev, uint8_t* bgr_dev, int B, cudaStream_t st) {
  const long long n = (long long)B * (e->cfg.frame_h / 2) * (e->cfg.frame_w / 2);
  nv12_to_bgr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nv12_dev, bgr_dev, B, e->cfg.frame_h, e->cfg.frame_w);
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

__global__ void fill_s2d_kernel(uint2* __restrict__ s2d, size_t n_quads, uint32_t w0, uint32_t w1) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_quads) s2d[i] = make_uint2(w0, w1);
}

int detector_fill_pad(gt_engine* e, cudaStream_t st) {  // constant letterbox border (value 114); the interior is rewritten per batch
  const int fp16 = e->cfg.act_dtype == GT_ACT_FP16;
  const uint16_t v = host_to_act(114.f, fp16);
  const size_t n_quads = (size_t)e->cfg.max_batch * (e->net_h / 4) * (e->net_w / 4) * 16;  // one quad = R,G,B,0
  fill_s2d_kernel<<<(unsigned)((n_quads + 255) / 256), 256, 0, st>>>(reinterpret_cast<uint2*>(e->net_s2d), n_quads, (uint32_t)v | ((uint32_t)v << 16),
                                                                      (uint32_t)v);
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

int detector_preprocess(gt_engine* e, const uint8_t* frames_dev, int B, cudaStream_t st) {
  // level 0 of each frame's pyramid slab receives the gray working image (CLAHE: written by clahe_run instead)
  uint8_t* gray = e->cfg.clahe ? nullptr : e->pyr;
  const int fp16 = e->cfg.act_dtype == GT_ACT_FP16;
  const ResizeTabs gt = {e->gw_tab[0], e->gw_tab[1], e->gw_tab[2], e->gw_tab[3], e->gw_tab[4], e->gw_tab[5], e->gw_tab[6], e->gw_tab[7], e->gw_mode};
  const long long n2 = (long long)B * e->work_h * e->work_w;
  if (e->lb_fast) {   // exact-1/2 letterbox: the fused vector kernel (+ the 1/2 gray working image in the same pass for the default preset)
    const long long threads = (long long)B * (e->new_h / 2) * (e->new_w / 8);
    preprocess_half_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(frames_dev, e->net_s2d, e->pre_fast ? gray : nullptr, e->pyr_bytes, B, e->cfg.frame_h,
                                                                              e->cfg.frame_w, e->net_h, e->net_w, e->pad_top, e->pad_left,
                                                                              e->new_h, e->new_w, fp16);
    e->launches++;
    if (!e->pre_fast && gray) {
      gray_work_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(frames_dev, gray, e->pyr_bytes, B, e->cfg.frame_h, e->cfg.frame_w, e->work_h, e->work_w, gt);
      e->launches++;
    }
  } else {            // any other geometry: two table-driven kernels (letterbox, gray working image)
    const ResizeTabs lt = {e->lb_tab[0], e->lb_tab[1], e->lb_tab[2], e->lb_tab[3], e->lb_tab[4], e->lb_tab[5], e->lb_tab[6], e->lb_tab[7], e->lb_mode};
    const long long n1 = (long long)B * e->new_h * e->new_w;
    letterbox_general_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(frames_dev, e->net_s2d, B, e->cfg.frame_h, e->cfg.frame_w, e->net_h, e->net_w,
                                                                           e->pad_top, e->pad_left, e->new_h, e->new_w, lt, fp16);
    e->launches++;
    if (gray) {
      gray_work_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(frames_dev, gray, e->pyr_bytes, B, e->cfg.frame_h, e->cfg.frame_w, e->work_h, e->work_w, gt);
      e->launches++;
    }
  }
  GT_CUDA(e, cudaGetLastError());
  e->cur_frames = frames_dev;
  return GT_OK;
}

// =====================================================================================================================
// SPPF max pool 5x5 / stride 1 / pad 2 on a channel slice (NHWC, 8 channels per thread)
// =====================================================================================================================
template <typename T2>
__global__ void __launch_bounds__(256) maxpool5_kernel(const bf16* __restrict__ in, int in_ctot, int in_coff, bf16* __restrict__ out,
                                                       int out_ctot, int out_coff, int B, int H, int W, int C) {
  const int c8 = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W * c8;
  if (idx >= total) return;
  const int cg = (int)(idx % c8);
  const int x = (int)((idx / c8) % W);
  const int y = (int)((idx / ((long long)c8 * W)) % H);
  const int b = (int)(idx / ((long long)c8 * W * H));
  T2 m[4];
  T2 ninf;
  if constexpr (sizeof(T2) == 4 && std::is_same<T2, __half2>::value) ninf = __float2half2_rn(-INFINITY); else ninf = __float2bfloat162_rn(-INFINITY);
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = ninf;
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -2; dx <= 2; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * H + yy) * W + xx) * in_ctot + in_coff + cg * 8));
      const T2* pv = reinterpret_cast<const T2*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], pv[i]);
    }
  }
  *reinterpret_cast<uint4*>(out + (((size_t)b * H + y) * W + x) * out_ctot + out_coff + cg * 8) = *reinterpret_cast<const uint4*>(m);
}

// SPPF chain in one launch: the three chained 5x5 / stride-1 pools (= 5x5, 9x9 and 13x13 windows of the input) of one image and one
// 8-channel group, separable (row max, column max) in shared memory; the plane of a 34 x 60 level is 32 KB.  Replaces three
// dependent launches that were pure latency (33 us each for an 8 MB tensor).
template <typename T2>
__global__ void __launch_bounds__(1024) sppf_pool3_kernel(bf16* __restrict__ buf, int ctot, int coff, int C, int H, int W) {
  extern __shared__ __align__(16) uint4 s_pool[];
  uint4* A = s_pool;
  uint4* R = s_pool + (size_t)H * W;
  const int cg = blockIdx.x, b = blockIdx.y;
  const int n = H * W;
  bf16* base = buf + (size_t)b * n * ctot + coff + cg * 8;
  for (int i = threadIdx.x; i < n; i += blockDim.x) A[i] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)i * ctot));
  __syncthreads();
  auto vmax = [](uint4 a, const uint4& c) {
    T2* pa = reinterpret_cast<T2*>(&a);
    const T2* pc = reinterpret_cast<const T2*>(&c);
#pragma unroll
    for (int k = 0; k < 4; ++k) pa[k] = __hmax2(pa[k], pc[k]);
    return a;
  };
  for (int stage = 1; stage <= 3; ++stage) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {       // row pass
      const int y = i / W, x = i - y * W;
      uint4 m = A[i];
      for (int dx = -2; dx <= 2; ++dx)
        if (dx != 0 && x + dx >= 0 && x + dx < W) m = vmax(m, A[i + dx]);
      R[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {       // column pass -> slice `stage`, and the next stage's input
      const int y = i / W;
      uint4 m = R[i];
      for (int dy = -2; dy <= 2; ++dy)
        if (dy != 0 && y + dy >= 0 && y + dy < H) m = vmax(m, R[i + dy * W]);
      A[i] = m;
      *reinterpret_cast<uint4*>(base + (size_t)i * ctot + (size_t)stage * C) = m;
    }
    __syncthreads();
  }
}

// =====================================================================================================================
// Decode + confidence filter: raw head rows [B][A][no] f32 -> dense candidate arrays indexed by anchor + a key list.
// key = conf bits << 32 | ~anchor : descending key order == (conf desc, anchor asc) == the order torchvision.ops.nms's
// stable descending sort gives to ultralytics' anchor-ordered candidate rows.
// =====================================================================================================================
struct ClassMask {            // allow-list over up to 96 classes (gt_create accepts nc <= 80); on == 0: every class passes
  uint32_t w[3];
  int on;
  __device__ __forceinline__ bool allows(int c) const { return !on || ((w[c >> 5] >> (c & 31)) & 1u); }
};
static ClassMask make_class_mask(const gt_engine* e, uint32_t classes_mask) {
  ClassMask m = {{0u, 0u, 0u}, 0};
  if (classes_mask) { m.w[0] = classes_mask; m.on = 1; }                         // per-call mask (classes 0..31) wins
  else if (e->cls_filter_on) { for (int i = 0; i < 3; ++i) m.w[i] = e->cls_filter[i]; m.on = 1; }   // gt_set_class_filter
  return m;
}

struct DecodeGeom {
  int lvl_w[3], lvl_h[3], lvl_off[3];
  float stride[3];
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float dfl_side(const float* p) {
  float m = p[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) m = fmaxf(m, p[i]);
  float s = 0.f, ws = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float ex = expf(p[i] - m);
    s += ex;
    ws += ex * (float)i;
  }
  return ws / s;
}

__global__ void __launch_bounds__(256) decode_filter_kernel(const float* __restrict__ raw_box, const float* __restrict__ raw_cls, const float* __restrict__ raw_ang,
                                                            int B, int A, int ncp, int nc, int obb,
                                                            DecodeGeom g, float conf_thr, ClassMask cm,
                                                            float* __restrict__ cand_box, float* __restrict__ cand_conf,
                                                            int* __restrict__ cand_cls, unsigned long long* __restrict__ keys,
                                                            int key_stride, int* __restrict__ count, int* __restrict__ nonfinite) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * A) return;
  const int b = (int)(idx / A), a = (int)(idx - (long long)b * A);
  // the class logits are a dense [anchor][ncp] array of their own: the confidence test of the ~99 % of anchors that fail it touches
  // 4 * ncp contiguous bytes per anchor, the 256-byte DFL row only for candidates
  const float* rc = raw_cls + (size_t)idx * ncp;
  const float* r = raw_box + (size_t)idx * 64;
  float best = rc[0];
  int bj = 0;
  bool bad = !isfinite(best);
  for (int j = 1; j < nc; ++j) {
    const float v = rc[j];
    bad |= !isfinite(v);
    if (v > best) { best = v; bj = j; }
  }
  // 16-bit overflow guard: an fp16 activation that left the format's range (|x| > 65504) reaches the head as inf / NaN.  Counted, never
  // silently dropped: gt_get_health() reports it and the Python front end switches the engine to bf16 storage.
  if (bad) { atomicAdd(nonfinite, 1); return; }
  const float conf = sigmoid_f(best);
  if (!(conf > conf_thr)) return;
  // the reference takes max/argmax over the *probabilities* (cls.sigmoid().max(1)): when the sigmoid saturates, several
  // classes tie at the same float and the first index wins
  for (int j = 0; j < bj; ++j)
    if (sigmoid_f(rc[j]) >= conf) { bj = j; break; }
  if (!cm.allows(bj)) return;
  int lvl = 0;
  if (a >= g.lvl_off[2]) lvl = 2; else if (a >= g.lvl_off[1]) lvl = 1;
  const int la = a - g.lvl_off[lvl];
  const float ax = (float)(la % g.lvl_w[lvl]) + 0.5f, ay = (float)(la / g.lvl_w[lvl]) + 0.5f;
  const float s = g.stride[lvl];
  float box[16];
  float d[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int i = 0; i < 16; ++i) box[i] = r[k * 16 + i];
    d[k] = dfl_side(box);
  }
  if (!isfinite(d[0] + d[1] + d[2] + d[3])) { atomicAdd(nonfinite, 1); return; }
  float* o = cand_box + ((size_t)b * A + a) * 5;
  if (!obb) {
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    const float cx = __fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), s), cy = __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), s);
    const float w = __fmul_rn(__fsub_rn(x2, x1), s), h = __fmul_rn(__fsub_rn(y2, y1), s);
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    o[0] = __fsub_rn(cx, hw); o[1] = __fsub_rn(cy, hh); o[2] = __fadd_rn(cx, hw); o[3] = __fadd_rn(cy, hh); o[4] = 0.f;
  } else {
    const float ang = (sigmoid_f(raw_ang[(size_t)idx * 4]) - 0.25f) * 3.14159265358979323846f;
    const float cs = cosf(ang), sn = sinf(ang);
    const float xf = (d[2] - d[0]) * 0.5f, yf = (d[3] - d[1]) * 0.5f;
    o[0] = (xf * cs - yf * sn + ax) * s; o[1] = (xf * sn + yf * cs + ay) * s;
    o[2] = (d[0] + d[2]) * s; o[3] = (d[1] + d[3]) * s; o[4] = ang;
  }
  cand_conf[(size_t)b * A + a] = conf;
  cand_cls[(size_t)b * A + a] = bj;
  const int slot = atomicAdd(&count[b], 1);
  keys[(size_t)b * key_stride + slot] = ((unsigned long long)__float_as_uint(conf) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)a);
}

// same filter for caller-supplied decoded predictions [B][A][4+nc(+1)] (gt_nms)
__global__ void __launch_bounds__(256) pred_filter_kernel(const float* __restrict__ pred, int B, int A, int nc, int rotated, float conf_thr,
                                                          ClassMask cm, float* __restrict__ cand_box, float* __restrict__ cand_conf,
                                                          int* __restrict__ cand_cls, unsigned long long* __restrict__ keys, int key_stride,
                                                          int* __restrict__ count) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * A) return;
  const int b = (int)(idx / A), a = (int)(idx - (long long)b * A);
  const int row = 4 + nc + (rotated ? 1 : 0);
  const float* r = pred + (size_t)idx * row;
  float best = r[4];
  int bj = 0;
  for (int j = 1; j < nc; ++j) {
    const float v = r[4 + j];
    if (v > best) { best = v; bj = j; }
  }
  if (!(best > conf_thr)) return;
  if (!cm.allows(bj)) return;
  float* o = cand_box + ((size_t)b * A + a) * 5;
  if (!rotated) {
    const float hw = __fmul_rn(r[2], 0.5f), hh = __fmul_rn(r[3], 0.5f);  // xywh2xyxy: xy -+ wh/2
    o[0] = __fsub_rn(r[0], hw); o[1] = __fsub_rn(r[1], hh); o[2] = __fadd_rn(r[0], hw); o[3] = __fadd_rn(r[1], hh); o[4] = 0.f;
  } else {
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3]; o[4] = r[4 + nc];
  }
  cand_conf[(size_t)b * A + a] = best;
  cand_cls[(size_t)b * A + a] = bj;
  const int slot = atomicAdd(&count[b], 1);
  keys[(size_t)b * key_stride + slot] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)a);
}

// ---- per-image descending bitonic sort of the key list (one CTA per image) -------------------------------------------
__global__ void __launch_bounds__(1024) sort_keys_kernel(unsigned long long* __restrict__ keys, int key_stride, const int* __restrict__ count) {
  extern __shared__ unsigned long long s_keys[];
  const int b = blockIdx.x;
  const int n = min(count[b], key_stride);
  if (n <= 1) return;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  unsigned long long* k = keys + (size_t)b * key_stride;
  const bool in_smem = np2 <= 4096;
  unsigned long long* w = in_smem ? s_keys : k;
  if (in_smem) {
    for (int i = threadIdx.x; i < np2; i += blockDim.x) s_keys[i] = i < n ? k[i] : 0ull;
  } else {
    for (int i = n + threadIdx.x; i < np2; i += blockDim.x) k[i] = 0ull;  // key_stride is a power of two >= A
  }
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = w[lo], c = w[hi];
        if ((a < c) == desc) { w[lo] = c; w[hi] = a; }
      }
      __syncthreads();
    }
  }
  if (in_smem)
    for (int i = threadIdx.x; i < n; i += blockDim.x) k[i] = s_keys[i];
}

// ---- gather sorted boxes (+ class offset) -----------------------------------------------------------------------------
__global__ void nms_gather_kernel(const unsigned long long* __restrict__ keys, int key_stride, const int* __restrict__ count, int A,
                                  int max_nms, const float* __restrict__ cand_box, const int* __restrict__ cand_cls, int agnostic,
                                  int rotated, float* __restrict__ sbox, int n_cap) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(min(count[b], max_nms), n_cap);
  if (i >= n) return;
  const unsigned long long key = keys[(size_t)b * key_stride + i];
  const int a = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
  const float* src = cand_box + ((size_t)b * A + a) * 5;
  const float c = agnostic ? 0.f : __fmul_rn((float)cand_cls[(size_t)b * A + a], 7680.0f);
  float* d = sbox + ((size_t)b * n_cap + i) * 5;
  d[0] = __fadd_rn(src[0], c); d[1] = __fadd_rn(src[1], c);
  if (rotated) { d[2] = src[2]; d[3] = src[3]; } else { d[2] = __fadd_rn(src[2], c); d[3] = __fadd_rn(src[3], c); }
  d[4] = src[4];
}

// ---- bitmask IoU matrix: mask[i][cb] bit j = IoU(i, cb*64+j) > thr, for j > i (torchvision semantics: strict >) -------
__device__ __forceinline__ bool iou_gt(const float* a, const float* b, float thr) {
  const float xx1 = fmaxf(a[0], b[0]), yy1 = fmaxf(a[1], b[1]);
  const float xx2 = fminf(a[2], b[2]), yy2 = fminf(a[3], b[3]);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  if (thr >= 0.f && (!(w > 0.f) || !(h > 0.f))) return false;   // disjoint boxes (almost every pair): inter = 0 -> 0 / union (or NaN) > thr is false; skips the IEEE division
  const float inter = __fmul_rn(w, h);
  const float aa = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
  const float ab = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
  return ovr > thr;
}

__device__ __forceinline__ void cov_abc(const float* o, float* a, float* b, float* c) {
  const float ga = o[2] * o[2] / 12.f, gb = o[3] * o[3] / 12.f;
  const float cs = cosf(o[4]), sn = sinf(o[4]);
  const float c2 = cs * cs, s2 = sn * sn;
  *a = ga * c2 + gb * s2;
  *b = ga * s2 + gb * c2;
  *c = (ga - gb) * cs * sn;
}

__device__ __forceinline__ float probiou(const float* o1, const float* o2) {
  const float eps = 1e-7f;
  float a1, b1, c1, a2, b2, c2;
  cov_abc(o1, &a1, &b1, &c1);
  cov_abc(o2, &a2, &b2, &c2);
  const float dx = o1[0] - o2[0], dy = o1[1] - o2[1];
  const float den = (a1 + a2) * (b1 + b2) - (c1 + c2) * (c1 + c2);
  const float t1 = (((a1 + a2) * dy * dy + (b1 + b2) * dx * dx) / (den + eps)) * 0.25f;
  const float t2 = (((c1 + c2) * (-dx) * dy) / (den + eps)) * 0.5f;
  const float d1 = fmaxf(a1 * b1 - c1 * c1, 0.f), d2 = fmaxf(a2 * b2 - c2 * c2, 0.f);
  const float t3 = logf(den / (4.f * sqrtf(d1 * d2) + eps) + eps) * 0.5f;
  const float bd = fminf(fmaxf(t1 + t2 + t3, eps), 100.f);
  const float hd = sqrtf(1.f - expf(-bd) + eps);
  return 1.f - hd;
}

__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ sbox, const int* __restrict__ count, int max_nms, int n_cap,
                                                      int words, float thr, unsigned long long* __restrict__ mask, int rotated) {
  const int b = blockIdx.z;
  const int n = min(min(count[b], max_nms), n_cap);
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb || rb * 64 >= n || cb * 64 >= n) return;
  __shared__ float s_col[64][5];
  const int cj = cb * 64 + threadIdx.x;
  if (cj < n) {
#pragma unroll
    for (int k = 0; k < 5; ++k) s_col[threadIdx.x][k] = sbox[((size_t)b * n_cap + cj) * 5 + k];
  }
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i >= n) return;
  float me[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) me[k] = sbox[((size_t)b * n_cap + i) * 5 + k];
  const int ncol = min(64, n - cb * 64);
  unsigned long long bits = 0ull;
  const int start = (rb == cb) ? threadIdx.x + 1 : 0;
  for (int j = start; j < ncol; ++j) {
    const bool sup = rotated ? (probiou(me, s_col[j]) >= thr) : iou_gt(me, s_col[j], thr);
    if (sup) bits |= 1ull << j;
  }
  mask[((size_t)b * n_cap + i) * words + cb] = bits;
}

// one kept candidate -> its det_out row: source-frame pixels (scale_boxes + clip / regularize_rboxes) or letterboxed pixels (scale == 0)
__device__ __forceinline__ void nms_emit_row(const float* __restrict__ bx, float cf, float cl, int rotated, int scale, float pad_x, float pad_y, float gain,
                                             float fw, float fh, float* __restrict__ o) {
  if (!rotated) {
    float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
    if (scale) {
      x1 = __fdiv_rn(__fsub_rn(x1, pad_x), gain); y1 = __fdiv_rn(__fsub_rn(y1, pad_y), gain);
      x2 = __fdiv_rn(__fsub_rn(x2, pad_x), gain); y2 = __fdiv_rn(__fsub_rn(y2, pad_y), gain);
      x1 = fminf(fmaxf(x1, 0.f), fw); x2 = fminf(fmaxf(x2, 0.f), fw);
      y1 = fminf(fmaxf(y1, 0.f), fh); y2 = fminf(fmaxf(y2, 0.f), fh);
    }
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = cf; o[5] = cl;
  } else {
    float x = bx[0], y = bx[1], w = bx[2], h = bx[3], t = bx[4];
    if (scale) {
      const float PI = 3.14159265358979323846f;
      float tm = fmodf(t, PI);
      if (tm < 0.f) tm += PI;
      const bool swap = tm >= PI * 0.5f;  // regularize_rboxes
      const float w2 = swap ? h : w, h2 = swap ? w : h;
      float t2 = fmodf(t, PI * 0.5f);
      if (t2 < 0.f) t2 += PI * 0.5f;
      x = __fdiv_rn(__fsub_rn(x, pad_x), gain); y = __fdiv_rn(__fsub_rn(y, pad_y), gain);
      w = __fdiv_rn(w2, gain); h = __fdiv_rn(h2, gain); t = t2;
    }
    o[0] = x; o[1] = y; o[2] = w; o[3] = h; o[4] = t; o[5] = cf; o[6] = cl;
  }
}

// ---- sweep + output.  HBB: greedy (a suppressed box does not suppress).  Rotated: Fast-NMS (every higher box suppresses).
// Writes det_out rows in source-frame pixels (scale_boxes + clip) or letterboxed pixels (scale == 0).
__global__ void __launch_bounds__(512) nms_sweep_kernel(const unsigned long long* __restrict__ mask, const float* __restrict__ sbox,
                                                        const unsigned long long* __restrict__ keys, int key_stride,
                                                        const int* __restrict__ count, int max_nms, int n_cap, int words, int A,
                                                        const float* __restrict__ cand_box, const float* __restrict__ cand_conf,
                                                        const int* __restrict__ cand_cls, int rotated, int max_det, int scale,
                                                        float pad_x, float pad_y, float gain, float fw, float fh, float* __restrict__ det_out,
                                                        int* __restrict__ det_count, int* __restrict__ det_keep, int* __restrict__ overflow) {
  extern __shared__ unsigned long long s_removed[];  // [words]
  __shared__ int s_kept[64];
  __shared__ unsigned long long s_diag[64];
  __shared__ int s_nk, s_total;
  const int b = blockIdx.x;
  const int ntrue = min(count[b], max_nms);
  if (ntrue > n_cap) {
    if (threadIdx.x == 0) { overflow[b] = 1; det_count[b] = 0; }
    return;
  }
  const int n = ntrue;
  const int w_used = (n + 63) >> 6;
  for (int i = threadIdx.x; i < w_used; i += blockDim.x) s_removed[i] = 0ull;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  const int row = rotated ? 7 : 6;
  for (int chunk = 0; chunk < w_used; ++chunk) {
    if (s_total >= max_det) break;
    // the 64 diagonal mask words of this chunk are fetched in parallel; the serial scan below then runs out of shared memory
    if (threadIdx.x < 64) {
      const int r = chunk * 64 + threadIdx.x;
      s_diag[threadIdx.x] = r < n ? mask[((size_t)b * n_cap + r) * words + chunk] : 0ull;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long rem = s_removed[chunk];
      int nk = 0;
      const int lim = min(64, n - chunk * 64);
      for (int j = 0; j < lim; ++j) {
        if (!((rem >> j) & 1ull)) {
          if (s_total + nk < max_det) s_kept[nk++] = chunk * 64 + j;
          if (!rotated) rem |= s_diag[j];
        }
        if (rotated) rem |= s_diag[j];
      }
      s_nk = nk;
    }
    __syncthreads();
    const int nk = s_nk;
    // propagate suppression to later chunks
    if (!rotated) {
      for (int wd = chunk + 1 + threadIdx.x; wd < w_used; wd += blockDim.x) {
        unsigned long long acc = 0ull;
        for (int k = 0; k < nk; ++k) acc |= mask[((size_t)b * n_cap + s_kept[k]) * words + wd];
        s_removed[wd] |= acc;
      }
    } else {
      const int lim = min(64, n - chunk * 64);
      for (int wd = chunk + 1 + threadIdx.x; wd < w_used; wd += blockDim.x) {
        unsigned long long acc = 0ull;
        for (int j = 0; j < lim; ++j) acc |= mask[((size_t)b * n_cap + chunk * 64 + j) * words + wd];
        s_removed[wd] |= acc;
      }
    }
    // emit kept rows of this chunk
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
      const int si = s_kept[k];
      const unsigned long long key = keys[(size_t)b * key_stride + si];
      const int a = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
      nms_emit_row(cand_box + ((size_t)b * A + a) * 5, cand_conf[(size_t)b * A + a], (float)cand_cls[(size_t)b * A + a], rotated, scale, pad_x, pad_y, gain,
                   fw, fh, det_out + ((size_t)b * max_det + s_total + k) * row);
      det_keep[(size_t)b * max_det + s_total + k] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_total += nk;
    __syncthreads();
  }
  if (threadIdx.x == 0) det_count[b] = s_total;
}

// ---- the common case in ONE launch: frames with at most kNmsSmall candidates (the default flight has ~320) ------------------------
// sort (bitonic, shared memory) -> gather -> suppression bit matrix (shared memory) -> sweep -> rows, one block per frame: the same
// arithmetic and the same order as sort_keys / nms_gather / nms_mask / nms_sweep (iou_gt / probiou, greedy for HBB, Fast-NMS for OBB),
// so keep indices stay bit-exact; it replaces four launches whose cost was launch geometry (65,536 mostly empty mask blocks for the
// 4096-candidate capacity) and global-memory round trips.  Frames with more candidates raise `over[b]` and the batch falls back to
// the multi-launch path.
constexpr int kNmsSmall = 1024, kNmsSmallWords = kNmsSmall / 64;
constexpr size_t kNmsSmallSmem = (size_t)kNmsSmall * 8 /*keys*/ + (size_t)kNmsSmall * 5 * 4 /*boxes*/ + (size_t)kNmsSmall * kNmsSmallWords * 8 /*mask*/;
__global__ void __launch_bounds__(1024) nms_small_kernel(unsigned long long* __restrict__ keys, int key_stride, const int* __restrict__ count, int max_nms, int A,
                                                         const float* __restrict__ cand_box, const float* __restrict__ cand_conf,
                                                         const int* __restrict__ cand_cls, int agnostic, int rotated, float thr, int max_det, int scale,
                                                         float pad_x, float pad_y, float gain, float fw, float fh, float* __restrict__ det_out,
                                                         int* __restrict__ det_count, int* __restrict__ det_keep, int* __restrict__ over) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(s_raw);
  float* s_box = reinterpret_cast<float*>(s_raw + (size_t)kNmsSmall * 8);
  unsigned long long* s_mask = reinterpret_cast<unsigned long long*>(s_raw + (size_t)kNmsSmall * 8 + (size_t)kNmsSmall * 5 * 4);
  __shared__ unsigned long long s_removed[kNmsSmallWords];
  __shared__ int s_kept[kNmsSmall];
  __shared__ int s_total;
  const int b = blockIdx.x;
  const int nall = min(count[b], key_stride);
  const int n = min(nall, max_nms);
  if (nall > kNmsSmall) {               // (sorting needs every candidate, not only the max_nms best)
    if (threadIdx.x == 0) { over[b] = 1; det_count[b] = 0; }
    return;
  }
  unsigned long long* k = keys + (size_t)b * key_stride;
  int np2 = 1;
  while (np2 < nall) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) s_keys[i] = i < nall ? k[i] : 0ull;
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = s_keys[lo], c = s_keys[hi];
        if ((a < c) == desc) { s_keys[lo] = c; s_keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < nall; i += blockDim.x) k[i] = s_keys[i];     // (the sorted order stays visible, as after sort_keys_kernel)
  // gather (class offset as in nms_gather_kernel)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int a = (int)(0xFFFFFFFFu - (uint32_t)(s_keys[i] & 0xFFFFFFFFull));
    const float* src = cand_box + ((size_t)b * A + a) * 5;
    const float c = agnostic ? 0.f : __fmul_rn((float)cand_cls[(size_t)b * A + a], 7680.0f);
    float* d = s_box + (size_t)i * 5;
    d[0] = __fadd_rn(src[0], c); d[1] = __fadd_rn(src[1], c);
    if (rotated) { d[2] = src[2]; d[3] = src[3]; } else { d[2] = __fadd_rn(src[2], c); d[3] = __fadd_rn(src[3], c); }
    d[4] = src[4];
  }
  const int w_used = (n + 63) >> 6;
  if (threadIdx.x < kNmsSmallWords) s_removed[threadIdx.x] = 0ull;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  // suppression bits: word (i, cb), cb >= i / 64, bit j = box i suppresses box cb * 64 + j (j beyond i only)
  for (int t = threadIdx.x; t < n * w_used; t += blockDim.x) {
    const int i = t / w_used, cb = t - i * w_used;
    unsigned long long bits = 0ull;
    if (cb >= (i >> 6)) {
      const float* me = s_box + (size_t)i * 5;
      const int ncol = min(64, n - cb * 64);
      const int start = (cb == (i >> 6)) ? (i & 63) + 1 : 0;
      for (int j = start; j < ncol; ++j) {
        const float* other = s_box + (size_t)(cb * 64 + j) * 5;
        const bool sup = rotated ? (probiou(me, other) >= thr) : iou_gt(me, other, thr);
        if (sup) bits |= 1ull << j;
      }
    }
    s_mask[(size_t)i * kNmsSmallWords + cb] = bits;
  }
  __syncthreads();
  // sweep: chunk by chunk, the serial part on one thread (shared memory only)
  __shared__ int s_nk;
  for (int chunk = 0; chunk < w_used; ++chunk) {
    if (s_total >= max_det) break;
    if (threadIdx.x == 0) {
      unsigned long long rem = s_removed[chunk];
      int nk = 0;
      const int lim = min(64, n - chunk * 64);
      for (int j = 0; j < lim; ++j) {
        const unsigned long long dg = s_mask[(size_t)(chunk * 64 + j) * kNmsSmallWords + chunk];
        if (!((rem >> j) & 1ull)) {
          if (s_total + nk < max_det) s_kept[s_total + nk++] = chunk * 64 + j;
          if (!rotated) rem |= dg;
        }
        if (rotated) rem |= dg;
      }
      s_nk = nk;
    }
    __syncthreads();
    const int nk = s_nk, base = s_total;
    if (!rotated) {
      for (int wd = chunk + 1 + threadIdx.x; wd < w_used; wd += blockDim.x) {
        unsigned long long acc = 0ull;
        for (int q = 0; q < nk; ++q) acc |= s_mask[(size_t)s_kept[base + q] * kNmsSmallWords + wd];
        s_removed[wd] |= acc;
      }
    } else {
      const int lim = min(64, n - chunk * 64);
      for (int wd = chunk + 1 + threadIdx.x; wd < w_used; wd += blockDim.x) {
        unsigned long long acc = 0ull;
        for (int j = 0; j < lim; ++j) acc |= s_mask[(size_t)(chunk * 64 + j) * kNmsSmallWords + wd];
        s_removed[wd] |= acc;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_total = base + nk;
    __syncthreads();
  }
  const int total = s_total, row = rotated ? 7 : 6;
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    const int a = (int)(0xFFFFFFFFu - (uint32_t)(s_keys[s_kept[q]] & 0xFFFFFFFFull));
    nms_emit_row(cand_box + ((size_t)b * A + a) * 5, cand_conf[(size_t)b * A + a], (float)cand_cls[(size_t)b * A + a], rotated, scale, pad_x, pad_y, gain, fw,
                 fh, det_out + ((size_t)b * max_det + q) * row);
    det_keep[(size_t)b * max_det + q] = a;
  }
  if (threadIdx.x == 0) det_count[b] = total;
}

int nms_run(gt_engine* e, const float* pred_dev, int B, int A, int nc, int rotated, float conf, float iou, int agnostic,
            uint32_t classes_mask, int max_det, bool scale_to_frame, cudaStream_t st) {
  // pred_dev == nullptr: candidates come from the raw head (decode_filter); else from decoded predictions.
  const int key_stride = e->cand_cap;
  GT_CHECK(e, A <= e->A, "nms: A=%d exceeds engine anchors %d", A, e->A);
  GT_CHECK(e, max_det <= e->cfg.max_det, "nms: max_det %d exceeds configured %d", max_det, e->cfg.max_det);
  GT_CUDA(e, cudaMemsetAsync(e->cand_count, 0, sizeof(int) * B, st));
  const long long total = (long long)B * A;
  const ClassMask cm = make_class_mask(e, classes_mask);
  if (pred_dev) {
    pred_filter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pred_dev, B, A, nc, rotated, conf, cm, e->cand_box,
                                                                        e->cand_conf, e->cand_cls, e->cand_key, key_stride, e->cand_count);
  } else {
    DecodeGeom g;
    for (int i = 0; i < 3; ++i) { g.lvl_w[i] = e->lvl_w[i]; g.lvl_h[i] = e->lvl_h[i]; g.lvl_off[i] = e->lvl_off[i]; g.stride[i] = (float)(8 << i); }
    decode_filter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e->raw_box, e->raw_cls, e->raw_ang, B, A, e->ncp, nc, rotated, g, conf, cm,
                                                                          e->cand_box, e->cand_conf, e->cand_cls, e->cand_key, key_stride,
                                                                          e->cand_count, e->nonfinite_dev);
    GT_CUDA(e, cudaMemcpyAsync(e->nonfinite_host, e->nonfinite_dev, sizeof(int), cudaMemcpyDeviceToHost, st));   // cumulative counter -> pinned mirror
  }
  e->launches++;
  const int max_nms = e->cfg.max_nms;
  int* overflow = e->det_keep + (size_t)e->cfg.max_batch * e->cfg.max_det;  // [max_batch] tail of det_keep
  const float fw0 = (float)e->cfg.frame_w, fh0 = (float)e->cfg.frame_h;
  std::vector<int> ov(B);
  if (e->nms_fused && max_det <= kNmsSmall) {
    // the common case: every frame has at most kNmsSmall candidates -> one launch does sort + mask + sweep per frame
    GT_CUDA(e, cudaMemsetAsync(overflow, 0, sizeof(int) * B, st));
    nms_small_kernel<<<B, 1024, kNmsSmallSmem, st>>>(e->cand_key, key_stride, e->cand_count, max_nms, A, e->cand_box, e->cand_conf, e->cand_cls, agnostic, rotated,
                                                     iou, max_det, scale_to_frame ? 1 : 0, (float)e->pad_left, (float)e->pad_top, e->gain, fw0, fh0, e->det_out,
                                                     e->det_count, e->det_keep, overflow);
    e->launches++;
    GT_CUDA(e, cudaGetLastError());
    // the stage's GPU work ends here: the postprocess stage time must not include the host's wait for the flags below
    GT_CUDA(e, cudaEventRecord(e->ev[4], st));
    GT_CUDA(e, cudaMemcpyAsync(ov.data(), overflow, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    GT_CUDA(e, cudaStreamSynchronize(st));
    bool any = false;
    for (int b = 0; b < B; ++b) any |= ov[b] != 0;
    if (!any) { e->post_event_done = true; return GT_OK; }
  }
  e->post_event_done = false;
  sort_keys_kernel<<<B, 1024, 4096 * sizeof(unsigned long long), st>>>(e->cand_key, key_stride, e->cand_count);
  e->launches++;
  // pass 1: batched, capacity nms_cap per image
  GT_CUDA(e, cudaMemsetAsync(overflow, 0, sizeof(int) * B, st));
  auto run_pass = [&](int b0, int nb, int n_cap) -> int {
    const int words = (n_cap + 63) / 64;
    float* sb = reinterpret_cast<float*>(e->nms_mask + (size_t)nb * n_cap * words);  // sorted boxes live after the mask
    dim3 gg((unsigned)((n_cap + 255) / 256), (unsigned)nb);
    nms_gather_kernel<<<gg, 256, 0, st>>>(e->cand_key + (size_t)b0 * key_stride, key_stride, e->cand_count + b0, A, max_nms,
                                          e->cand_box + (size_t)b0 * A * 5, e->cand_cls + (size_t)b0 * A, agnostic, rotated, sb, n_cap);
    dim3 gm((unsigned)words, (unsigned)words, (unsigned)nb);
    nms_mask_kernel<<<gm, 64, 0, st>>>(sb, e->cand_count + b0, max_nms, n_cap, words, iou, e->nms_mask, rotated);
    const float fw = (float)e->cfg.frame_w, fh = (float)e->cfg.frame_h;
    nms_sweep_kernel<<<nb, 512, words * sizeof(unsigned long long), st>>>(
        e->nms_mask, sb, e->cand_key + (size_t)b0 * key_stride, key_stride, e->cand_count + b0, max_nms, n_cap, words, A,
        e->cand_box + (size_t)b0 * A * 5, e->cand_conf + (size_t)b0 * A, e->cand_cls + (size_t)b0 * A, rotated, max_det,
        scale_to_frame ? 1 : 0, (float)e->pad_left, (float)e->pad_top, e->gain, fw, fh,
        e->det_out + (size_t)b0 * max_det * (rotated ? 7 : 6), e->det_count + b0, e->det_keep + (size_t)b0 * max_det, overflow + b0);
    e->launches += 3;
    GT_CUDA(e, cudaGetLastError());
    return GT_OK;
  };
  GT_TRY(run_pass(0, B, e->nms_cap));
  // pass 2 (rare): images with more than nms_cap candidates, one at a time with the full max_nms capacity
  GT_CUDA(e, cudaMemcpyAsync(ov.data(), overflow, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
  GT_CUDA(e, cudaStreamSynchronize(st));
  for (int b = 0; b < B; ++b)
    if (ov[b]) GT_TRY(run_pass(b, 1, ((max_nms + 63) / 64) * 64));
  return GT_OK;
}

// =====================================================================================================================
// Conv kernel variant per layer: shipped table keyed by the layer signature, else a fixed rule (see plan_both)
// =====================================================================================================================
ConvSig conv_signature(const ConvPlanArgs& a) {
  ConvSig s;
  const int pad = a.pad >= 0 ? a.pad : a.k / 2;
  s.cin = a.cin; s.cout = a.cout; s.k = a.k; s.stride = a.stride;
  s.H = a.Ho > 0 ? a.Ho : (a.in.H + 2 * pad - a.k) / a.stride + 1;
  s.W = a.Wo > 0 ? a.Wo : (a.in.W + 2 * pad - a.k) / a.stride + 1;
  s.flags = (a.res ? 1 : 0) | (a.up ? 2 : 0) | (a.out_f32 ? 4 : 0) | (a.out_s2d ? 8 : 0) | (a.pre ? 16 : 0) | (a.chain_cout ? 32 : 0);
  return s;
}

struct TuneRow { int cin, cout, k, stride, H, W, flags, variant; };
static const TuneRow kTuneTable[] = {
#include "conv_tune.inc"
};

// Layers the table does not know (other frame sizes, other heads): the pattern of the tuned table, as a rule.
static int rule_variant(const ConvSig& s) {
  if (s.flags & (4 | 8)) return 3;                                     // f32 head rows, layer 0: pixel-major, two CTAs / SM
  if (s.flags & 32) return s.cout <= 64 ? 3 : 0;                       // chained 1x1 conv: pixel-major epilogue only
  const long long px = (long long)s.H * s.W;
  if (s.k == 1) {
    if (s.flags & 16) return s.cout <= 128 ? 3 : 0;                    // half-resolution pre-activation add: pixel-major epilogue only
    if (s.flags & 2) return 0;                                         // upsampled copy: pixel-major slabs
    if ((s.cout % 256) == 0) return s.cin >= 384 ? 6 : 1;
    return s.cin >= 384 ? 0 : 3;
  }
  if (s.k == 2) return 3;
  if (s.stride == 2) {
    if ((s.cout % 256) == 0) return 6;
    if (s.cout <= 64) return 3;
    return s.cin <= 64 ? 1 : 0;
  }
  if ((s.cout % 256) == 0) return 6;
  if (s.cout > 128) return 0;
  if (s.cout == 128) return px >= 136 * 240 ? 2 : (px >= 68 * 120 ? 1 : 0);
  if (s.cout > 32) return (s.flags & 1) ? 5 : 4;
  return 5;
}

int conv_choose_variant(const ConvSig& s) {
  for (const TuneRow& r : kTuneTable)
    if (r.cin == s.cin && r.cout == s.cout && r.k == s.k && r.stride == s.stride && r.H == s.H && r.W == s.W && r.flags == s.flags) return r.variant;
  return rule_variant(s);
}

// =====================================================================================================================
// Network plan
// =====================================================================================================================
namespace {

struct Builder {
  gt_engine* e;
  int B;
  std::map<std::string, int> idx;
  int rc = GT_OK;

  View alloc(int C, int H, int W) {
    View v;
    bf16* p = nullptr;
    if (e->dev_alloc((void**)&p, (size_t)B * H * W * C * sizeof(bf16)) != GT_OK) rc = GT_ERR_NOMEM;
    v.ptr = p; v.C = C; v.ctot = C; v.coff = 0; v.H = H; v.W = W;
    return v;
  }
  int find(const std::string& name) {
    auto it = idx.find(name);
    if (it == idx.end()) { gt_set_error(e, "plan: unknown conv %s", name.c_str()); rc = GT_ERR_INVALID; return 0; }
    return it->second;
  }
  void push(const ConvOp& op) {
    e->conv_ops.push_back(op);
    PlanOp po; po.type = OP_CONV; po.conv = (int)e->conv_ops.size() - 1;
    e->plan.push_back(po);
  }
  // plans the op in the forced variant, or in every variant when the engine autotunes (conv_ops gets variant 0, conv_var[v] the others;
  // a variant that does not apply to the layer -- halo on a 1x1, two CTAs with BN > 128 -- is marked invalid and never timed)
  static bool variant_applies(int v, const ConvOp& o) {
    switch (v) {
      case 1: return o.swapped && !o.p.halo;
      case 2: return o.swapped && o.p.halo;
      case 3: return !o.swapped && o.occ2 && !o.p.halo;
      case 4: return !o.swapped && !o.occ2 && o.p.halo;
      case 5: return !o.swapped && o.occ2 && o.p.halo;
      case 6: return o.swapped && o.pair;
      default: return true;
    }
  }
  int plan_both(ConvOp& op, const ConvPlanArgs& a_in) {
    ConvPlanArgs a = a_in;
    // MUFU-bound layers (>= silu_tanh_px output pixels per image; layer 0 counts its 2 x 2 pixels per GEMM row) use the one-MUFU SiLU
    if (a.act == 1 && e->silu_tanh_px > 0) {
      const ConvSig sg = conv_signature(a);
      if ((long long)sg.H * sg.W * (a.out_s2d ? 4 : 1) >= e->silu_tanh_px) a.act = 2;
    }
    if (e->swap_mode >= 0) {               // GT_SWAP=v: one variant forced wherever it applies (tests)
      e->plan_variant = e->swap_mode;
      return conv_tc_plan(e, &op, a);
    }
    if (!e->tune_mode) {
      // Default: the variant is a fixed function of the layer's signature (shipped table, else a rule) -- never of a timing -- so every
      // process / rank / engine runs the same kernels in the same accumulation order and outputs are bit-identical across them.
      e->plan_variant = conv_choose_variant(conv_signature(a));
      const int r = conv_tc_plan(e, &op, a);   // (a variant that does not apply to the layer plans as its plain base kernel)
      e->plan_variant = 0;
      return r;
    }
    // GT_TUNE=1 (development): plan every variant; detector_autotune times them at weight load and prints the table rows
    e->plan_variant = 0;
    int r = conv_tc_plan(e, &op, a);
    if (r != GT_OK) return r;
    for (int v = 1; v < GT_CONV_VARIANTS; ++v) {
      ConvOp alt = op;
      if (v == 6 && !e->pair_mode) {   // GT_PAIR=0: not planned (no weights, never timed)
        e->conv_var[v].push_back(alt);
        e->conv_var_ok[v].push_back(0);
        continue;
      }
      e->plan_variant = v;
      r = conv_tc_plan(e, &alt, a);
      e->plan_variant = 0;
      if (r != GT_OK) return r;
      e->conv_var[v].push_back(alt);
      e->conv_var_ok[v].push_back(variant_applies(v, alt) ? 1 : 0);
    }
    e->conv_sig.push_back(conv_signature(a));
    return r;
  }
  // 16-bit conv writing a channel slice; several canonical convs reading the same input are fused along cout
  void conv(std::vector<std::string> names, const View& in, const View& out, const View* res = nullptr, const View* up = nullptr) {
    if (rc != GT_OK) return;
    ConvOp op;
    int cout = 0;
    op.n_src = (int)names.size();
    for (int i = 0; i < op.n_src; ++i) { op.src[i] = find(names[i]); cout += e->conv_descs[op.src[i]].cout; }
    const gt_conv_desc& d = e->conv_descs[op.src[0]];
    ConvPlanArgs a;
    a.in = in; a.Bmax = B; a.cin = d.cin; a.cout = cout; a.k = d.k; a.stride = d.stride; a.act = d.act; a.out = &out; a.res = res; a.up = up;
    rc = plan_both(op, a);
    if (rc == GT_OK) push(op);
  }
  // conv `name` followed, inside the same kernel, by the 1x1 conv `name2` (ConvParams::chain_n): only name2's output reaches HBM
  void conv_chain(const std::string& name, const std::string& name2, const View& in, const View& out) {
    if (rc != GT_OK) return;
    ConvOp op;
    op.n_src = 1; op.src[0] = find(name); op.chain_src = find(name2);
    const gt_conv_desc& d = e->conv_descs[op.src[0]];
    const gt_conv_desc& d2 = e->conv_descs[op.chain_src];
    if (d2.k != 1 || d2.stride != 1 || d2.cin != d.cout) { gt_set_error(e, "plan: %s cannot be chained after %s", name2.c_str(), name.c_str()); rc = GT_ERR_INVALID; return; }
    ConvPlanArgs a;
    a.in = in; a.Bmax = B; a.cin = d.cin; a.cout = d.cout; a.k = d.k; a.stride = d.stride; a.act = d.act; a.out = &out;
    a.chain_cout = d2.cout; a.chain_act = d2.act;
    if (a.chain_act == 1 && e->silu_tanh_px > 0 && (long long)out.H * out.W >= e->silu_tanh_px) a.chain_act = 2;
    rc = plan_both(op, a);
    if (rc == GT_OK) push(op);
  }
  // final head conv writing fp32 rows of one of the three raw-head arrays (box logits [A][64], class logits [A][ncp], angle [A][4])
  void conv_raw(const std::string& name, const View& in, int lvl_off, float* base, int ctot) {
    if (rc != GT_OK) return;
    ConvOp op;
    op.n_src = 1; op.src[0] = find(name);
    const gt_conv_desc& d = e->conv_descs[op.src[0]];
    ConvPlanArgs a;
    a.in = in; a.Bmax = B; a.cin = d.cin; a.cout = d.cout; a.k = d.k; a.stride = d.stride; a.act = d.act;
    a.out_f32 = base + (size_t)lvl_off * ctot; a.out_img_stride = e->A; a.out_ctot_f32 = ctot; a.out_coff_f32 = 0;
    rc = plan_both(op, a);
    if (rc == GT_OK) push(op);
  }
  // layer 0: Conv(3 -> 32, k3, s2) as a stride-1 2x2 convolution over the 64-channel 4x4 space-to-depth input producing a 2x2
  // block of output pixels x 32 channels per GEMM row (see stage 1)
  void conv0(const View& s2d, const View& out) {
    if (rc != GT_OK) return;
    ConvOp op;
    op.n_src = 1; op.src[0] = find("model.0");
    ConvPlanArgs a;
    a.in = s2d; a.Bmax = B; a.cin = 64; a.cout = 128; a.k = 2; a.stride = 1; a.pad = 1; a.Ho = s2d.H; a.Wo = s2d.W;
    a.act = 1; a.scale = 1.0f / 255.0f; a.out = &out; a.out_s2d = true;
    rc = plan_both(op, a);
    if (rc != GT_OK) return;
    op.flops = 2.0 * out.H * out.W * 32.0 * 27.0;   // the algorithmic 3x3x3 work, not the zero-padded 2x2x64 -> 128
    for (int v = 1; v < GT_CONV_VARIANTS; ++v) if (!e->conv_var[v].empty()) e->conv_var[v].back().flops = op.flops;
    e->conv0_op = (int)e->conv_ops.size();
    push(op);
  }
  // 1x1 conv over cat(up2x(lo), hi) WITHOUT materialising the upsampled tensor: a 1x1 convolution commutes with nearest upsampling,
  // so the `lo` branch (the first lo.C input channels of the canonical conv) runs at lo's resolution into a half-resolution
  // pre-activation partial sum Z (no bias, no activation: 4x fewer MACs, and neither the 4 upsampled stores of the producer nor their
  // read-back exist any more), and the `hi` branch adds Z[y/2][x/2] + bias before its activation (ConvParams::res_pre).
  void conv_cat_up(const std::string& name, const View& lo, const View& hi, const View& out) {
    if (rc != GT_OK) return;
    const int ci = find(name);
    const gt_conv_desc& d = e->conv_descs[ci];
    if (d.k != 1 || d.stride != 1 || d.cin != lo.C + hi.C || hi.H != 2 * lo.H || hi.W != 2 * lo.W) { gt_set_error(e, "plan: %s is not a 1x1 conv over cat(up2x, .)", name.c_str()); rc = GT_ERR_INVALID; return; }
    View Z = alloc(d.cout, lo.H, lo.W);
    ConvOp part;
    part.n_src = 1; part.src[0] = ci; part.w_cin_total = d.cin; part.w_cin_off = 0; part.no_bias = 1;
    ConvPlanArgs a;
    a.in = lo; a.Bmax = B; a.cin = lo.C; a.cout = d.cout; a.k = 1; a.stride = 1; a.act = 0; a.out = &Z;
    rc = plan_both(part, a);
    if (rc != GT_OK) return;
    part.flops = 0;                                            // the algorithmic FLOPs of the canonical conv are booked on the final op
    for (int v = 1; v < GT_CONV_VARIANTS; ++v) if (!e->conv_var[v].empty()) e->conv_var[v].back().flops = 0;
    push(part);
    ConvOp fin;
    fin.n_src = 1; fin.src[0] = ci; fin.w_cin_total = d.cin; fin.w_cin_off = lo.C;
    ConvPlanArgs b;
    b.in = hi; b.Bmax = B; b.cin = hi.C; b.cout = d.cout; b.k = 1; b.stride = 1; b.act = d.act; b.out = &out; b.pre = &Z;
    rc = plan_both(fin, b);
    if (rc != GT_OK) return;
    fin.flops = 2.0 * hi.H * hi.W * (double)d.cout * d.cin;   // what ultralytics' conv over the materialised concat costs
    for (int v = 1; v < GT_CONV_VARIANTS; ++v) if (!e->conv_var[v].empty()) e->conv_var[v].back().flops = fin.flops;
    push(fin);
  }
  // chain_from / chain_in: the conv that produces the block's input runs in the same kernel as the block's cv1 (conv_chain); `in` then
  // only gives the geometry (its buffer is never written)
  void c2f(const std::string& pre, const View& in, int c2, int n, bool shortcut, const View& out, const View* up = nullptr, const View* in_lo = nullptr,
           const char* chain_from = nullptr, const View* chain_in = nullptr) {
    const int c = c2 / 2;
    View cat = alloc((2 + n) * c, in.H, in.W);
    View tmp = alloc(c, in.H, in.W);
    if (chain_from) conv_chain(chain_from, pre + ".cv1", *chain_in, cat.slice(0, 2 * c));
    else if (in_lo) conv_cat_up(pre + ".cv1", *in_lo, in, cat.slice(0, 2 * c));   // the block's input is cat(up2x(in_lo), in)
    else conv({pre + ".cv1"}, in, cat.slice(0, 2 * c));
    for (int i = 0; i < n; ++i) {
      View src = cat.slice(c * (1 + i), c);
      conv({pre + ".m." + std::to_string(i) + ".cv1"}, src, tmp);
      View dst = cat.slice(c * (2 + i), c);
      conv({pre + ".m." + std::to_string(i) + ".cv2"}, tmp, dst, shortcut ? &src : nullptr);
    }
    conv({pre + ".cv2"}, cat, out, nullptr, up);
  }
};

void add_desc(gt_engine* e, const std::string& name, int cin, int cout, int k, int s, int act) {
  gt_conv_desc d;
  memset(&d, 0, sizeof(d));
  snprintf(d.name, sizeof(d.name), "%s", name.c_str());
  d.cin = cin; d.cout = cout; d.k = k; d.stride = s; d.act = act;
  e->conv_descs.push_back(d);
}
void add_c2f_desc(gt_engine* e, const std::string& pre, int c1, int c2, int n) {
  const int c = c2 / 2;
  add_desc(e, pre + ".cv1", c1, 2 * c, 1, 1, 1);
  add_desc(e, pre + ".cv2", (2 + n) * c, c2, 1, 1, 1);
  for (int i = 0; i < n; ++i) {
    add_desc(e, pre + ".m." + std::to_string(i) + ".cv1", c, c, 3, 1, 1);
    add_desc(e, pre + ".m." + std::to_string(i) + ".cv2", c, c, 3, 1, 1);
  }
}

}  // namespace

int detector_build(gt_engine* e) {
  GT_CUDA(e, cudaFuncSetAttribute(nms_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmallSmem));
  const int B = e->cfg.max_batch, nc = e->cfg.nc;
  const bool obb = e->cfg.task == GT_TASK_OBB;
  const int c1 = 32, c2 = 64, c3 = 128, c4 = 256, c5 = 512;  // YOLOv8 s-scale widths
  // canonical conv list (ultralytics module order)
  add_desc(e, "model.0", 3, c1, 3, 2, 1);
  add_desc(e, "model.1", c1, c2, 3, 2, 1);
  add_c2f_desc(e, "model.2", c2, c2, 1);
  add_desc(e, "model.3", c2, c3, 3, 2, 1);
  add_c2f_desc(e, "model.4", c3, c3, 2);
  add_desc(e, "model.5", c3, c4, 3, 2, 1);
  add_c2f_desc(e, "model.6", c4, c4, 2);
  add_desc(e, "model.7", c4, c5, 3, 2, 1);
  add_c2f_desc(e, "model.8", c5, c5, 1);
  add_desc(e, "model.9.cv1", c5, c5 / 2, 1, 1, 1);
  add_desc(e, "model.9.cv2", c5 * 2, c5, 1, 1, 1);
  add_c2f_desc(e, "model.12", c5 + c4, c4, 1);
  add_c2f_desc(e, "model.15", c4 + c3, c3, 1);
  add_desc(e, "model.16", c3, c3, 3, 2, 1);
  add_c2f_desc(e, "model.18", c3 + c4, c4, 1);
  add_desc(e, "model.19", c4, c4, 3, 2, 1);
  add_c2f_desc(e, "model.21", c4 + c5, c5, 1);
  const int ch[3] = {c3, c4, c5};
  const int hc2 = 64, hc3 = std::max(c3, std::min(nc, 100)), hc4 = std::max(c3 / 4, 1);
  for (int i = 0; i < 3; ++i) {
    const std::string p = "model.22.cv2." + std::to_string(i);
    add_desc(e, p + ".0", ch[i], hc2, 3, 1, 1);
    add_desc(e, p + ".1", hc2, hc2, 3, 1, 1);
    add_desc(e, p + ".2", hc2, 64, 1, 1, 0);
  }
  for (int i = 0; i < 3; ++i) {
    const std::string p = "model.22.cv3." + std::to_string(i);
    add_desc(e, p + ".0", ch[i], hc3, 3, 1, 1);
    add_desc(e, p + ".1", hc3, hc3, 3, 1, 1);
    add_desc(e, p + ".2", hc3, nc, 1, 1, 0);
  }
  if (obb)
    for (int i = 0; i < 3; ++i) {
      const std::string p = "model.22.cv4." + std::to_string(i);
      add_desc(e, p + ".0", ch[i], hc4, 3, 1, 1);
      add_desc(e, p + ".1", hc4, hc4, 3, 1, 1);
      add_desc(e, p + ".2", hc4, 1, 1, 1, 0);
    }

  Builder bl;
  bl.e = e; bl.B = B;
  for (size_t i = 0; i < e->conv_descs.size(); ++i) bl.idx[e->conv_descs[i].name] = (int)i;

  const int H = e->net_h, W = e->net_w;
  const int H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4, H3 = H / 8, W3 = W / 8, H4 = H / 16, W4 = W / 16, H5 = H / 32, W5 = W / 32;
  e->lvl_h[0] = H3; e->lvl_w[0] = W3; e->lvl_h[1] = H4; e->lvl_w[1] = W4; e->lvl_h[2] = H5; e->lvl_w[2] = W5;
  e->lvl_off[0] = 0; e->lvl_off[1] = H3 * W3; e->lvl_off[2] = H3 * W3 + H4 * W4;
  e->A = H3 * W3 + H4 * W4 + H5 * W5;
  e->no = 64 + nc + (obb ? 1 : 0);
  // The raw head lives in three dense fp32 arrays -- box logits [B][A][64] (256-byte rows), class logits [B][A][ncp] (ncp = nc rounded up
  // to 4: 16-byte rows for the 4-class head), OBB angle [B][A][4] -- so that every row and every slab chunk the final 1x1 convs store
  // starts on a 32-byte DRAM sector and a tile's class rows are contiguous.  (One interleaved [A][68] array put odd rows 16 bytes off
  // a sector: the 64->64 box conv took 63 us instead of 37, and decode read a 32-byte sector for every 16 bytes of class logits.)
  e->ncp = (nc + 3) / 4 * 4;
  const size_t nA = (size_t)B * e->A;
  GT_TRY(e->dev_alloc((void**)&e->raw_box, nA * 64 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->raw_cls, nA * e->ncp * sizeof(float)));
  GT_CUDA(e, cudaMemset(e->raw_box, 0, nA * 64 * sizeof(float)));
  GT_CUDA(e, cudaMemset(e->raw_cls, 0, nA * e->ncp * sizeof(float)));
  if (obb) {
    GT_TRY(e->dev_alloc((void**)&e->raw_ang, nA * 4 * sizeof(float)));
    GT_CUDA(e, cudaMemset(e->raw_ang, 0, nA * 4 * sizeof(float)));
  }

  View S2D = bl.alloc(64, H2, W2);          // 4x4 space-to-depth network input, written by stage 1
  e->net_s2d = S2D.ptr;
  View T0 = bl.alloc(c1, H1, W1);
  bl.conv0(S2D, T0);
  // model.1 (3x3 s2, 32 -> 64) + model.2.cv1 (1x1, 64 -> 64) as ONE kernel: model.1's output (16.7 MB per 4K frame) never exists in HBM
  const bool chain1 = e->chain_mode != 0;
  View T1;
  if (chain1) { T1.ptr = nullptr; T1.C = c2; T1.ctot = c2; T1.coff = 0; T1.H = H2; T1.W = W2; }
  else { T1 = bl.alloc(c2, H2, W2); bl.conv({"model.1"}, T0, T1); }
  View T2 = bl.alloc(c2, H2, W2);
  if (chain1) bl.c2f("model.2", T1, c2, 1, true, T2, nullptr, nullptr, "model.1", &T0);
  else bl.c2f("model.2", T1, c2, 1, true, T2);
  View T3 = bl.alloc(c3, H3, W3);
  bl.conv({"model.3"}, T2, T3);
  View L4 = bl.alloc(c3, H3, W3);
  bl.c2f("model.4", T3, c3, 2, true, L4);
  View T5 = bl.alloc(c4, H4, W4);
  bl.conv({"model.5"}, L4, T5);
  View L6 = bl.alloc(c4, H4, W4);
  bl.c2f("model.6", T5, c4, 2, true, L6);
  View T7 = bl.alloc(c5, H5, W5);
  bl.conv({"model.7"}, L6, T7);
  View T8 = bl.alloc(c5, H5, W5);
  bl.c2f("model.8", T7, c5, 1, true, T8);
  // SPPF
  View s9 = bl.alloc(c5 * 2, H5, W5);
  bl.conv({"model.9.cv1"}, T8, s9.slice(0, c5 / 2));
  for (int i = 0; i < 3; ++i) {
    PlanOp po; po.type = OP_MAXPOOL;
    po.pool.in = s9.slice(i * (c5 / 2), c5 / 2);
    po.pool.out = s9.slice((i + 1) * (c5 / 2), c5 / 2);
    e->plan.push_back(po);
  }
  View cat20 = bl.alloc(c4 + c5, H5, W5);   // [19 | 9]
  View L9 = cat20.slice(c4, c5);
  bl.conv({"model.9.cv2"}, s9, L9);
  View cat17 = bl.alloc(c3 + c4, H4, W4);   // [16 | 12]
  View L12 = cat17.slice(c3, c4);
  // Upsample + Concat (model.10/11, model.13/14) are never materialised: model.12.cv1 / model.15.cv1 are 1x1 convs over
  // cat(up2x(lo), hi) and run as conv_cat_up (lo branch at lo's resolution, added before the activation of the hi branch)
  bl.c2f("model.12", L6, c4, 1, false, L12, nullptr, &L9);
  View P3 = bl.alloc(c3, H3, W3);
  bl.c2f("model.15", L4, c3, 1, false, P3, nullptr, &L12);
  bl.conv({"model.16"}, P3, cat17.slice(0, c3));
  View P4 = bl.alloc(c4, H4, W4);
  bl.c2f("model.18", cat17, c4, 1, false, P4);
  bl.conv({"model.19"}, P4, cat20.slice(0, c4));
  View P5 = bl.alloc(c5, H5, W5);
  bl.c2f("model.21", cat20, c5, 1, false, P5);
  // head
  const View feats[3] = {P3, P4, P5};
  for (int i = 0; i < 3; ++i) {
    const std::string s = std::to_string(i);
    const int ha = hc2 + hc3 + (obb ? hc4 : 0);
    View a = bl.alloc(ha, feats[i].H, feats[i].W);
    std::vector<std::string> first = {"model.22.cv2." + s + ".0", "model.22.cv3." + s + ".0"};
    if (obb) first.push_back("model.22.cv4." + s + ".0");
    bl.conv(first, feats[i], a);
    View b2 = bl.alloc(hc2, feats[i].H, feats[i].W), b3 = bl.alloc(hc3, feats[i].H, feats[i].W);
    bl.conv({"model.22.cv2." + s + ".1"}, a.slice(0, hc2), b2);
    bl.conv({"model.22.cv3." + s + ".1"}, a.slice(hc2, hc3), b3);
    bl.conv_raw("model.22.cv2." + s + ".2", b2, e->lvl_off[i], e->raw_box, 64);
    bl.conv_raw("model.22.cv3." + s + ".2", b3, e->lvl_off[i], e->raw_cls, e->ncp);
    if (obb) {
      View b4 = bl.alloc(hc4, feats[i].H, feats[i].W);
      bl.conv({"model.22.cv4." + s + ".1"}, a.slice(hc2 + hc3, hc4), b4);
      bl.conv_raw("model.22.cv4." + s + ".2", b4, e->lvl_off[i], e->raw_ang, 4);
    }
  }
  if (bl.rc != GT_OK) return bl.rc;
  e->feat_views[0] = T0; e->feat_views[1] = T1; e->feat_views[2] = T2; e->feat_views[3] = T3; e->feat_views[4] = L4;
  e->feat_views[5] = T5; e->feat_views[6] = L6; e->feat_views[7] = T7; e->feat_views[8] = T8; e->feat_views[9] = L9;
  e->feat_views[12] = L12; e->feat_views[15] = P3; e->feat_views[16] = cat17.slice(0, c3); e->feat_views[18] = P4;
  e->feat_views[19] = cat20.slice(0, c4); e->feat_views[21] = P5;

  e->conv_flops = 0;
  for (const ConvOp& op : e->conv_ops) e->conv_flops += op.flops;

  // decode / NMS workspaces
  int cap = 1;
  while (cap < e->A) cap <<= 1;
  e->cand_cap = cap;
  GT_TRY(e->dev_alloc((void**)&e->cand_box, (size_t)B * e->A * 5 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->cand_conf, (size_t)B * e->A * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->cand_cls, (size_t)B * e->A * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->cand_key, (size_t)B * cap * sizeof(unsigned long long)));
  GT_TRY(e->dev_alloc((void**)&e->cand_count, (size_t)B * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->nonfinite_dev, sizeof(int)));
  GT_CUDA(e, cudaMemset(e->nonfinite_dev, 0, sizeof(int)));
  GT_TRY(e->host_alloc((void**)&e->nonfinite_host, sizeof(int)));
  *e->nonfinite_host = 0;
  e->nms_cap = 4096;
  const size_t big_cap = (size_t)((e->cfg.max_nms + 63) / 64) * 64;
  const size_t batched = (size_t)B * e->nms_cap * (e->nms_cap / 64) * 8 + (size_t)B * e->nms_cap * 5 * 4;
  const size_t single = big_cap * (big_cap / 64) * 8 + big_cap * 5 * 4;
  GT_TRY(e->dev_alloc((void**)&e->nms_mask, std::max(batched, single)));
  GT_TRY(e->dev_alloc((void**)&e->det_out, (size_t)B * e->cfg.max_det * 7 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->det_count, (size_t)B * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->det_keep, ((size_t)B * e->cfg.max_det + B) * sizeof(int)));
  return GT_OK;
}

static int load_op_weights(gt_engine* e, ConvOp& op, bool is_conv0, const float* const* w, const float* const* b) {
  const int fp16 = e->cfg.act_dtype == GT_ACT_FP16;
  if (is_conv0) {
    // layer 0: [32][3][3][3] (cout, RGB, ky, kx) -> [n = (oy*2+ox)*32 + co][tap = dY*2+dX][c64 = (r*4+c)*4 + ch].  GEMM row =
    // super-pixel (Y, X); output pixel (2Y+oy, 2X+ox) reads input row 4(Y-1+dY) + r as kernel row ky = 4 dY + r - 2 oy - 3 (columns
    // likewise); combinations outside 0..2 are zero.  Unscaled: the 1/255 is the epilogue scale.
    const int taps = 4, cpad = op.cin_pad;   // 64
    std::vector<uint16_t> hw((size_t)op.cout_pad * taps * cpad, 0);
    std::vector<float> hb(op.cout_pad, 0.f);
    for (int oy = 0; oy < 2; ++oy)
      for (int ox = 0; ox < 2; ++ox)
        for (int co = 0; co < 32; ++co) {
          const int n = (oy * 2 + ox) * 32 + co;
          for (int dY = 0; dY < 2; ++dY)
            for (int dX = 0; dX < 2; ++dX)
              for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                  const int ky = 4 * dY + r - 2 * oy - 3, kx = 4 * dX + c - 2 * ox - 3;
                  if (ky < 0 || ky > 2 || kx < 0 || kx > 2) continue;
                  for (int chn = 0; chn < 3; ++chn)
                    hw[((size_t)n * taps + dY * 2 + dX) * cpad + (r * 4 + c) * 4 + chn] = host_to_act(w[0][((co * 3 + chn) * 3 + ky) * 3 + kx], fp16);
                }
          hb[n] = b[0][co];
        }
    return conv_tc_upload_packed(e, &op, hw.data(), hb.data());
  }
  const float* ws[3];
  const float* bs[3];
  int couts[3];
  for (int i = 0; i < op.n_src; ++i) { ws[i] = w[op.src[i]]; bs[i] = b[op.src[i]]; couts[i] = e->conv_descs[op.src[i]].cout; }
  GT_TRY(conv_tc_pack_weights(e, &op, ws, bs, couts, op.n_src));
  if (op.chain_src >= 0) GT_TRY(conv_tc_pack_chain(e, &op, w[op.chain_src], b[op.chain_src]));
  return GT_OK;
}

int detector_load_weights(gt_engine* e, const float* const* w, const float* const* b, int n) {
  GT_CHECK(e, n == (int)e->conv_descs.size(), "load_weights: expected %d convs, got %d", (int)e->conv_descs.size(), n);
  for (size_t oi = 0; oi < e->conv_ops.size(); ++oi) {
    GT_TRY(load_op_weights(e, e->conv_ops[oi], (int)oi == e->conv0_op, w, b));
    for (int v = 1; v < GT_CONV_VARIANTS; ++v)
      if (!e->conv_var[v].empty() && e->conv_var_ok[v][oi]) GT_TRY(load_op_weights(e, e->conv_var[v][oi], (int)oi == e->conv0_op, w, b));
  }
  e->weights_loaded = true;
  return GT_OK;
}

// GT_TUNE=1 (development only): times every applicable variant of every conv on the full batch, runs the fastest, and prints one
// `[gt tune]` line per layer; tools/make_tune_table.py turns that log into csrc/conv_tune.inc, the table that SHIPS.  A normal engine
// never times anything: its per-layer variant comes from that table (plan_both), so it is the same in every process.
static void apply_choice(gt_engine* e, size_t i, int v) {
  if (v > 0) std::swap(e->conv_ops[i], e->conv_var[v][i]);
}

int detector_autotune(gt_engine* e, cudaStream_t st) {
  if (!e->tune_mode || e->conv_var[1].empty() || e->tuned) return GT_OK;
  const int B = e->cfg.max_batch;
  cudaEvent_t a, b;
  GT_CUDA(e, cudaEventCreate(&a));
  GT_CUDA(e, cudaEventCreate(&b));
  int64_t tune_launches = 0;
  constexpr int kReps = 5;
  auto time_op = [&](const ConvOp& op, float* us) -> int {
    GT_TRY(conv_tc_launch(e, &op, B, st));   // warm-up
    GT_CUDA(e, cudaEventRecord(a, st));
    for (int i = 0; i < kReps; ++i) GT_TRY(conv_tc_launch(e, &op, B, st));
    GT_CUDA(e, cudaEventRecord(b, st));
    GT_CUDA(e, cudaEventSynchronize(b));
    float ms = 0.f;
    GT_CUDA(e, cudaEventElapsedTime(&ms, a, b));
    *us = ms * 1000.f / kReps;
    tune_launches += kReps + 1;
    return GT_OK;
  };
  static const char* vname[GT_CONV_VARIANTS] = {"tc", "sw", "sw-halo", "tc2", "tc-halo", "tc2-halo", "sw-pair"};
  for (size_t i = 0; i < e->conv_ops.size(); ++i) {
    float t[GT_CONV_VARIANTS];
    int best = 0;
    GT_TRY(time_op(e->conv_ops[i], &t[0]));
    for (int v = 1; v < GT_CONV_VARIANTS; ++v) {
      t[v] = 0.f;
      if (!e->conv_var_ok[v][i]) continue;
      GT_TRY(time_op(e->conv_var[v][i], &t[v]));
      if (t[v] < t[best]) best = v;
    }
    const ConvOp& o = e->conv_ops[i];
    const ConvSig& sg = e->conv_sig[i];
    fprintf(stderr, "[gt tune] op %2zu src %2d cin %4d cout %4d k %d s %d out %4dx%-4d flags %d ", i, o.src[0], sg.cin, sg.cout, sg.k, sg.stride, sg.H, sg.W, sg.flags);
    for (int v = 0; v < GT_CONV_VARIANTS; ++v) fprintf(stderr, " %s %6.1f", vname[v], t[v]);
    fprintf(stderr, "  -> %s  %6.1f TFLOP/s  %6.1f GB/s\n", vname[best], o.flops * B / (t[best] * 1e-6) * 1e-12, o.bytes * B / (t[best] * 1e-6) * 1e-9);
    apply_choice(e, i, best);
  }
  e->launches -= tune_launches;   // tuning launches are not part of any step
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  e->tuned = true;
  return GT_OK;
}

static int launch_pool(gt_engine* e, const MaxpoolOp& po, int B, cudaStream_t st) {
  const View& i = po.in;
  const View& o = po.out;
  const long long total = (long long)B * i.H * i.W * (i.C / 8);
  if (e->cfg.act_dtype == GT_ACT_FP16)
    maxpool5_kernel<__half2><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(i.ptr, i.ctot, i.coff, o.ptr, o.ctot, o.coff, B, i.H, i.W, i.C);
  else
    maxpool5_kernel<__nv_bfloat162><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(i.ptr, i.ctot, i.coff, o.ptr, o.ctot, o.coff, B, i.H, i.W, i.C);
  e->launches++;
  return GT_OK;
}

// the SPPF pools: in = slice i, out = slice i + 1 of one buffer, three in a row -> one fused launch when the plane fits shared memory
static bool sppf_chain(const std::vector<PlanOp>& plan, size_t i) {
  if (i + 2 >= plan.size()) return false;
  for (int k = 0; k < 3; ++k) {
    const PlanOp& p = plan[i + k];
    if (p.type != OP_MAXPOOL || p.pool.in.ptr != plan[i].pool.in.ptr || p.pool.out.ptr != p.pool.in.ptr || p.pool.in.C != plan[i].pool.in.C ||
        p.pool.in.coff != plan[i].pool.in.coff + k * p.pool.in.C || p.pool.out.coff != p.pool.in.coff + p.pool.in.C)
      return false;
  }
  return (size_t)plan[i].pool.in.H * plan[i].pool.in.W * 32 <= 200 * 1024 && (plan[i].pool.in.C % 8) == 0;
}

int detector_forward(gt_engine* e, int B, cudaStream_t st) {
  GT_CHECK(e, e->weights_loaded, "detect: weights not loaded");
  for (size_t pi = 0; pi < e->plan.size(); ++pi) {
    const PlanOp& po = e->plan[pi];
    if (po.type == OP_CONV) GT_TRY(conv_tc_launch(e, &e->conv_ops[po.conv], B, st));
    else if (sppf_chain(e->plan, pi)) {
      const View& v = po.pool.in;
      const size_t smem = (size_t)v.H * v.W * 32;
      dim3 g((unsigned)(v.C / 8), (unsigned)B);
      if (e->cfg.act_dtype == GT_ACT_FP16) {
        GT_CUDA(e, cudaFuncSetAttribute(sppf_pool3_kernel<__half2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sppf_pool3_kernel<__half2><<<g, 1024, smem, st>>>(v.ptr, v.ctot, v.coff, v.C, v.H, v.W);
      } else {
        GT_CUDA(e, cudaFuncSetAttribute(sppf_pool3_kernel<__nv_bfloat162>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sppf_pool3_kernel<__nv_bfloat162><<<g, 1024, smem, st>>>(v.ptr, v.ctot, v.coff, v.C, v.H, v.W);
      }
      e->launches++;
      pi += 2;
    } else GT_TRY(launch_pool(e, po.pool, B, st));
  }
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

int detector_postprocess(gt_engine* e, int B, float conf, float iou, int agnostic, uint32_t classes_mask, cudaStream_t st) {
  return nms_run(e, nullptr, B, e->A, e->cfg.nc, e->cfg.task == GT_TASK_OBB, conf, iou, agnostic, classes_mask, e->cfg.max_det, true, st);
}
