// Stage 3a: ORB key points + rBRIEF descriptors on the half-resolution gray frame (batched over frames).
//
// Restates cv2.ORB_create(nfeatures).detectAndCompute(gray, mask) as stabilo calls it
// (/root/reference/geotrax/extract.py:177,181 -> stabilo.Stabilizer; OpenCV defaults scaleFactor 1.2, nlevels 8,
// edgeThreshold 31, HARRIS_SCORE, patchSize 31, fastThreshold 20 -- SURVEY.md 8a-10, Appendix A-3).  Every integer stage
// is bit-exact with OpenCV 4.13 (chained INTER_LINEAR_EXACT pyramid, FAST-9/16 score + 3x3 NMS, mask / border filter,
// "retain best" with ties kept); the float stages follow OpenCV's operation order (Harris 7x7, intensity-centroid angle
// with fastAtan2, float32 separable 7x7 sigma-2 blur, rotated 256-pair test pattern recovered by probing cv2).
#include <algorithm>
#include <cmath>

#include "engine.cuh"

namespace {

constexpr int kFastThr = 20;
constexpr int kEdge = 31;
constexpr int kSelCap = 8192;   // per (frame, level) candidates surviving the FAST-score cut
constexpr int kLvlKeep = 2048;  // per (frame, level) final key points


__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
__constant__ float c_gauss[7] = {0.07015932351350784f, 0.13107487559318542f, 0.1907128244638443f, 0.21610593795776367f,
                                 0.1907128244638443f, 0.13107487559318542f, 0.07015932351350784f};
__device__ const signed char d_pattern[256][4] = {
#include "orb_pattern.inc"
};

// ---- mask level 0 from boxes (xywh, source-frame pixels): rect grown by margin, scaled, floor/ceil, clipped ---------------
__global__ void mask_rects_kernel(uint8_t* __restrict__ mask, size_t slab, const float* __restrict__ boxes, const int* __restrict__ nboxes,
                                  int max_det, int slot0, int w, int h, float margin, float ratio) {
  const int slot = slot0 + blockIdx.y;
  const int n = min(nboxes[slot], max_det);
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
  const float* b = boxes + ((size_t)slot * max_det + i) * 4;
  const double gw = (double)b[2] * (1.0 + (double)margin), gh = (double)b[3] * (1.0 + (double)margin);
  int x0 = (int)floor(((double)b[0] - gw / 2) * (double)ratio), y0 = (int)floor(((double)b[1] - gh / 2) * (double)ratio);
  int x1 = (int)ceil(((double)b[0] + gw / 2) * (double)ratio), y1 = (int)ceil(((double)b[1] + gh / 2) * (double)ratio);
  x0 = min(max(x0, 0), w); x1 = min(max(x1, 0), w); y0 = min(max(y0, 0), h); y1 = min(max(y1, 0), h);
  uint8_t* m = mask + (size_t)slot * slab;
  const int rw = x1 - x0;
  for (int y = y0 + (int)(threadIdx.x / 32); y < y1; y += blockDim.x / 32)
    for (int x = threadIdx.x % 32; x < rw; x += 32) m[(size_t)y * w + x0 + x] = 0;
  }
}

// ---- sparse mask pyramid: a level of the vehicle-mask pyramid recomputed only around the boxes ---------------------------------------
// The mask is 255 everywhere except inside the (grown) vehicle rectangles -- ~3 % of the pixels for the 132 boxes of a frame -- and a
// pyramid pixel whose 2 x 2 source pixels are all 255 is 255 exactly ((255*256*256 + 32768) >> 16).  So every level is preset to 255 and
// only the pixels whose source footprint can reach a rectangle are computed, with pyr_resize_kernel<true>'s arithmetic (chained
// INTER_LINEAR_EXACT + threshold): bit-identical to the dense chain, ~30x less work.  Levels are separate launches because a pixel near
// two boxes needs the finished previous level of both (overlapping regions write equal values); a block walks boxes b, b + gridDim.x, ...
// (A single launch with one block per frame and __syncthreads between levels was measured at 0.9-2.5 ms: one block has too little
// memory-level parallelism for the dependent table -> pixel loads; the seven small launches take ~250 us beside the FAST kernel.)
struct MaskLevels { int w[GT_ORB_LEVELS], h[GT_ORB_LEVELS]; unsigned long long off[GT_ORB_LEVELS]; double sx[GT_ORB_LEVELS], sy[GT_ORB_LEVELS]; };

__global__ void __launch_bounds__(128) mask_pyr_sparse_kernel(uint8_t* __restrict__ mask, size_t slab, const float* __restrict__ boxes, const int* __restrict__ nboxes,
                                                              int max_det, int slot0, float margin, float ratio, const MaskLevels ml, int level,
                                                              const int* __restrict__ xofs, const int* __restrict__ xc1, const int* __restrict__ yofs,
                                                              const int* __restrict__ yc1) {
  const int slot = slot0 + blockIdx.y;
  const int nb = min(nboxes[slot], max_det);
  uint8_t* base = mask + (size_t)slot * slab;
  const uint8_t* src = base + ml.off[level - 1];
  uint8_t* dst = base + ml.off[level];
  const int sw = ml.w[level - 1], sh = ml.h[level - 1], dw = ml.w[level];
  for (int bi = blockIdx.x; bi < nb; bi += gridDim.x) {
    const float* b = boxes + ((size_t)slot * max_det + bi) * 4;
    // level-0 rectangle exactly as mask_rects_kernel draws it, then the range of each further level that can see it (conservative by
    // one pixel: extra pixels are computed with the same arithmetic, hence exact)
    const double gw = (double)b[2] * (1.0 + (double)margin), gh = (double)b[3] * (1.0 + (double)margin);
    int x0 = (int)floor(((double)b[0] - gw / 2) * (double)ratio), y0 = (int)floor(((double)b[1] - gh / 2) * (double)ratio);
    int x1 = (int)ceil(((double)b[0] + gw / 2) * (double)ratio) - 1, y1 = (int)ceil(((double)b[1] + gh / 2) * (double)ratio) - 1;   // inclusive
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, ml.w[0] - 1); y1 = min(y1, ml.h[0] - 1);
    if (x1 < x0 || y1 < y0) continue;
    for (int l = 1; l <= level; ++l) {
      x0 = max((int)floor(((double)x0 - 1.5) / ml.sx[l] - 0.5) - 1, 0); x1 = min((int)ceil(((double)x1 + 1.5) / ml.sx[l]) + 1, ml.w[l] - 1);
      y0 = max((int)floor(((double)y0 - 1.5) / ml.sy[l] - 0.5) - 1, 0); y1 = min((int)ceil(((double)y1 + 1.5) / ml.sy[l]) + 1, ml.h[l] - 1);
    }
    const int rw = x1 - x0 + 1, n = rw * (y1 - y0 + 1);
    for (int i = threadIdx.x; i < n; i += 128) {
      const int yy = i / rw, y = y0 + yy, x = x0 + i - yy * rw;
      const int xo = __ldg(xofs + x), c1 = __ldg(xc1 + x), c0 = 256 - c1, yo = __ldg(yofs + y), cy1 = __ldg(yc1 + y), cy0 = 256 - cy1;
      const uint8_t* r0 = src + (size_t)yo * sw;
      const uint8_t* r1 = src + (size_t)min(yo + 1, sh - 1) * sw;
      const int xb = min(xo + 1, sw - 1);
      const int h0 = (int)r0[xo] * c0 + (int)r0[xb] * c1, h1 = (int)r1[xo] * c0 + (int)r1[xb] * c1;
      int v = (h0 * cy0 + h1 * cy1 + 32768) >> 16;
      if (v <= 254) v = 0;
      dst[(size_t)y * dw + x] = (uint8_t)v;
    }
  }
}

// ---- one pyramid level from the previous one: cv2.resize(INTER_LINEAR_EXACT) in Q8.8 x Q8.8, (v + 2^15) >> 16 -------------
// One thread = 4 adjacent output pixels.  MASK = false: image plane.  MASK = true: mask plane, thresholded (<= 254 -> 0); the
// mask pyramid depends on the detections and is built later than the image pyramid (see orb_front / orb_back).
constexpr int kResizeRows = 4;
template <bool MASK>
__global__ void __launch_bounds__(128) pyr_resize_kernel(uint8_t* __restrict__ plane, size_t slab, int slot0, size_t src_off, int sw, int sh,
                                                         size_t dst_off, int dw, int dh, const int* __restrict__ xofs, const int* __restrict__ xc1,
                                                         const int* __restrict__ yofs, const int* __restrict__ yc1) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x4 >= dw) return;
  const int slot = slot0 + blockIdx.z;
  uint8_t* base = plane + (size_t)slot * slab;
  // the tables are padded to a multiple of four entries: one 16-byte load each
  const int4 xo = __ldg(reinterpret_cast<const int4*>(xofs + x4)), xc = __ldg(reinterpret_cast<const int4*>(xc1 + x4));
  const int a[4] = {xo.x, xo.y, xo.z, xo.w}, c1v[4] = {xc.x, xc.y, xc.z, xc.w};
  const int nvalid = min(4, dw - x4);
  // the four outputs read source columns a[0] .. a[3] + 1 <= a[0] + 7 (scale 1.2): an 8-byte window per source row, fetched as three
  // aligned words (the rows of a level have arbitrary alignment) and shifted into place.  A thread produces kResizeRows output rows
  // and issues all of their loads before the first use (the kernel is latency bound, not bandwidth bound).
  const uintptr_t hi = (reinterpret_cast<uintptr_t>(base + src_off) + (size_t)sw * sh - 1) & ~(uintptr_t)3;
  const int y0 = blockIdx.y * kResizeRows;
  uint32_t w[kResizeRows][2][3];
  int cy[kResizeRows], sft[kResizeRows][2];
#pragma unroll
  for (int i = 0; i < kResizeRows; ++i) {
    const int y = min(y0 + i, dh - 1);
    const int yo = __ldg(yofs + y);
    cy[i] = __ldg(yc1 + y);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uintptr_t p = reinterpret_cast<uintptr_t>(base + src_off + (size_t)min(yo + r, sh - 1) * sw + a[0]);
      const uintptr_t p0 = p & ~(uintptr_t)3;
      w[i][r][0] = __ldg(reinterpret_cast<const uint32_t*>(min(p0, hi)));
      w[i][r][1] = __ldg(reinterpret_cast<const uint32_t*>(min(p0 + 4, hi)));
      w[i][r][2] = __ldg(reinterpret_cast<const uint32_t*>(min(p0 + 8, hi)));
      sft[i][r] = 8 * (int)(p & 3);
    }
  }
  // per output column: a byte-permute selector that picks the two source bytes (a[k], min(a[k] + 1, sw - 1)) out of the 8-byte window, and the
  // two horizontal weights as a 16-bit pair -> one PRMT + one IDP.2A per source row instead of 64-bit shifts, masks and two IMADs
  uint32_t sel[4], cw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int o0 = a[k] - a[0], o1 = min(a[k] + 1, sw - 1) - a[0];     // 0 .. 7
    sel[k] = (uint32_t)o0 | ((uint32_t)o1 << 4);                       // (bytes 2, 3 of the result are don't-care: dp2a_lo reads bytes 0, 1)
    cw[k] = (uint32_t)(256 - c1v[k]) | ((uint32_t)c1v[k] << 16);
  }
#pragma unroll
  for (int i = 0; i < kResizeRows; ++i) {
    const int y = y0 + i;
    if (y >= dh) break;
    uint32_t lo[2], hi2[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      lo[r] = __funnelshift_r(w[i][r][0], w[i][r][1], sft[i][r]);
      hi2[r] = __funnelshift_r(w[i][r][1], w[i][r][2], sft[i][r]);
    }
    const int cy1 = cy[i], cy0 = 256 - cy1;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int h0 = (int)__dp2a_lo(cw[k], __byte_perm(lo[0], hi2[0], sel[k]), 0u);
      const int h1 = (int)__dp2a_lo(cw[k], __byte_perm(lo[1], hi2[1], sel[k]), 0u);
      int v = (h0 * cy0 + h1 * cy1 + 32768) >> 16;
      if (MASK && v <= 254) v = 0;
      packed |= (uint32_t)v << (8 * k);
    }
    uint8_t* d = base + dst_off + (size_t)y * dw + x4;
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(d) & 3) == 0) *reinterpret_cast<uint32_t*>(d) = packed;
    else
      for (int k = 0; k < nvalid; ++k) d[k] = (uint8_t)(packed >> (8 * k));
  }
}

// ---- FAST-9/16 score + 3x3 non-max suppression + border filter -> candidate list --------------------------------------------
// One launch covers all pyramid levels (blockIdx.x walks the 64 x 64 tiles of every level's candidate region
// [kEdge, w - kEdge) x [kEdge, h - kEdge); blockIdx.y = frame slot).  Per tile:
//   0. stage the tile + apron in shared memory with aligned 32-bit global loads (rows of the pyramid have arbitrary alignment: two
//      aligned words + a funnel shift per staged word);
//   1a. compass quick-reject (ring pixels 0, 4, 8, 12: any 9-arc contains two of them) on FOUR horizontally adjacent pixels per
//      thread with byte-SIMD compares; groups with a surviving pixel are compacted;
//   1b. the full 16-pixel arc test on the surviving groups, again four pixels per thread: 16 ring words (aligned words + funnel
//      shifts), byte masks "brighter" / "darker", 9-in-a-row by three-input ANDs; corners are appended to a shared list;
//   2. exact corner score (OpenCV cornerScore<16>) of the listed corners;
//   3. 3x3 non-max suppression and compaction into the per-(frame, level) candidate list.
// The vehicle mask is NOT applied here (it depends on the detections, which are computed concurrently): orb_select_kernel
// drops masked candidates before its score cut, which is equivalent to OpenCV's order (mask, then retain-best).
#define FT_X 64
#define FT_Y 64
#define FT_SW 76                         // staged columns: gx in [x0 - 5, x0 + 71)
#define FT_SH (FT_Y + 8)                 // staged rows:    gy in [y0 - 4, y0 + 36)
#define FT_G 17                          // 4-pixel groups per tested row: sx = 4 g + j, gx = x0 - 1 + sx
#define FT_ROWS (FT_Y + 2)               // tested rows: sy in [0, 34), gy = y0 - 1 + sy
#define FT_LIST ((FT_X + 2) * (FT_Y + 2))

struct FastLevels {                      // per-level geometry of the single FAST launch
  int w[GT_ORB_LEVELS], h[GT_ORB_LEVELS], tiles_x[GT_ORB_LEVELS], tile0[GT_ORB_LEVELS + 1], cand_cap[GT_ORB_LEVELS];
  unsigned long long off[GT_ORB_LEVELS], cand_off[GT_ORB_LEVELS];
};

__device__ __forceinline__ void fast_ring(const uint8_t (*t)[FT_SW], int x, int y, int v, int* d) {
  d[0] = v - t[y + 3][x];      d[1] = v - t[y + 3][x + 1];  d[2] = v - t[y + 2][x + 2];  d[3] = v - t[y + 1][x + 3];
  d[4] = v - t[y][x + 3];      d[5] = v - t[y - 1][x + 3];  d[6] = v - t[y - 2][x + 2];  d[7] = v - t[y - 3][x + 1];
  d[8] = v - t[y - 3][x];      d[9] = v - t[y - 3][x - 1];  d[10] = v - t[y - 2][x - 2]; d[11] = v - t[y - 1][x - 3];
  d[12] = v - t[y][x - 3];     d[13] = v - t[y + 1][x - 3]; d[14] = v - t[y + 2][x - 2]; d[15] = v - t[y + 3][x - 1];
}

__device__ __forceinline__ int fast_corner_score(const uint8_t (*t)[FT_SW], int x, int y) {
  const int v = t[y][x];
  int d[16];
  fast_ring(t, x, y, v, d);
  // exact corner score: the largest threshold for which the pixel is still a corner = max over the 16 circular 9-arcs of
  // min(d) (darker arcs) and of min(-d) (brighter arcs), minus 1.  Written as a doubling min-network on d and on an
  // explicitly negated copy: the straightforward "max(mn, -mx)" form is miscompiled by ptxas 12.9 -O1..-O3 for sm_100a
  // (wrong VIMNMX3 fusion; repro in tools/fast_test.cu: 7,325 wrong scores at -O3, 0 at -Xptxas -O0).
  int nd[16], a2[16], a4[16], a8[16], b2[16], b4[16], b8[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) nd[k] = -d[k];
#pragma unroll
  for (int k = 0; k < 16; ++k) { a2[k] = min(d[k], d[(k + 1) & 15]); b2[k] = min(nd[k], nd[(k + 1) & 15]); }
#pragma unroll
  for (int k = 0; k < 16; ++k) { a4[k] = min(a2[k], a2[(k + 2) & 15]); b4[k] = min(b2[k], b2[(k + 2) & 15]); }
#pragma unroll
  for (int k = 0; k < 16; ++k) { a8[k] = min(a4[k], a4[(k + 4) & 15]); b8[k] = min(b4[k], b4[(k + 4) & 15]); }
  int best = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    best = max(best, min(a8[k], d[(k + 8) & 15]));
    best = max(best, min(b8[k], nd[(k + 8) & 15]));
  }
  return best > kFastThr ? best - 1 : 0;
}

// Byte-wise unsigned compares with the result in the MSB of each byte only (the other bits are don't-care: every consumer is a
// bitwise AND / OR followed by a final mask with 0x80808080).  a > b  <=>  (a7 & ~b7) | (~(a7 ^ b7) & ~msb((b | H) - (a & ~H))): four
// instructions instead of the ~6 of the emulated __vcmpgtu4, and `r & ~H`, `r | H` are shared by the two polarities of a ring word.
struct RingCmp {            // per group: thresholds hi = sat(c + t), lo = sat(c - t)
  uint32_t hi, hi_h, lo, lo_l;
  __device__ __forceinline__ RingCmp(uint32_t c, uint32_t thr4) {
    hi = __vaddus4(c, thr4); lo = __vsubus4(c, thr4);
    hi_h = hi | 0x80808080u; lo_l = lo & 0x7f7f7f7fu;
  }
  __device__ __forceinline__ uint32_t brighter(uint32_t r) const {   // r > hi
    const uint32_t d = hi_h - (r & 0x7f7f7f7fu);
    return (r & ~hi) | (~(r ^ hi) & ~d);
  }
  __device__ __forceinline__ uint32_t darker(uint32_t r) const {     // lo > r
    const uint32_t d = (r | 0x80808080u) - lo_l;
    return (lo & ~r) | (~(lo ^ r) & ~d);
  }
};

// bytes [o, o + 4) of the 12-byte window (w0, w1, w2), o in 1..7
__device__ __forceinline__ uint32_t win4(uint32_t w0, uint32_t w1, uint32_t w2, int o) {
  return o < 4 ? __funnelshift_r(w0, w1, 8 * o) : (o == 4 ? w1 : __funnelshift_r(w1, w2, 8 * (o - 4)));
}
// per byte: set where two ADJACENT compass points (0-4, 4-8, 8-12 or 12-0) are both set.  Nine circularly consecutive ring positions always
// contain two adjacent compass points, so this is a necessary condition for a 9-arc -- strictly stronger than "any two of four" (which
// also passes the opposite pairs 0-8 / 4-12) and one instruction shorter: (a&b)|(b&c)|(c&d)|(d&a) = (a|c)&(b|d).
__device__ __forceinline__ uint32_t two_of_four(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return (a | c) & (b | d);
}
// per byte: 0xFF where some 9 circularly consecutive masks of m[0..15] are all set
__device__ __forceinline__ uint32_t arc9(const uint32_t* m) {
  uint32_t a3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a3[k] = m[k] & m[(k + 1) & 15] & m[(k + 2) & 15];
  uint32_t any = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) any |= a3[k] & a3[(k + 3) & 15] & a3[(k + 6) & 15];
  return any;
}

__global__ void __launch_bounds__(256, 5) fast_kernel(const uint8_t* __restrict__ img, size_t slab, int slot0, const FastLevels fl,
                                                   unsigned int* __restrict__ cand, uint8_t* __restrict__ cscore, size_t cand_slab,
                                                   int* __restrict__ counts, int tile_first) {
  __shared__ __align__(16) uint8_t s_img[FT_SH][FT_SW];
  __shared__ __align__(16) uint8_t s_sc_buf[((FT_Y + 2) * (FT_X + 2) + 15) & ~15];   // corner scores, cleared with 16-byte stores
  uint8_t (*s_sc)[FT_X + 2] = reinterpret_cast<uint8_t (*)[FT_X + 2]>(s_sc_buf);
  __shared__ int s_geo[5];
  __shared__ unsigned short s_cand[FT_ROWS * FT_G];   // (task << 4 | pixel mask) of the groups that pass the compass test
  __shared__ unsigned short s_list[FT_LIST];          // corner positions sy * (FT_X + 2) + sx
  __shared__ int s_na, s_nl, s_n, s_base;
  __shared__ unsigned int s_xy[FT_X * FT_Y / 4];
  __shared__ uint8_t s_s[FT_X * FT_Y / 4];
  // tile -> (level, tile row, tile column): one thread does the search and the division, the block reads the five numbers (the per-thread
  // form was 6 % of the kernel's instructions)
  if (threadIdx.x == 0) {
    const int gtile = (int)blockIdx.x + tile_first;   // (a launch may cover a sub-range of the levels: orb_fast)
    int lv = 0;
#pragma unroll
    for (int l = 1; l < GT_ORB_LEVELS; ++l) lv += gtile >= fl.tile0[l];
    const int tl = gtile - fl.tile0[lv];
    const int ty_ = tl / fl.tiles_x[lv];
    s_geo[0] = lv; s_geo[1] = fl.w[lv]; s_geo[2] = fl.h[lv]; s_geo[3] = ty_; s_geo[4] = tl - ty_ * fl.tiles_x[lv];
    s_n = 0; s_nl = 0; s_na = 0;
  }
  for (int i = threadIdx.x; i < (int)(sizeof(s_sc_buf) / 16); i += 256) reinterpret_cast<uint4*>(s_sc_buf)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const int level = s_geo[0], w = s_geo[1], h = s_geo[2], by = s_geo[3], bx = s_geo[4];
  const int slot = slot0 + blockIdx.y;
  const uint8_t* im = img + (size_t)slot * slab + fl.off[level];
  const int x0 = kEdge + bx * FT_X, y0 = kEdge + by * FT_Y;
  // phase 0: staged word (ty, tq) = image bytes gx = x0 - 5 + 4 tq .. + 3 of row gy = y0 - 4 + ty (rows clamped into the image;
  // columns outside it only feed positions that are never tested).  All loads of a thread are issued before the first use.
  {
    const int mis = (int)(reinterpret_cast<uintptr_t>(im) & 3);
    const uint32_t* base = reinterpret_cast<const uint32_t*>(im - mis);   // 4-byte aligned; word indices clamped to the level's words
    const int last_word = (mis + w * h - 1) >> 2;
    constexpr int kWords = FT_SH * (FT_SW / 4), kIter = (kWords + 255) / 256;
    uint32_t v0[kIter], v1[kIter];
    int sh[kIter];
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int i = threadIdx.x + it * 256;
      v0[it] = v1[it] = 0; sh[it] = 0;
      if (i < kWords) {
        const int ty = i / (FT_SW / 4), tq = i - ty * (FT_SW / 4);
        const int gy = min(max(y0 - 4 + ty, 0), h - 1);
        const int boff = mis + gy * w + (x0 - 5 + 4 * tq);
        const int wi = boff >> 2;
        v0[it] = __ldg(base + min(wi, last_word));
        v1[it] = __ldg(base + min(wi + 1, last_word));
        sh[it] = (boff & 3) * 8;
      }
    }
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int i = threadIdx.x + it * 256;
      if (i < kWords) reinterpret_cast<uint32_t*>(&s_img[0][0])[i] = __funnelshift_r(v0[it], v1[it], sh[it]);
    }
  }
  __syncthreads();
  const bool right_edge = x0 + FT_X + 1 > w - kEdge;   // only the last tile column can run past the tested range
  const uint32_t thr4 = 0x01010101u * (uint32_t)kFastThr;
  // tested positions: gx in [kEdge - 1, w - kEdge] (candidates + their 3x3 neighbours), same for gy
  // phase 1a: compass quick-reject, four pixels per thread
  for (int t0 = 0; t0 < FT_ROWS * FT_G; t0 += 256) {
    const int t = t0 + threadIdx.x;
    uint32_t pass = 0;
    if (t < FT_ROWS * FT_G) {
      const int sy = t / FT_G, g = t - sy * FT_G;
      const int gy = y0 - 1 + sy;
      if (gy >= kEdge - 1 && gy <= h - kEdge) {
        const uint32_t* rc = reinterpret_cast<const uint32_t*>(&s_img[sy + 3][0]) + g;
        const uint32_t w0 = rc[0], c = rc[1], w2 = rc[2];
        const uint32_t r0 = reinterpret_cast<const uint32_t*>(&s_img[sy + 6][0])[g + 1], r8 = reinterpret_cast<const uint32_t*>(&s_img[sy][0])[g + 1];
        const uint32_t r4 = __funnelshift_r(c, w2, 24), r12 = __funnelshift_r(w0, c, 8);
        const RingCmp rc4(c, thr4);
        const uint32_t br = two_of_four(rc4.brighter(r0), rc4.brighter(r4), rc4.brighter(r8), rc4.brighter(r12));
        const uint32_t dk = two_of_four(rc4.darker(r0), rc4.darker(r4), rc4.darker(r8), rc4.darker(r12));
        uint32_t m = (br | dk) & 0x80808080u;
        // validity of the four pixels (sx = 4 g + j): j in [0, jhi] -- the left limit never cuts (gx >= x0 - 1 >= kEdge - 1); the right one
        // only in the last group of a row (two of its four pixels lie beyond the tile's apron) and in the last tile column
        if (g == FT_G - 1 || right_edge) {
          const int gx = x0 - 1 + 4 * g;
          const int jhi = min(min(w - kEdge - gx, FT_X + 1 - 4 * g), 3);
          if (jhi < 3) m = jhi < 0 ? 0u : (m & (0xFFFFFFFFu >> (8 * (3 - jhi))));
        }
        pass = m;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass != 0);
    if (bal) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_na, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pass) {
        const unsigned bits = ((pass >> 7) & 1u) | ((pass >> 14) & 2u) | ((pass >> 21) & 4u) | ((pass >> 28) & 8u);
        s_cand[base + __popc(bal & ((1u << lane) - 1))] = (unsigned short)((t << 4) | bits);
      }
    }
  }
  __syncthreads();
  // phase 1b: full arc test of the surviving groups
  const int na = s_na;
  for (int k = threadIdx.x; k < na; k += 256) {
    const unsigned e = s_cand[k];
    const int t = e >> 4;
    const int sy = t / FT_G, g = t - sy * FT_G;
    uint32_t r[16], c;
    {
      const uint32_t* q;
      uint32_t w0, w1, w2;
#define FT_ROW(dy) q = reinterpret_cast<const uint32_t*>(&s_img[sy + 3 + (dy)][0]) + g; w0 = q[0]; w1 = q[1]; w2 = q[2];
      FT_ROW(3)  r[15] = win4(w0, w1, w2, 3); r[0] = w1; r[1] = win4(w0, w1, w2, 5);
      FT_ROW(2)  r[14] = win4(w0, w1, w2, 2); r[2] = win4(w0, w1, w2, 6);
      FT_ROW(1)  r[13] = win4(w0, w1, w2, 1); r[3] = win4(w0, w1, w2, 7);
      FT_ROW(0)  r[12] = win4(w0, w1, w2, 1); r[4] = win4(w0, w1, w2, 7); c = w1;
      FT_ROW(-1) r[11] = win4(w0, w1, w2, 1); r[5] = win4(w0, w1, w2, 7);
      FT_ROW(-2) r[10] = win4(w0, w1, w2, 2); r[6] = win4(w0, w1, w2, 6);
      FT_ROW(-3) r[9] = win4(w0, w1, w2, 3); r[8] = w1; r[7] = win4(w0, w1, w2, 5);
#undef FT_ROW
    }
    const RingCmp rc4(c, thr4);
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = rc4.brighter(r[i]);
    uint32_t corner = arc9(m);
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = rc4.darker(r[i]);
    corner |= arc9(m);
    unsigned bits = (((corner >> 7) & 1u) | ((corner >> 14) & 2u) | ((corner >> 21) & 4u) | ((corner >> 28) & 8u)) & (e & 15u);
    if (bits) {
      const int base = atomicAdd(&s_nl, __popc(bits));
      int n = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (bits & (1u << j)) s_list[base + n++] = (unsigned short)(sy * (FT_X + 2) + 4 * g + j);
    }
  }
  __syncthreads();
  // phase 2: exact score of the listed corners (tile coordinates: x = sx + 4, y = sy + 3)
  for (int k = threadIdx.x; k < s_nl; k += 256) {
    const int i = s_list[k];
    const int sy = i / (FT_X + 2), sx = i - sy * (FT_X + 2);
    s_sc[sy][sx] = (uint8_t)fast_corner_score(s_img, sx + 4, sy + 3);
  }
  __syncthreads();
  // phase 3: 3x3 NMS over the listed corners that lie inside the tile, border filter
  for (int k = threadIdx.x; k < s_nl; k += 256) {
    const int i = s_list[k];
    const int sy = i / (FT_X + 2), sx = i - sy * (FT_X + 2);
    if (sx < 1 || sx > FT_X || sy < 1 || sy > FT_Y) continue;
    const int lx = sx - 1, ly = sy - 1;
    const int gx = x0 + lx, gy = y0 + ly;
    const int sc = s_sc[sy][sx];
    if (sc < kFastThr) continue;
    if (gx < kEdge || gx >= w - kEdge || gy < kEdge || gy >= h - kEdge) continue;
    if (!(sc > s_sc[ly][lx] && sc > s_sc[ly][lx + 1] && sc > s_sc[ly][lx + 2] && sc > s_sc[ly + 1][lx] && sc > s_sc[ly + 1][lx + 2] &&
          sc > s_sc[ly + 2][lx] && sc > s_sc[ly + 2][lx + 1] && sc > s_sc[ly + 2][lx + 2]))
      continue;
    const int q = atomicAdd(&s_n, 1);  // NMS guarantees <= 1 survivor per 2x2 block, so q < FT_X*FT_Y/4
    s_xy[q] = ((unsigned)gy << 16) | (unsigned)gx;
    s_s[q] = (uint8_t)sc;
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_n) s_base = atomicAdd(&counts[slot * GT_ORB_LEVELS + level], s_n);
  __syncthreads();
  const int cap = fl.cand_cap[level];
  for (int k = threadIdx.x; k < s_n; k += 256) {
    const int dst = s_base + k;
    if (dst < cap) {
      cand[(size_t)slot * cand_slab + fl.cand_off[level] + dst] = s_xy[k];
      cscore[(size_t)slot * cand_slab + fl.cand_off[level] + dst] = s_s[k];
    }
  }
}

// ---- per (frame, level): FAST-score cut (2 x quota, ties kept) -> Harris -> quota cut (ties kept) -> sort by position ------
__device__ __forceinline__ float harris_at(const uint8_t* im, int w, int x0, int y0) {
  // 7 x 7 block of 3 x 3 Sobel responses = a 9 x 9 window: three rows of nine pixels live in registers and slide down, so every
  // pixel is loaded once (81 loads instead of 588; the loop was bound by its byte-load latency).  Integer sums: order-free, exact.
  int a = 0, b = 0, c = 0;
  int r0[9], r1[9], r2[9];
  const uint8_t* p = im + (size_t)(y0 - 4) * w + (x0 - 4);
#pragma unroll
  for (int i = 0; i < 9; ++i) { r0[i] = p[i]; r1[i] = p[w + i]; }
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    const uint8_t* q = p + (size_t)(dy + 2) * w;
#pragma unroll
    for (int i = 0; i < 9; ++i) r2[i] = q[i];
#pragma unroll
    for (int i = 1; i <= 7; ++i) {
      const int Ix = (r1[i + 1] - r1[i - 1]) * 2 + (r0[i + 1] - r0[i - 1]) + (r2[i + 1] - r2[i - 1]);
      const int Iy = (r2[i] - r0[i]) * 2 + (r2[i - 1] - r0[i - 1]) + (r2[i + 1] - r0[i + 1]);
      a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) { r0[i] = r1[i]; r1[i] = r2[i]; }
  }
  const float scale = 1.0f / (4.0f * 7.0f * 255.0f);
  const float s2 = __fmul_rn(scale, scale), s4 = __fmul_rn(s2, s2);
  const float fa = (float)a, fb = (float)b, fc = (float)c;
  const float ab = __fadd_rn(fa, fb);
  const float r = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, ab), ab));
  return __fmul_rn(r, s4);
}

__device__ __forceinline__ unsigned f2key(float f) {  // order-preserving float -> uint
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Warp-parallel "walk the 256-bin histogram from the top until `need` entries are covered": returns the bucket where the running
// count reaches `need` and the count accumulated ABOVE that bucket (both in every lane); bucket = -1 if the total is below `need`.
// Lane l owns buckets 255 - 8 l .. 248 - 8 l (descending), so lane order = descending bucket order.
__device__ __forceinline__ void hist_cut_warp(const int* __restrict__ hist, int need, int* bucket, int* above) {
  const int lane = threadIdx.x & 31;
  int mine = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine += hist[255 - 8 * lane - j];
  int incl = mine;                              // inclusive prefix over lanes (descending buckets)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
  int bkt = -1, acc = 0;
  if (hit) {
    const int src = __ffs(hit) - 1;             // first lane whose running count reaches `need`
    if (lane == src) {
      acc = incl - mine;
      for (int j = 0; j < 8; ++j) {
        const int bb = 255 - 8 * lane - j;
        if (acc + hist[bb] >= need) { bkt = bb; break; }
        acc += hist[bb];
      }
    }
    bkt = __shfl_sync(0xffffffffu, bkt, src);
    acc = __shfl_sync(0xffffffffu, acc, src);
  }
  *bucket = bkt;
  *above = acc;
}

__global__ void __launch_bounds__(1024) orb_select_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ msk, size_t slab, int slot0,
                                                          const OrbLevel* __restrict__ lv,
                                                          const unsigned int* __restrict__ cand, const uint8_t* __restrict__ cscore,
                                                          size_t cand_slab, const int* __restrict__ fast_count, int as_ref,
                                                          unsigned int* __restrict__ sel_xy, float* __restrict__ sel_resp,
                                                          int* __restrict__ sel_count) {
  extern __shared__ unsigned char s_raw[];
  unsigned* s_keys = reinterpret_cast<unsigned*>(s_raw);                           // [kSelCap] response keys
  unsigned long long* s_sort = reinterpret_cast<unsigned long long*>(s_raw);       // [kLvlKeep] aliases s_keys after the cut
  __shared__ int s_hist[256];
  __shared__ int s_cut, s_m, s_k, s_need, s_nu;
  __shared__ unsigned s_prefix;
  const int level = blockIdx.x, slot = slot0 + blockIdx.y;
  const OrbLevel L = lv[level];
  const int quota = as_ref ? L.quota_ref : L.quota_cur;
  const int n = min(fast_count[slot * GT_ORB_LEVELS + level], L.cand_cap);
  const unsigned int* cxy = cand + (size_t)slot * cand_slab + L.cand_off;
  const uint8_t* csc = cscore + (size_t)slot * cand_slab + L.cand_off;
  unsigned int* oxy = sel_xy + ((size_t)slot * GT_ORB_LEVELS + level) * kSelCap;
  float* orsp = sel_resp + ((size_t)slot * GT_ORB_LEVELS + level) * kSelCap;
  const uint8_t* im = img + (size_t)slot * slab + L.off;
  const uint8_t* mk = msk + (size_t)slot * slab + L.off;
  auto unmasked = [&](unsigned xy) { return mk[(size_t)(xy >> 16) * L.w + (xy & 0xFFFF)] != 0; };

  // (a) FAST-score histogram over the unmasked candidates and cut: keep score >= score of the (2*quota)-th best
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
  if (threadIdx.x == 0) { s_m = 0; s_k = 0; }
  __syncthreads();
  if (threadIdx.x == 0) s_nu = 0;
  __syncthreads();
  // The loop is latency bound (candidate -> mask address -> mask byte): four candidates per thread are in flight at a time, the
  // mask verdicts are kept as a per-thread bit mask for pass (b), and lanes with equal scores merge their histogram updates.
  int my_unmasked = 0;
  unsigned long long um_bits = 0;       // bit j: this thread's j-th candidate (i = threadIdx.x + j * blockDim.x) is unmasked
  const bool bits_ok = n <= 64 * (int)blockDim.x;
  for (int i0 = threadIdx.x, j0 = 0; i0 < n; i0 += 4 * blockDim.x, j0 += 4) {
    unsigned xy[4];
    int sc[4];
    bool um[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      xy[u] = i < n ? cxy[i] : 0u;
      sc[u] = i < n ? (int)csc[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) um[u] = (i0 + u * (int)blockDim.x < n) && unmasked(xy[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned active = __ballot_sync(__activemask(), um[u]);
      if (um[u]) {
        const unsigned peers = __match_any_sync(active, sc[u]);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&s_hist[sc[u]], __popc(peers));
        ++my_unmasked;
        if (bits_ok) um_bits |= 1ull << (j0 + u);
      }
    }
  }
  if (my_unmasked) atomicAdd(&s_nu, my_unmasked);
  __syncthreads();
  if (threadIdx.x < 32) {
    int cut = 0;
    const int want = 2 * quota;
    if (s_nu > want) {
      int bkt, above;
      hist_cut_warp(s_hist, want, &bkt, &above);
      cut = max(bkt, 0);
    }
    if (threadIdx.x == 0) s_cut = cut;
  }
  __syncthreads();
  const int cut = s_cut;
  // (b) compact survivors + Harris response
  for (int i = threadIdx.x, j = 0; i < n; i += blockDim.x, ++j) {
    const bool um = bits_ok ? ((um_bits >> j) & 1ull) != 0 : unmasked(cxy[i]);
    if (um && csc[i] >= cut) {
      const int k = atomicAdd(&s_m, 1);
      if (k < kSelCap) {
        const unsigned xy = cxy[i];
        const float r = harris_at(im, L.w, (int)(xy & 0xFFFF), (int)(xy >> 16));
        oxy[k] = xy;
        orsp[k] = r;
        s_keys[k] = f2key(r);
      }
    }
  }
  __syncthreads();
  const int m = min(s_m, kSelCap);
  // (c) Harris cut: key of the quota-th largest response (MSB-first radix select), ties kept
  unsigned cutkey = 0;
  if (m > quota && quota > 0) {
    if (threadIdx.x == 0) { s_prefix = 0; s_need = quota; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
      __syncthreads();
      const unsigned prefix = s_prefix;
      const unsigned himask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const unsigned k = s_keys[i];
        if ((k & himask) == (prefix & himask)) atomicAdd(&s_hist[(k >> shift) & 0xFF], 1);
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        int bkt, above;
        const int need = s_need;
        hist_cut_warp(s_hist, need, &bkt, &above);
        if (threadIdx.x == 0 && bkt >= 0) { s_prefix = prefix | ((unsigned)bkt << shift); s_need = need - above; }
      }
      __syncthreads();
    }
    cutkey = s_prefix;
  } else if (quota <= 0) {
    cutkey = 0xFFFFFFFFu;
  }
  // (d) gather the kept ones (position-major key) and sort ascending by (y, x)
  unsigned long long mine[(kSelCap + 1023) / 1024];
  int nmine = 0;
  for (int i = threadIdx.x; i < m; i += blockDim.x)
    if (s_keys[i] >= cutkey && quota > 0) mine[nmine++] = ((unsigned long long)oxy[i] << 32) | (unsigned long long)__float_as_uint(orsp[i]);
  __syncthreads();  // s_keys is dead from here on; s_sort aliases it
  for (int j = 0; j < nmine; ++j) {
    const int k = atomicAdd(&s_k, 1);
    if (k < kLvlKeep) s_sort[k] = mine[j];
  }
  __syncthreads();
  const int kept = min(s_k, kLvlKeep);
  int np2 = 1;
  while (np2 < kept) np2 <<= 1;
  for (int i = kept + threadIdx.x; i < np2; i += blockDim.x) s_sort[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        const unsigned long long a = s_sort[lo], c = s_sort[hi];
        if ((a > c) == asc) { s_sort[lo] = c; s_sort[hi] = a; }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < kept; i += blockDim.x) {
    oxy[i] = (unsigned)(s_sort[i] >> 32);
    orsp[i] = __uint_as_float((unsigned)(s_sort[i] & 0xFFFFFFFFull));
  }
  if (threadIdx.x == 0) sel_count[slot * GT_ORB_LEVELS + level] = kept;
}

// ---- one warp per key point: intensity-centroid angle, key-point record, 256-bit rotated BRIEF -------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = 0.9997878412794807f * 57.29577951308232f, p3 = -0.3258083974640975f * 57.29577951308232f;
  const float p5 = 0.1555786518463281f * 57.29577951308232f, p7 = -0.04432655554792128f * 57.29577951308232f;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, ax + 2.220446049250313e-16f);
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, ay + 2.220446049250313e-16f);
    c2 = __fmul_rn(c, c);
    a = 90.f - __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  }
  if (x < 0.f) a = 180.f - a;
  if (y < 0.f) a = 360.f - a;
  return a;
}

// Per key point (one warp): the 45 x 45 source patch is staged in shared memory; the intensity centroid reads it; the 7-tap
// horizontal Gaussian pass runs over the whole patch (45 rows x 39 columns, one lane per half row with the row's bytes held in
// registers); the vertical pass is evaluated ON DEMAND at the 512 rotated test points only (~1/3 of the 39 x 39 blurred pixels, each
// 7 shared-memory reads).  float32 separable 7x7 sigma-2 blur with every product and sum rounded separately in OpenCV's
// accumulation order -- bit-identical to blurring the whole level, which is what cv2.ORB does, because key points stay 31 px away
// from the border so no reflection is involved.  (Round 1 blurred the full 39 x 39 patch with per-element index arithmetic: 6,800
// warp instructions per key point; this form needs ~1,700.)
constexpr int kDescWarps = 8;
constexpr int kSrcR = 22, kSrcW = 2 * kSrcR + 1;       // 45: descriptor reach 19 (pattern radius 13*sqrt2 rounded) + blur 3
constexpr int kBlurR = 19, kBlurW = 2 * kBlurR + 1;    // 39
constexpr int kSrcPitchW = 13;                         // source rows are 13 words (52 B) apart: lanes on different rows hit different banks
constexpr int kDescSmemPerWarp = kSrcW * kSrcPitchW * 4 + kSrcW * kBlurW * 4;   // 2340 + 7020

__global__ void __launch_bounds__(kDescWarps * 32) orb_describe_kernel(const uint8_t* __restrict__ img, size_t slab, int slot0,
                                                                       const OrbLevel* __restrict__ lv, const unsigned int* __restrict__ sel_xy,
                                                                       const float* __restrict__ sel_resp, const int* __restrict__ sel_count,
                                                                       float* __restrict__ kp_all, uint8_t* __restrict__ desc_all,
                                                                       int* __restrict__ kp_count, int* __restrict__ lvl_kp_off) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ signed char s_pat[256][4];
  __shared__ int s_off[GT_ORB_LEVELS + 1];
  const int slot = slot0 + blockIdx.y;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) *reinterpret_cast<uint32_t*>(s_pat[i]) = *reinterpret_cast<const uint32_t*>(d_pattern[i]);
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < GT_ORB_LEVELS; ++l) {
      s_off[l] = acc;
      acc = min(acc + sel_count[slot * GT_ORB_LEVELS + l], GT_MAX_KP);
    }
    s_off[GT_ORB_LEVELS] = acc;
    if (blockIdx.x == 0) {
      kp_count[slot] = acc;
      for (int l = 0; l <= GT_ORB_LEVELS; ++l) lvl_kp_off[slot * (GT_ORB_LEVELS + 1) + l] = s_off[l];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kpi = blockIdx.x * kDescWarps + warp;
  if (kpi >= s_off[GT_ORB_LEVELS]) return;
  int level = 0;
  while (kpi >= s_off[level + 1]) ++level;
  const int li = kpi - s_off[level];
  const OrbLevel L = lv[level];
  const unsigned xy = sel_xy[((size_t)slot * GT_ORB_LEVELS + level) * kSelCap + li];
  const int x = (int)(xy & 0xFFFF), y = (int)(xy >> 16);
  const uint8_t* im = img + (size_t)slot * slab + L.off;
  unsigned char* base = s_dyn + (size_t)warp * ((kDescSmemPerWarp + 15) & ~15);
  uint8_t* s_src = base;                                             // [45][52]
  const uint32_t* s_srcw = reinterpret_cast<const uint32_t*>(base);  // the same rows as words
  float* s_hp = reinterpret_cast<float*>(base + kSrcW * kSrcPitchW * 4);   // [45][39] horizontally blurred rows
  // stage the source patch: key points are >= 31 px from every border (edgeThreshold), the patch reaches 22 (+ 3 bytes of word slack) -> always
  // inside the image.  Two rows per step (one per half-warp): lane j loads ALIGNED word j of its row, all 23 loads of a lane are issued
  // before the first use, and word j of the patch row is funnel-shifted out of the aligned words j, j + 1 (shuffle) -- the byte-wise form
  // (90 LDG.U8 + 90 STS.U8 per lane, five rows in flight) was 41 % of the kernel's stall samples.
  {
    const uint8_t* p0 = im + (size_t)(y - kSrcR) * L.w + (x - kSrcR);
    const int half = lane >> 4, j = lane & 15;
    constexpr int kSteps = (kSrcW + 1) / 2;   // 23
    uint32_t wv[kSteps];
#pragma unroll
    for (int st = 0; st < kSteps; ++st) {
      const int r = min(2 * st + half, kSrcW - 1);
      const uintptr_t pa = reinterpret_cast<uintptr_t>(p0 + (size_t)r * L.w);
      wv[st] = j < 13 ? __ldg(reinterpret_cast<const uint32_t*>(pa & ~(uintptr_t)3) + j) : 0u;
    }
    uint32_t* s_w = reinterpret_cast<uint32_t*>(base);
#pragma unroll
    for (int st = 0; st < kSteps; ++st) {
      const int r = 2 * st + half;
      const uint32_t nx = __shfl_down_sync(0xffffffffu, wv[st], 1);
      const int sh = 8 * (int)(reinterpret_cast<uintptr_t>(p0 + (size_t)min(r, kSrcW - 1) * L.w) & 3);
      if (j < 12 && r < kSrcW) s_w[r * kSrcPitchW + j] = __funnelshift_r(wv[st], nx, sh);
    }
  }
  __syncwarp();
  // intensity centroid over the radius-15 disc: lane = column u
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int u = lane - 15, au = abs(u);
    for (int v = -15; v <= 15; ++v) {
      if (au <= c_umax[abs(v)]) {
        const int I = s_src[(kSrcR + v) * (kSrcPitchW * 4) + kSrcR + u];
        m10 += u * I;
        m01 += v * I;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  const float angle = fast_atan2_deg((float)m01, (float)m10);
  float* kp = kp_all + ((size_t)slot * GT_MAX_KP + kpi) * 6;
  if (lane == 0) {
    kp[0] = __fmul_rn((float)x, L.scale);
    kp[1] = __fmul_rn((float)y, L.scale);
    kp[2] = __fmul_rn(31.0f, L.scale);
    kp[3] = angle;
    kp[4] = sel_resp[((size_t)slot * GT_ORB_LEVELS + level) * kSelCap + li];
    kp[5] = (float)level;
  }
  // horizontal pass: 90 work items = (row, half); half 0 -> output columns 0..19 from source bytes 0..25 (words 0..6),
  // half 1 -> output columns 20..38 from source bytes 20..44 (words 5..11)
  const float g0 = c_gauss[0], g1 = c_gauss[1], g2 = c_gauss[2], g3 = c_gauss[3];   // symmetric kernel: g4 = g2, g5 = g1, g6 = g0
#pragma unroll 1
  for (int item = lane; item < 2 * kSrcW; item += 32) {
    const int hf = item >= kSrcW ? 1 : 0, r = item - hf * kSrcW;
    const uint32_t* rw = s_srcw + r * kSrcPitchW + hf * 5;
    uint32_t wv[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) wv[k] = rw[k];
    float f[28];
#pragma unroll
    for (int j = 0; j < 28; ++j) f[j] = (float)((wv[j >> 2] >> (8 * (j & 3))) & 0xFFu);
    float* out = s_hp + r * kBlurW + hf * 20;
    const int nout = hf ? 19 : 20;
#pragma unroll
    for (int o = 0; o < 20; ++o) {
      float acc = __fmul_rn(f[o], g0);
      acc = __fadd_rn(acc, __fmul_rn(f[o + 1], g1));
      acc = __fadd_rn(acc, __fmul_rn(f[o + 2], g2));
      acc = __fadd_rn(acc, __fmul_rn(f[o + 3], g3));
      acc = __fadd_rn(acc, __fmul_rn(f[o + 4], g2));
      acc = __fadd_rn(acc, __fmul_rn(f[o + 5], g1));
      acc = __fadd_rn(acc, __fmul_rn(f[o + 6], g0));
      if (o < nout) out[o] = acc;
    }
  }
  __syncwarp();
  // descriptor: lane computes byte `lane` (bits 8*lane .. 8*lane+7); the vertical blur pass is evaluated at the test points only
  const float ar = __fmul_rn(angle, 0.017453292519943295f);  // (float)(CV_PI/180)
  const float ca = (float)cos((double)ar), sa = (float)sin((double)ar);
  auto blurred = [&](int ix, int iy) -> int {                 // blurred pixel (x + ix, y + iy), rounded to u8 like the blurred image
    const float* c = s_hp + (kBlurR + iy) * kBlurW + (kBlurR + ix);
    float acc = __fmul_rn(c[0], g0);
    acc = __fadd_rn(acc, __fmul_rn(c[kBlurW], g1));
    acc = __fadd_rn(acc, __fmul_rn(c[2 * kBlurW], g2));
    acc = __fadd_rn(acc, __fmul_rn(c[3 * kBlurW], g3));
    acc = __fadd_rn(acc, __fmul_rn(c[4 * kBlurW], g2));
    acc = __fadd_rn(acc, __fmul_rn(c[5 * kBlurW], g1));
    acc = __fadd_rn(acc, __fmul_rn(c[6 * kBlurW], g0));
    return min(max(__float2int_rn(acc), 0), 255);
  };
  unsigned byte = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const signed char* p = s_pat[lane * 8 + j];
    const float px = (float)p[0], py = (float)p[1], qx = (float)p[2], qy = (float)p[3];
    const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(px, ca), __fmul_rn(py, sa)));
    const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(px, sa), __fmul_rn(py, ca)));
    const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(qx, ca), __fmul_rn(qy, sa)));
    const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(qx, sa), __fmul_rn(qy, ca)));
    byte |= (unsigned)(blurred(ix0, iy0) < blurred(ix1, iy1)) << j;
  }
  desc_all[((size_t)slot * GT_MAX_KP + kpi) * 32 + lane] = (uint8_t)byte;
}

void resize_tables(int n_src, int n_dst, std::vector<int>& ofs, std::vector<int>& c1) {
  ofs.resize(n_dst);
  c1.resize(n_dst);
  const double scale = (double)n_src / n_dst;
  for (int d = 0; d < n_dst; ++d) {
    double f = (d + 0.5) * scale - 0.5;
    int s = (int)std::floor(f);
    double fr = f - s;
    if (s < 0) { s = 0; fr = 0; }
    if (s >= n_src - 1) { s = n_src - 1; fr = 0; }
    ofs[d] = s;
    c1[d] = (int)std::nearbyint(fr * 256.0);
  }
}

}  // namespace

int orb_build(gt_engine* e) {
  const int B = e->cfg.max_batch, S = B + 1;
  const int nfeat_cur = e->cfg.max_features;
  const int nfeat_ref = (int)(e->cfg.max_features * e->cfg.ref_multiplier);
  GT_CHECK(e, nfeat_ref <= GT_MAX_KP && nfeat_cur >= 8, "orb: max_features * ref_multiplier = %d exceeds %d", nfeat_ref, GT_MAX_KP);
  auto quotas = [](int n, int* out) {
    const float factor = (float)(1.0 / 1.2);
    float nd = n * (1 - factor) / (1 - (float)std::pow((double)factor, (double)GT_ORB_LEVELS));
    int sum = 0;
    for (int l = 0; l < GT_ORB_LEVELS - 1; ++l) {
      out[l] = (int)std::nearbyint(nd);
      sum += out[l];
      nd *= factor;
    }
    out[GT_ORB_LEVELS - 1] = std::max(n - sum, 0);
  };
  int qc[GT_ORB_LEVELS], qr[GT_ORB_LEVELS];
  quotas(nfeat_cur, qc);
  quotas(nfeat_ref, qr);
  size_t off = 0, coff = 0;
  for (int l = 0; l < GT_ORB_LEVELS; ++l) {
    OrbLevel& L = e->lv[l];
    L.scale = (float)std::pow(1.2, (double)l);
    L.w = (int)std::nearbyint((float)e->work_w / L.scale);
    L.h = (int)std::nearbyint((float)e->work_h / L.scale);
    L.off = off;
    off += (((size_t)L.w * L.h) + 255) & ~(size_t)255;
    L.quota_cur = qc[l]; L.quota_ref = qr[l];
    GT_CHECK(e, qr[l] <= kLvlKeep, "orb: per-level quota %d exceeds %d", qr[l], kLvlKeep);
    L.cand_cap = std::max(4096, ((L.w * L.h / 16 + 1023) / 1024) * 1024);
    L.cand_off = coff;
    coff += L.cand_cap;
  }
  e->pyr_bytes = off;
  e->cand_total = coff;
  e->sel_cap = kSelCap;
  GT_TRY(e->dev_alloc((void**)&e->pyr, (size_t)S * off));
  GT_TRY(e->dev_alloc((void**)&e->pyr_mask, (size_t)S * off));
  GT_TRY(e->dev_alloc((void**)&e->fast_cand, (size_t)S * coff * sizeof(unsigned)));
  GT_TRY(e->dev_alloc((void**)&e->fast_score, (size_t)S * coff));
  GT_TRY(e->dev_alloc((void**)&e->fast_count, (size_t)S * GT_ORB_LEVELS * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->sel_xy, (size_t)S * GT_ORB_LEVELS * kSelCap * sizeof(unsigned)));
  GT_TRY(e->dev_alloc((void**)&e->sel_resp, (size_t)S * GT_ORB_LEVELS * kSelCap * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->sel_count, (size_t)S * GT_ORB_LEVELS * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->kp_all, (size_t)S * GT_MAX_KP * 6 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->desc_all, (size_t)S * GT_MAX_KP * 32));
  GT_TRY(e->dev_alloc((void**)&e->kp_count, (size_t)S * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->lvl_kp_off, (size_t)S * (GT_ORB_LEVELS + 1) * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->boxes_dev, (size_t)S * e->cfg.max_det * 4 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->nboxes_dev, (size_t)S * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->det_xywh_dev, (size_t)S * e->cfg.max_det * 4 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->det_nbox_dev, (size_t)S * sizeof(int)));
  GT_CUDA(e, cudaMemset(e->nboxes_dev, 0, (size_t)S * sizeof(int)));
  GT_CUDA(e, cudaMemset(e->kp_count, 0, (size_t)S * sizeof(int)));
  GT_CUDA(e, cudaMemset(e->pyr_mask, 255, (size_t)S * off));
  GT_TRY(e->dev_alloc((void**)&e->lv_dev, sizeof(OrbLevel) * GT_ORB_LEVELS));
  GT_CUDA(e, cudaMemcpy(e->lv_dev, e->lv, sizeof(OrbLevel) * GT_ORB_LEVELS, cudaMemcpyHostToDevice));
  for (int l = 1; l < GT_ORB_LEVELS; ++l) {
    std::vector<int> xo, xc, yo, yc;
    resize_tables(e->lv[l - 1].w, e->lv[l].w, xo, xc);
    resize_tables(e->lv[l - 1].h, e->lv[l].h, yo, yc);
    while (xo.size() % 4) { xo.push_back(xo.back()); xc.push_back(xc.back()); }   // the resize kernel loads four x entries at once
    const std::vector<int>* src[4] = {&xo, &xc, &yo, &yc};
    for (int k = 0; k < 4; ++k) {
      GT_TRY(e->dev_alloc((void**)&e->rs_tab[l][k], src[k]->size() * 4));
      GT_CUDA(e, cudaMemcpy(e->rs_tab[l][k], src[k]->data(), src[k]->size() * 4, cudaMemcpyHostToDevice));
    }
  }
  GT_CUDA(e, cudaFuncSetAttribute(orb_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelCap * 4));
  GT_CUDA(e, cudaFuncSetAttribute(orb_describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDescWarps * ((kDescSmemPerWarp + 15) & ~15)));
  return GT_OK;
}

// levels 1..7 of one plane (image or mask) for slots [slot0, slot0 + nslots): seven chained launches.  (A single-launch form -- a block owns
// a tile of the last level and the halo regions of every level below it in shared memory -- was bit-exact but slower, 255 vs 160 us per
// plane: 1.4x the pixels with byte-granular shared-memory traffic; profiles/round2_summary.md section 2.)
template <bool MASK>
static int pyramid_run(gt_engine* e, uint8_t* plane, int slot0, int nslots, cudaStream_t st) {
  const size_t slab = e->pyr_bytes;
  for (int l = 1; l < GT_ORB_LEVELS; ++l) {
    const OrbLevel& S = e->lv[l - 1];
    const OrbLevel& D = e->lv[l];
    int* const* t = e->rs_tab[l];
    dim3 g((unsigned)ceil_div(D.w, 4 * 128), (unsigned)ceil_div(D.h, kResizeRows), (unsigned)nslots);
    pyr_resize_kernel<MASK><<<g, 128, 0, st>>>(plane, slab, slot0, S.off, S.w, S.h, D.off, D.w, D.h, t[0], t[1], t[2], t[3]);
    e->launches++;
  }
  return GT_OK;
}

// Mask-independent half of ORB: image pyramid, blurred pyramid, FAST candidates.  Needs only the gray level 0, so it can run on
// a second stream while the detector works on the same frames.
int orb_pyramid(gt_engine* e, int slot0, int nslots, cudaStream_t st) {   // image pyramid only (needs the gray level 0)
  GT_TRY(pyramid_run<false>(e, e->pyr, slot0, nslots, st));
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

// FAST candidates of pyramid levels [level_lo, level_hi) (needs those levels of the image pyramid); `reset` clears the per-level counters
int orb_fast(gt_engine* e, int slot0, int nslots, cudaStream_t st, int level_lo, int level_hi, bool reset) {
  const size_t slab = e->pyr_bytes;
  if (reset) GT_CUDA(e, cudaMemsetAsync(e->fast_count + (size_t)slot0 * GT_ORB_LEVELS, 0, (size_t)nslots * GT_ORB_LEVELS * sizeof(int), st));
  FastLevels fl;
  int tiles = 0;
  for (int l = 0; l < GT_ORB_LEVELS; ++l) {
    const OrbLevel& L = e->lv[l];
    const int tx = std::max(0, ceil_div(L.w - 2 * kEdge, FT_X)), ty = std::max(0, ceil_div(L.h - 2 * kEdge, FT_Y));
    fl.w[l] = L.w; fl.h[l] = L.h; fl.tiles_x[l] = std::max(tx, 1); fl.tile0[l] = tiles; fl.cand_cap[l] = L.cand_cap;
    fl.off[l] = L.off; fl.cand_off[l] = L.cand_off;
    tiles += tx * ty;
  }
  fl.tile0[GT_ORB_LEVELS] = tiles;
  const int first = fl.tile0[level_lo], n = fl.tile0[level_hi] - first;
  if (n > 0) {
    fast_kernel<<<dim3((unsigned)n, (unsigned)nslots), 256, 0, st>>>(e->pyr, slab, slot0, fl, e->fast_cand, e->fast_score, e->cand_total, e->fast_count, first);
    e->launches++;
  }
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

int orb_front(gt_engine* e, int slot0, int nslots, cudaStream_t st) {
  GT_TRY(orb_pyramid(e, slot0, nslots, st));
  return orb_fast(e, slot0, nslots, st, 0, GT_ORB_LEVELS, true);
}

// The same on two streams: FAST on level 0 (53 % of the pixels; needs only the gray frame) runs on `st` WHILE the seven latency-bound
// pyramid launches run on `st2`; FAST on levels 1..7 follows on `st` once both are done.  ev_a / ev_b order the two streams.
int orb_front_split(gt_engine* e, int slot0, int nslots, cudaStream_t st, cudaStream_t st2, cudaEvent_t ev_a, cudaEvent_t ev_b) {
  GT_CUDA(e, cudaEventRecord(ev_a, st));
  GT_CUDA(e, cudaStreamWaitEvent(st2, ev_a, 0));
  GT_TRY(orb_pyramid(e, slot0, nslots, st2));
  GT_CUDA(e, cudaEventRecord(ev_b, st2));
  GT_TRY(orb_fast(e, slot0, nslots, st, 0, 1, true));
  GT_CUDA(e, cudaStreamWaitEvent(st, ev_b, 0));
  return orb_fast(e, slot0, nslots, st, 1, GT_ORB_LEVELS, false);
}

// Mask-dependent half, part 1: vehicle mask (level 0 from the boxes unless the caller supplied one) + its pyramid.  Needs the boxes
// (detections / tracker boxes) but NOT the image pyramid or FAST: the caller runs it before waiting for orb_front, so it overlaps the
// tail of the FAST kernel on the aux stream.
int orb_mask(gt_engine* e, int slot0, int nslots, bool build_mask, cudaStream_t st) {
  const size_t slab = e->pyr_bytes;
  const OrbLevel& L0 = e->lv[0];
  const dim3 gbox((unsigned)std::min(e->cfg.max_det, 160), (unsigned)nslots);   // blocks walk the boxes
  if (build_mask && e->mask_sparse) {
    // box-derived masks: every level preset to 255, rectangles drawn at level 0, then levels 1..7 recomputed around the boxes only
    GT_CUDA(e, cudaMemsetAsync(e->pyr_mask + (size_t)slot0 * slab, 255, (size_t)nslots * slab, st));
    if (e->cfg.mask_use) {
      mask_rects_kernel<<<gbox, 128, 0, st>>>(e->pyr_mask, slab, e->boxes_dev, e->nboxes_dev, e->cfg.max_det, slot0, L0.w, L0.h,
                                              e->cfg.mask_margin_ratio, e->cfg.downsample_ratio);
      e->launches++;
      MaskLevels ml;
      for (int l = 0; l < GT_ORB_LEVELS; ++l) {
        ml.w[l] = e->lv[l].w; ml.h[l] = e->lv[l].h; ml.off[l] = e->lv[l].off;
        ml.sx[l] = l ? (double)e->lv[l - 1].w / e->lv[l].w : 1.0; ml.sy[l] = l ? (double)e->lv[l - 1].h / e->lv[l].h : 1.0;
      }
      for (int l = 1; l < GT_ORB_LEVELS; ++l) {
        int* const* t = e->rs_tab[l];
        mask_pyr_sparse_kernel<<<gbox, 128, 0, st>>>(e->pyr_mask, slab, e->boxes_dev, e->nboxes_dev, e->cfg.max_det, slot0, e->cfg.mask_margin_ratio,
                                                     e->cfg.downsample_ratio, ml, l, t[0], t[1], t[2], t[3]);
        e->launches++;
      }
    }
    GT_CUDA(e, cudaGetLastError());
    return GT_OK;
  }
  if (build_mask) {
    GT_CUDA(e, cudaMemset2DAsync(e->pyr_mask + (size_t)slot0 * slab, slab, 255, (size_t)L0.w * L0.h, nslots, st));
    if (e->cfg.mask_use) {
      mask_rects_kernel<<<gbox, 128, 0, st>>>(e->pyr_mask, slab, e->boxes_dev, e->nboxes_dev, e->cfg.max_det, slot0, L0.w, L0.h,
                                              e->cfg.mask_margin_ratio, e->cfg.downsample_ratio);
      e->launches++;
    }
  }
  GT_TRY(pyramid_run<true>(e, e->pyr_mask, slot0, nslots, st));   // caller-supplied level-0 masks (gt_orb_detect) / GT_MASK_SPARSE=0: the dense chain
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

// Part 2: masked candidate selection (score cut, Harris, quota cut), orientation + descriptors.  Needs orb_front and orb_mask.
int orb_back(gt_engine* e, int slot0, int nslots, bool as_reference, cudaStream_t st) {
  const size_t slab = e->pyr_bytes;
  {
    dim3 g(GT_ORB_LEVELS, (unsigned)nslots);
    orb_select_kernel<<<g, 1024, kSelCap * 4, st>>>(e->pyr, e->pyr_mask, slab, slot0, e->lv_dev, e->fast_cand, e->fast_score, e->cand_total,
                                                    e->fast_count, as_reference ? 1 : 0, e->sel_xy, e->sel_resp, e->sel_count);
    e->launches++;
  }
  {
    const int nkp = std::min(GT_MAX_KP, (as_reference ? (int)(e->cfg.max_features * e->cfg.ref_multiplier) : e->cfg.max_features) + GT_ORB_LEVELS * 64);
    dim3 g((unsigned)ceil_div(nkp, kDescWarps), (unsigned)nslots);  // one warp per key point; warps past the count exit
    const size_t dsm = (size_t)kDescWarps * ((kDescSmemPerWarp + 15) & ~15);
    orb_describe_kernel<<<g, kDescWarps * 32, dsm, st>>>(e->pyr, slab, slot0, e->lv_dev, e->sel_xy, e->sel_resp, e->sel_count, e->kp_all,
                                                         e->desc_all, e->kp_count, e->lvl_kp_off);
    e->launches++;
  }
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

int orb_run(gt_engine* e, int slot0, int nslots, bool as_reference, bool build_mask, cudaStream_t st) {
  GT_TRY(orb_front(e, slot0, nslots, st));
  GT_TRY(orb_mask(e, slot0, nslots, build_mask, st));
  return orb_back(e, slot0, nslots, as_reference, st);
}
