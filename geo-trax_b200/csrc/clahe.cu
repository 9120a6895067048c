// CLAHE front end of the stabilizer (the reference's `stable` preset: /root/reference/geotrax/cfg/stable.yaml:115 `clahe: true`;
// stabilo applies cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)) to the gray frame before the working-resolution resize).
//
// Bit-exact restatement of OpenCV's CLAHE for 8-bit images (imgproc/src/clahe.cpp; pinned on the CPU by oracle/prepost.py:clahe_u8
// against cv2 in tests/test_oracle_model.py): per-tile 256-bin histogram (tiles of the image extended to a multiple of the grid with
// BORDER_REFLECT_101), clip at max(1, int(clip * area / 256)), redistribution (batch + strided residual), LUT = rint(cumsum * 255 /
// area), then per pixel the bilinear blend of the four neighbouring tile LUTs in float32 with every product and sum rounded.
//
// Launches per batch: BGR -> gray plane (full resolution), tile histograms, LUTs, apply (+ the working-image resize when
// downsample_ratio < 1).  HBM-bound: 3 B read + 1 B written per source pixel for the gray plane, then 1 + 1 (+ 1) B for the rest.
#include <algorithm>
#include <cmath>
#include <vector>

#include "engine.cuh"

namespace {

constexpr int kTilesX = 8, kTilesY = 8, kTiles = kTilesX * kTilesY;

__device__ __forceinline__ uint32_t gray15c(uint32_t b, uint32_t g, uint32_t r) { return (9798u * r + 19235u * g + 3735u * b + 16384u) >> 15; }

// one thread = 16 pixels of one row: three 128-bit loads -> one 128-bit store (rows are 16-pixel aligned when W % 16 == 0, else scalar tail)
__global__ void __launch_bounds__(256) gray_full_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ gray, int B, int H, int W) {
  const int groups = (W + 15) >> 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * groups) return;
  const int g = (int)(idx % groups);
  const long long row = idx / groups;   // b * H + y
  const uint8_t* src = frames + (size_t)row * W * 3 + (size_t)g * 48;
  uint8_t* dst = gray + (size_t)row * W + (size_t)g * 16;
  if ((W & 15) == 0) {
    uint4 v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const uint8_t* p = reinterpret_cast<const uint8_t*>(v);
    __align__(16) uint8_t o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = (uint8_t)gray15c(p[3 * j], p[3 * j + 1], p[3 * j + 2]);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
  } else {
    const int n = min(16, W - g * 16);
    for (int j = 0; j < n; ++j) dst[j] = (uint8_t)gray15c(src[3 * j], src[3 * j + 1], src[3 * j + 2]);
  }
}

// grid (chunks, tiles, B): a block accumulates `rows_per_block` rows of one tile in shared memory, then merges into the tile histogram
__global__ void __launch_bounds__(256) clahe_hist_kernel(const uint8_t* __restrict__ gray, int H, int W, int tw, int th, int rows_per_block,
                                                         int* __restrict__ hist) {
  __shared__ int s_h[256];
  const int tile = blockIdx.y, b = blockIdx.z;
  const int tx = tile % kTilesX, ty = tile / kTilesX;
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const uint8_t* g = gray + (size_t)b * H * W;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, th);
  for (int r = r0; r < r1; ++r) {
    int y = ty * th + r;
    if (y >= H) y = 2 * H - 2 - y;                      // BORDER_REFLECT_101 extension (bottom)
    for (int c = threadIdx.x; c < tw; c += 256) {
      int x = tx * tw + c;
      if (x >= W) x = 2 * W - 2 - x;                    // (right)
      atomicAdd(&s_h[g[(size_t)y * W + x]], 1);
    }
  }
  __syncthreads();
  const int v = s_h[threadIdx.x];
  if (v) atomicAdd(&hist[((size_t)b * kTiles + tile) * 256 + threadIdx.x], v);
}

// grid (tiles, B), 256 threads = 256 bins
__global__ void __launch_bounds__(256) clahe_lut_kernel(const int* __restrict__ hist, int clip_limit, float lut_scale, uint8_t* __restrict__ lut) {
  __shared__ int s_scan[256];
  __shared__ int s_warp[8];
  const int i = threadIdx.x, lane = i & 31, warp = i >> 5;
  const size_t base = ((size_t)blockIdx.y * kTiles + blockIdx.x) * 256;
  int h = hist[base + i];
  if (clip_limit > 0) {
    int over = max(h - clip_limit, 0);
    h = min(h, clip_limit);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) over += __shfl_xor_sync(0xffffffffu, over, o);
    if (lane == 0) s_warp[warp] = over;
    __syncthreads();
    int clipped = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) clipped += s_warp[w];
    const int batch = clipped / 256;
    int resid = clipped - batch * 256;
    h += batch;
    if (resid != 0) {
      const int step = max(256 / resid, 1);
      if ((i % step) == 0 && i / step < resid) ++h;      // bins 0, step, 2 step, ... (resid of them, all below 256)
    }
    __syncthreads();
  }
  // inclusive prefix sum over the 256 bins
  s_scan[i] = h;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const int v = i >= o ? s_scan[i - o] : 0;
    __syncthreads();
    s_scan[i] += v;
    __syncthreads();
  }
  const float f = __fmul_rn((float)s_scan[i], lut_scale);
  lut[base + i] = (uint8_t)min(max(__float2int_rn(f), 0), 255);
}

struct ClaheAxis { const int *i1, *i2; const float *a, *a1; };

// one thread = 4 adjacent pixels; out_stride / out_frame_stride let the result land directly in level 0 of the pyramid slabs
__global__ void __launch_bounds__(256) clahe_apply_kernel(const uint8_t* __restrict__ gray, const uint8_t* __restrict__ lut, int B, int H, int W,
                                                          ClaheAxis ax, ClaheAxis ay, uint8_t* __restrict__ out, size_t out_frame_stride) {
  const int groups = (W + 3) >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * groups) return;
  const int g = (int)(idx % groups);
  const long long row = idx / groups;
  const int b = (int)(row / H), y = (int)(row - (long long)b * H);
  const uint8_t* src = gray + (size_t)row * W;
  const uint8_t* L = lut + (size_t)b * kTiles * 256;
  const uint8_t* l1 = L + (size_t)ay.i1[y] * kTilesX * 256;
  const uint8_t* l2 = L + (size_t)ay.i2[y] * kTilesX * 256;
  const float ya = ay.a[y], ya1 = ay.a1[y];
  uint8_t* dst = out + (size_t)b * out_frame_stride + (size_t)y * W;
  const int n = min(4, W - g * 4);
  for (int j = 0; j < n; ++j) {
    const int x = g * 4 + j;
    const int v = src[x];
    const int o1 = ax.i1[x] * 256 + v, o2 = ax.i2[x] * 256 + v;
    const float xa = ax.a[x], xa1 = ax.a1[x];
    const float t1 = __fadd_rn(__fmul_rn((float)l1[o1], xa1), __fmul_rn((float)l1[o2], xa));
    const float t2 = __fadd_rn(__fmul_rn((float)l2[o1], xa1), __fmul_rn((float)l2[o2], xa));
    const float r = __fadd_rn(__fmul_rn(t1, ya1), __fmul_rn(t2, ya));
    dst[x] = (uint8_t)min(max(__float2int_rn(r), 0), 255);
  }
}

// working-image resize of a u8 plane: cv2.resize(INTER_LINEAR) integer arithmetic (same tables as gray_work_kernel in detector.cu)
struct PlaneTabs { const int *x0, *x1, *a0, *a1, *y0, *y1, *b0, *b1; int mode; };
__global__ void __launch_bounds__(256) resize_plane_kernel(const uint8_t* __restrict__ src, int B, int H, int W, uint8_t* __restrict__ dst,
                                                           size_t dst_frame_stride, int oh, int ow, PlaneTabs t) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)oh * ow;
  if (idx >= per * B) return;
  const int b = (int)(idx / per);
  const int rem = (int)(idx - (long long)b * per);
  const int oy = rem / ow, ox = rem - oy * ow;
  const uint8_t* f = src + (size_t)b * H * W;
  uint32_t v;
  if (t.mode == 1) {
    const uint8_t* p = f + (size_t)(2 * oy) * W + 2 * ox;
    v = ((uint32_t)p[0] + p[1] + p[W] + p[W + 1] + 2) >> 2;
  } else {
    const int x0 = t.x0[ox], x1 = t.x1[ox], y0 = t.y0[oy], y1 = t.y1[oy];
    const int S0 = t.a0[ox] * (int)f[(size_t)y0 * W + x0] + t.a1[ox] * (int)f[(size_t)y0 * W + x1];
    const int S1 = t.a0[ox] * (int)f[(size_t)y1 * W + x0] + t.a1[ox] * (int)f[(size_t)y1 * W + x1];
    v = (uint32_t)((((t.b0[oy] * (S0 >> 4)) >> 16) + ((t.b1[oy] * (S1 >> 4)) >> 16) + 2) >> 2);
  }
  dst[(size_t)b * dst_frame_stride + (size_t)oy * ow + ox] = (uint8_t)v;
}

}  // namespace

int clahe_build(gt_engine* e) {
  if (!e->cfg.clahe) return GT_OK;
  const int H = e->cfg.frame_h, W = e->cfg.frame_w, B = e->cfg.max_batch;
  const int We = (W % kTilesX) ? W + kTilesX - (W % kTilesX) : W, He = (H % kTilesY) ? H + kTilesY - (H % kTilesY) : H;
  e->clahe_tw = We / kTilesX; e->clahe_th = He / kTilesY;
  GT_CHECK(e, e->clahe_tw >= 2 && e->clahe_th >= 2 && We - W < W && He - H < H, "clahe: frame %dx%d too small for an 8x8 tile grid", W, H);
  const int area = e->clahe_tw * e->clahe_th;
  e->clahe_clip = std::max((int)(2.0 * area / 256), 1);            // clipLimit 2.0 (stabilo's constant)
  e->clahe_scale = 255.0f / (float)area;
  GT_TRY(e->dev_alloc((void**)&e->gray_full, (size_t)B * H * W));
  if (!(e->work_w == W && e->work_h == H)) GT_TRY(e->dev_alloc((void**)&e->gray_eq, (size_t)B * H * W));
  GT_TRY(e->dev_alloc((void**)&e->clahe_hist, (size_t)B * kTiles * 256 * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->clahe_lut, (size_t)B * kTiles * 256));
  auto axis = [&](int n, int tile, int ntiles, int** i1, int** i2, float** a, float** a1) -> int {
    std::vector<int> v1(n), v2(n);
    std::vector<float> va(n), va1(n);
    const float inv = 1.0f / (float)tile;
    for (int x = 0; x < n; ++x) {
      const float tf = (float)x * inv - 0.5f;
      int t1 = (int)floorf(tf);
      const int t2 = t1 + 1;
      va[x] = tf - (float)t1; va1[x] = 1.0f - va[x];
      v1[x] = std::max(t1, 0); v2[x] = std::min(t2, ntiles - 1);
    }
    GT_TRY(e->dev_alloc((void**)i1, n * 4)); GT_TRY(e->dev_alloc((void**)i2, n * 4));
    GT_TRY(e->dev_alloc((void**)a, n * 4)); GT_TRY(e->dev_alloc((void**)a1, n * 4));
    GT_CUDA(e, cudaMemcpy(*i1, v1.data(), n * 4, cudaMemcpyHostToDevice)); GT_CUDA(e, cudaMemcpy(*i2, v2.data(), n * 4, cudaMemcpyHostToDevice));
    GT_CUDA(e, cudaMemcpy(*a, va.data(), n * 4, cudaMemcpyHostToDevice)); GT_CUDA(e, cudaMemcpy(*a1, va1.data(), n * 4, cudaMemcpyHostToDevice));
    return GT_OK;
  };
  GT_TRY(axis(W, e->clahe_tw, kTilesX, &e->clahe_xi[0], &e->clahe_xi[1], &e->clahe_xa[0], &e->clahe_xa[1]));
  GT_TRY(axis(H, e->clahe_th, kTilesY, &e->clahe_yi[0], &e->clahe_yi[1], &e->clahe_ya[0], &e->clahe_ya[1]));
  return GT_OK;
}

// frames_dev: u8 BGR [B][H][W][3] -> CLAHE-equalised working image in level 0 of the pyramid slabs
int clahe_run(gt_engine* e, const uint8_t* frames_dev, int B, cudaStream_t st) {
  const int H = e->cfg.frame_h, W = e->cfg.frame_w;
  const long long n16 = (long long)B * H * ((W + 15) / 16);
  gray_full_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>(frames_dev, e->gray_full, B, H, W);
  GT_CUDA(e, cudaMemsetAsync(e->clahe_hist, 0, (size_t)B * kTiles * 256 * sizeof(int), st));
  const int rows_per_block = 16;
  dim3 gh((unsigned)ceil_div(e->clahe_th, rows_per_block), kTiles, (unsigned)B);
  clahe_hist_kernel<<<gh, 256, 0, st>>>(e->gray_full, H, W, e->clahe_tw, e->clahe_th, rows_per_block, e->clahe_hist);
  clahe_lut_kernel<<<dim3(kTiles, (unsigned)B), 256, 0, st>>>(e->clahe_hist, e->clahe_clip, e->clahe_scale, e->clahe_lut);
  const ClaheAxis ax = {e->clahe_xi[0], e->clahe_xi[1], e->clahe_xa[0], e->clahe_xa[1]};
  const ClaheAxis ay = {e->clahe_yi[0], e->clahe_yi[1], e->clahe_ya[0], e->clahe_ya[1]};
  const long long n4 = (long long)B * H * ((W + 3) / 4);
  const bool direct = e->gray_eq == nullptr;   // working image = full resolution: write level 0 of the pyramid directly
  clahe_apply_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(e->gray_full, e->clahe_lut, B, H, W, ax, ay, direct ? e->pyr : e->gray_eq,
                                                                   direct ? e->pyr_bytes : (size_t)H * W);
  e->launches += 4;
  if (!direct) {
    const PlaneTabs t = {e->gw_tab[0], e->gw_tab[1], e->gw_tab[2], e->gw_tab[3], e->gw_tab[4], e->gw_tab[5], e->gw_tab[6], e->gw_tab[7], e->gw_mode};
    const long long n = (long long)B * e->work_h * e->work_w;
    resize_plane_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->gray_eq, B, H, W, e->pyr, e->pyr_bytes, e->work_h, e->work_w, t);
    e->launches++;
  }
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}
