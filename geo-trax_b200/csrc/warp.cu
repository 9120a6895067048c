// Frame warp into the stabilised (reference) frame: cv2.warpPerspective(frame, H, (w, h)) as the reference's visualisation applies it to
// every frame with its transform (/root/reference/geotrax/visualize.py:285-289).  Bit-exact restatement of
// OpenCV's INTER_LINEAR / BORDER_CONSTANT(0) path for 8-bit images (pinned on the CPU: oracle/prepost.py:warp_perspective_u8):
//   * H is inverted with OpenCV's closed-form 3x3 inverse (double), on the host;
//   * the destination is walked in 64-pixel-wide blocks; per block row X0 = M0*bx + M1*y + M2 (same for Y0, W0), per pixel
//     X = round((X0 + M0*x1) * (32 / (W0 + M6*x1))): 1/32-pixel fixed point, every double operation rounded on its own (no FMA
//     contraction: __dmul_rn / __dadd_rn / __ddiv_rn), round-half-even;
//   * the 2 x 2 source pixels are blended with the 15-bit weights (32-ay)(32-ax)*32 ..., result = (sum + 2^14) >> 15; pixels outside
//     the source contribute 0.
// HBM-bound by design: 3 B written + <= 12 B read (mostly L1 / L2 hits) per destination pixel; one thread = one pixel, 3 channels.
#include "engine.cuh"

namespace {

__global__ void __launch_bounds__(256) warp_perspective_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const double* __restrict__ Minv,
                                                               int B, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int b = (int)(idx / ((long long)W * H));
  const double* M = Minv + (size_t)b * 9;
  const double bx = (double)((x >> 6) << 6), x1 = (double)(x & 63), yd = (double)y;
  const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], bx), __dmul_rn(M[1], yd)), M[2]);
  const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], bx), __dmul_rn(M[4], yd)), M[5]);
  const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], bx), __dmul_rn(M[7], yd)), M[8]);
  double w = __dadd_rn(W0, __dmul_rn(M[6], x1));
  w = w != 0.0 ? __ddiv_rn(32.0, w) : 0.0;
  const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], x1)), w)));
  const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], x1)), w)));
  const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
  const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
  const int ax = X & 31, ay = Y & 31;
  const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
  const uint8_t* f = src + (size_t)b * H * W * 3;
  int acc[3] = {0, 0, 0};
  auto add = [&](int yy, int xx, int wt) {
    if (wt == 0 || xx < 0 || xx >= W || yy < 0 || yy >= H) return;
    const uint8_t* p = f + ((size_t)yy * W + xx) * 3;
    acc[0] += (int)p[0] * wt; acc[1] += (int)p[1] * wt; acc[2] += (int)p[2] * wt;
  };
  add(sy, sx, w00); add(sy, sx + 1, w01); add(sy + 1, sx, w10); add(sy + 1, sx + 1, w11);
  uint8_t* o = dst + (size_t)idx * 3;
  o[0] = (uint8_t)min(max((acc[0] + (1 << 14)) >> 15, 0), 255);
  o[1] = (uint8_t)min(max((acc[1] + (1 << 14)) >> 15, 0), 255);
  o[2] = (uint8_t)min(max((acc[2] + (1 << 14)) >> 15, 0), 255);
}

// OpenCV's closed-form inverse of a 3 x 3 double matrix (core/src/lapack.cpp, cv::invert, the Matx33 branch)
bool invert3x3(const double* S, double* t) {
  const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (det == 0.0) return false;
  const double d = 1. / det;
  t[0] = (S[4] * S[8] - S[5] * S[7]) * d;
  t[1] = (S[2] * S[7] - S[1] * S[8]) * d;
  t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  t[3] = (S[5] * S[6] - S[3] * S[8]) * d;
  t[4] = (S[0] * S[8] - S[2] * S[6]) * d;
  t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  t[6] = (S[3] * S[7] - S[4] * S[6]) * d;
  t[7] = (S[1] * S[6] - S[0] * S[7]) * d;
  t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  return true;
}

}  // namespace

int warp_frames_run(gt_engine* e, const uint8_t* src_dev, uint8_t* dst_dev, const double* H_host, int B, cudaStream_t st) {
  std::vector<double> inv((size_t)B * 9);
  for (int b = 0; b < B; ++b)
    GT_CHECK(e, invert3x3(H_host + (size_t)b * 9, inv.data() + (size_t)b * 9), "gt_warp_frames: transform %d is singular", b);
  if (!e->warp_minv) GT_TRY(e->dev_alloc((void**)&e->warp_minv, (size_t)e->cfg.max_batch * 9 * sizeof(double)));
  GT_CUDA(e, cudaMemcpyAsync(e->warp_minv, inv.data(), inv.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  GT_CUDA(e, cudaStreamSynchronize(st));   // `inv` is a stack-owned vector
  const long long n = (long long)B * e->cfg.frame_h * e->cfg.frame_w;
  warp_perspective_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src_dev, dst_dev, e->warp_minv, B, e->cfg.frame_h, e->cfg.frame_w);
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}
