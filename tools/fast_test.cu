// scratch: isolates the FAST score kernel on one image and compares against the host version of the same code
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
using std::min; using std::max;
#define FT_X 64
#define FT_Y 16
static const int kFastThrH = 20;
__host__ __device__ inline int fast_score(const uint8_t (*t)[FT_X + 8], int x, int y) {
  const int kFastThr = 20;
  const int v = t[y][x];
  int d[16];
  d[0] = v - t[y + 3][x];      d[1] = v - t[y + 3][x + 1];  d[2] = v - t[y + 2][x + 2];  d[3] = v - t[y + 1][x + 3];
  d[4] = v - t[y][x + 3];      d[5] = v - t[y - 1][x + 3];  d[6] = v - t[y - 2][x + 2];  d[7] = v - t[y - 3][x + 1];
  d[8] = v - t[y - 3][x];      d[9] = v - t[y - 3][x - 1];  d[10] = v - t[y - 2][x - 2]; d[11] = v - t[y - 1][x - 3];
  d[12] = v - t[y][x - 3];     d[13] = v - t[y + 1][x - 3]; d[14] = v - t[y + 2][x - 2]; d[15] = v - t[y + 3][x - 1];
#ifndef NO_EARLY
  unsigned dark = 0, bright = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) { dark |= (unsigned)(d[k] > kFastThr) << k; bright |= (unsigned)(d[k] < -kFastThr) << k; }
  auto run9 = [](unsigned m) { m |= m << 16; unsigned a = m & (m >> 1); a = a & (a >> 2); a = a & (a >> 4); a = a & (m >> 8); return (a & 0xFFFFu) != 0; };
  if (!run9(dark) && !run9(bright)) return 0;
#endif
#ifdef USE_NET
  int best = 0;
  {
    int nd[16], a2[16], a4[16], a8[16], b2[16], b4[16], b8[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) nd[k] = -d[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) { a2[k] = min(d[k], d[(k + 1) & 15]); b2[k] = min(nd[k], nd[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { a4[k] = min(a2[k], a2[(k + 2) & 15]); b4[k] = min(b2[k], b2[(k + 2) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { a8[k] = min(a4[k], a4[(k + 4) & 15]); b8[k] = min(b4[k], b4[(k + 4) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { best = max(best, min(a8[k], d[(k + 8) & 15])); best = max(best, min(b8[k], nd[(k + 8) & 15])); }
  }
  return best > kFastThr ? best - 1 : 0;
#else
  int best = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    int mn = d[k], mx = d[k];
#pragma unroll
#ifdef USE_MINMAX
    for (int j = 1; j < 9; ++j) { const int e = d[(k + j) & 15]; mn = min(mn, e); mx = max(mx, e); }
    best = max(best, max(mn, -mx));
#else
    for (int j = 1; j < 9; ++j) { const int e = d[(k + j) & 15]; mn = mn < e ? mn : e; mx = mx > e ? mx : e; }
    const int c = mn > -mx ? mn : -mx;
    best = best > c ? best : c;
#endif
  }
  return best > kFastThr ? best - 1 : 0;
#endif
}
__global__ void score_kernel(const uint8_t* im, int w, int h, uint8_t* out) {
  __shared__ uint8_t s_img[FT_Y + 8][FT_X + 8];
  const int x0 = blockIdx.x * FT_X, y0 = blockIdx.y * FT_Y;
  for (int i = threadIdx.x; i < (FT_Y + 8) * (FT_X + 8); i += 256) {
    const int ty = i / (FT_X + 8), tx = i - ty * (FT_X + 8);
    const int gx = x0 - 4 + tx, gy = y0 - 4 + ty;
    s_img[ty][tx] = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? im[(size_t)gy * w + gx] : 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FT_Y * FT_X; i += 256) {
    const int ly = i / FT_X, lx = i - ly * FT_X;
    const int gx = x0 + lx, gy = y0 + ly;
    if (gx >= w || gy >= h) continue;
    int sc = 0;
    if (gx >= 3 && gx < w - 3 && gy >= 3 && gy < h - 3) sc = fast_score(s_img, lx + 4, ly + 4);
    out[(size_t)gy * w + gx] = (uint8_t)sc;
  }
}
int main() {
  const int w = 960, h = 540;
  std::vector<uint8_t> im(w * h), ref(w * h, 0), got(w * h);
  FILE* f = fopen("gpurun_out/g.bin", "rb"); if (!f) { printf("no g.bin\n"); return 1; } size_t n = fread(im.data(), 1, w * h, f); fclose(f); (void)n;
  static uint8_t tile[FT_Y + 8][FT_X + 8];
  for (int y = 3; y < h - 3; ++y) for (int x = 3; x < w - 3; ++x) {
    for (int ty = 0; ty < 9; ++ty) for (int tx = 0; tx < 9; ++tx) tile[ty][tx] = im[(y - 4 + ty) < 0 || (y - 4 + ty) >= h || (x - 4 + tx) < 0 || (x - 4 + tx) >= w ? 0 : (size_t)(y - 4 + ty) * w + (x - 4 + tx)];
    ref[(size_t)y * w + x] = (uint8_t)fast_score(tile, 4, 4);
  }
  uint8_t *dim, *dout; cudaMalloc(&dim, w * h); cudaMalloc(&dout, w * h);
  cudaMemcpy(dim, im.data(), w * h, cudaMemcpyHostToDevice);
  dim3 g((w + FT_X - 1) / FT_X, (h + FT_Y - 1) / FT_Y);
  score_kernel<<<g, 256>>>(dim, w, h, dout);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(got.data(), dout, w * h, cudaMemcpyDeviceToHost);
  long bad = 0, nz = 0; int shown = 0;
  for (int i = 0; i < w * h; ++i) { nz += ref[i] != 0; if (ref[i] != got[i]) { ++bad; if (shown++ < 5) printf("  (%d,%d) ref %d got %d\n", i % w, i / w, ref[i], got[i]); } }
  printf("%s: err=%s corners(ref)=%ld mismatches=%ld\n", VARIANT, cudaGetErrorString(e), nz, bad);
  return 0;
}
