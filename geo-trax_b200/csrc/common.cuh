// Shared declarations for the geotrax_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/geotrax_b200.h"

typedef __nv_bfloat16 bf16;

#define GT_NUM_SMS 148

// ---- error plumbing: no exceptions across the ABI -----------------------------------------------------------------
struct gt_engine;
void gt_set_error(gt_engine* e, const char* fmt, ...);

#define GT_CUDA(e, call)                                                                          \
  do {                                                                                            \
    cudaError_t _err = (call);                                                                    \
    if (_err != cudaSuccess) {                                                                    \
      gt_set_error((e), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_err));  \
      return GT_ERR_CUDA;                                                                         \
    }                                                                                             \
  } while (0)

#define GT_CHECK(e, cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      gt_set_error((e), __VA_ARGS__); \
      return GT_ERR_INVALID;        \
    }                               \
  } while (0)

#define GT_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != GT_OK) return _rc; \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- NHWC channel-slice view of an activation buffer ----------------------------------------------------------------
struct View {
  bf16* ptr = nullptr;  // base of the full buffer
  int C = 0;            // channels of this slice
  int ctot = 0;         // channels of the full buffer (pixel stride in elements)
  int coff = 0;         // first channel of the slice
  int H = 0, W = 0;
  View slice(int off, int c) const {
    View v = *this;
    v.coff = coff + off;
    v.C = c;
    return v;
  }
};

// ---- tcgen05 implicit-GEMM convolution ------------------------------------------------------------------------------
struct ConvParams {
  int B, H, W;            // output spatial dims
  int tw, th;             // spatial tile (tw*th == 128 GEMM rows)
  int tiles_x, tiles_y;
  int stride, ksize, pad;
  int kc_blocks;          // ceil(cin / 64)
  int num_kb;             // ksize*ksize*kc_blocks
  int BN;                 // GEMM N tile (<= 256, multiple of 16)
  int tmem_cols;          // power of two >= max(32, BN)
  int stages;
  int cout;               // real output channels
  int act;                // 1 = SiLU
  int out_f32;            // 1: out is float (raw head), 0: 16-bit act dtype
  int fp16;               // 16-bit format: 1 = fp16, 0 = bf16
  void* out;
  long long out_img_stride;  // destination pixels per image
  int out_ctot, out_coff;
  const bf16* res;        // optional residual (same spatial dims as the output)
  int res_ctot, res_coff;
  bf16* up;               // optional second, 2x nearest-upsampled destination (2H x 2W)
  int up_ctot, up_coff;
  const float* bias;      // [n_tiles * BN]
};

struct ConvOp {
  CUtensorMap tmA, tmB;
  ConvParams p;
  dim3 grid;
  size_t smem = 0;
  bf16* w_dev = nullptr;     // [cout_pad][taps][cin_pad]
  float* b_dev = nullptr;    // [cout_pad]
  int cin = 0, cout = 0, cout_pad = 0, cin_pad = 0, k = 1, stride = 1;
  int n_src = 0;             // 1..3 canonical convs fused along cout
  int src[3] = {0, 0, 0};    // canonical conv indices
  double flops = 0;
};

int conv_tc_init(gt_engine* e);  // resolves cuTensorMapEncodeTiled, sets kernel attributes
// builds tensor maps + launch geometry; in/out views may be channel slices.  out_f32_ptr != null -> fp32 raw-head store
int conv_tc_plan(gt_engine* e, ConvOp* op, const View& in, int Bmax, int cin, int cout_total, int k, int stride, int act,
                 const View* out, float* out_f32_ptr, long long out_img_stride, int out_ctot_f32, int out_coff_f32,
                 const View* res, const View* up);
int conv_tc_pack_weights(gt_engine* e, ConvOp* op, const float* const* w, const float* const* b, const int* couts, int n);
int conv_tc_launch(gt_engine* e, const ConvOp* op, int B, cudaStream_t st);

// ---- small helpers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// 16-bit activation formats: storage type is always a 2-byte word (typedef bf16 in signatures); `fp16` picks the encoding
__device__ __forceinline__ uint32_t pack2_act(float a, float b, int fp16) {
  if (fp16) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2_act(uint32_t u, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}
static inline uint16_t host_to_act(float f, int fp16) {
  if (fp16) { __half h = __float2half_rn(f); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  return *reinterpret_cast<uint16_t*>(&h);
}
