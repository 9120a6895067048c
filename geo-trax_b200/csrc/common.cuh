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
  int B, H, W;            // images in this launch, output spatial dims
  int img0;               // first image of this launch (tiles address images img0 .. img0 + B - 1)
  int tw, th;             // spatial tile (tw*th == 128 GEMM rows)
  int tiles_x, tiles_y;
  int stride, ksize, pad;
  int kb_elems;           // channels per k-block: 64 (SWIZZLE_128B rows) or 16 (SWIZZLE_32B rows)
  int kc_blocks;          // ceil(cin / kb_elems)
  int num_kb;             // ksize*ksize*kc_blocks
  int BN;                 // GEMM N tile (<= 256, multiple of 16)
  int n_tiles;            // N tiles (cout_pad / BN)
  int total_tiles;        // tiles_x * tiles_y * B * n_tiles
  int tmem_cols;          // power of two >= max(32, 2*BN): two accumulators
  int acc_stride;         // TMEM column offset of the second accumulator
  int stages;
  int halo;               // 0: one TMA box per (tap, k-block); 1: one halo box per k-block, taps are shifted smem descriptors
  int a_stages;           // halo ring depth
  int halo_bytes;         // halo ring stage stride (1024-multiple); the box itself is 16 x (th + k - 1) rows of one k-block
  int halo_tx;            // bytes one halo box delivers
  int b_resident;         // 1: all weight k-blocks stay in shared memory for the CTA's lifetime (ring holds A only)
  int epi_mode;           // 0: 16-bit 128-B slabs, 1: 16-bit 64-B slab, 2: f32 128-B slabs, 3: f32 64-B slab
  int cout;               // real output channels
  int act;                // 0 linear, 1 SiLU (ex2 + rcp: 2 MUFU), 2 SiLU in the one-MUFU tanh form (MUFU-bound layers, see tc_ptx.cuh)
  int out_f32;            // 1: out is float (raw head), 0: 16-bit act dtype
  int fp16;               // 16-bit format: 1 = fp16, 0 = bf16
  float scale;            // accumulator scale applied before the bias (1/255 for layer 0, else 1)
  int out_s2d;            // layer 0: the 128 GEMM outputs of a row are a 2x2 block of output pixels x 32 channels; each 64-channel
                          // slab (one row parity) goes out through the tensor map {64 = (col parity, c), row parity, X, n*rows + Y}
  int out_rows;           // rows per image of that map
  void* out;
  long long out_img_stride;  // destination pixels per image
  int out_ctot, out_coff;
  const bf16* res;        // optional residual (same spatial dims as the output), added AFTER the activation; or, with res_pre:
  int res_ctot, res_coff;
  int res_pre;            // 1: `res` is a HALF-resolution pre-activation partial sum (H/2 x W/2): out = act(acc + bias + res[y/2][x/2]) -- the
                          // upsampled branch of a 1x1 conv over cat(up2x(a), b), computed at a's resolution (1x1 convs commute with nearest upsampling)
  bf16* up;               // optional second, 2x nearest-upsampled destination (2H x 2W)
  int up_ctot, up_coff;
  const float* bias;      // [n_tiles * BN]
  // chained 1x1 conv (conv_tc.cu, pixel-major 16-bit slabs only): the activated tile of this conv never goes to HBM -- its staging slabs
  // are the A operand of a second GEMM against resident weights [chain_n][BN], whose activated result is what gets stored
  int chain_n;            // output channels of the chained conv (64 or 128; 0 = no chain); its K is this conv's BN
  int chain_act;          // its activation (same codes as `act`)
  int chain_tmem;         // TMEM column offset of its accumulator (after the two accumulators of the main conv)
  const float* chain_bias;   // [chain_n]
};

struct ConvUpMaps { CUtensorMap m[4]; };   // the four (dy, dx) phases of the 2x nearest-upsampled destination

struct ConvOp {
  CUtensorMap tmA, tmB, tmOut, tmRes, tmC;   // tmC: weights of the chained 1x1 conv (ConvParams::chain_n)
  ConvUpMaps tmUp;
  ConvParams p;
  size_t smem = 0;
  bf16* w_dev = nullptr;     // [cout_pad][taps][cin_pad]
  float* b_dev = nullptr;    // [cout_pad]
  bf16* w2_dev = nullptr;    // chained conv: [chain_n][BN]
  float* b2_dev = nullptr;   // [chain_n]
  int chain_src = -1;        // canonical conv index of the chained 1x1 conv (-1: none)
  int cin = 0, cout = 0, cout_pad = 0, cin_pad = 0, k = 1, stride = 1;
  int occ2 = 0;              // 1: conv_tc.cu compiled for two CTAs per SM (variant 3)
  int pair = 0;              // 1: conv_sw.cu as CTA pairs (cta_group::2, 256 channels x 256 pixels per pair tile; variant 6)
  int swapped = 0;           // 1: conv_sw.cu (weights = A operand, 256-pixel tile = B operand); tmA = activations, tmB = weights either way
  int n_src = 0;             // 1..3 canonical convs fused along cout
  int w_cin_total = 0, w_cin_off = 0;   // the op uses input channels [w_cin_off, w_cin_off + cin) of a canonical conv with w_cin_total of them (0: all)
  int no_bias = 0;           // 1: bias is zero (partial-sum op; the consumer adds the conv's bias)
  int src[3] = {0, 0, 0};    // canonical conv indices
  double flops = 0;          // per image
  double bytes = 0;          // algorithmic HBM bytes per image (unfused: input + output (+ residual))
};

struct ConvPlanArgs {
  View in;
  int Bmax = 1;
  int cin = 0, cout = 0, k = 1, stride = 1, act = 1;
  int pad = -1;              // -1: k / 2
  int Ho = -1, Wo = -1;      // -1: derived from the input dims
  int kb_elems = 64;
  float scale = 1.0f;
  bool out_s2d = false;        // see ConvParams::out_s2d (out then is the full-resolution NHWC tensor, 2Ho x 2Wo x cout/4)
  const View* out = nullptr;   // 16-bit NHWC destination slice, or
  float* out_f32 = nullptr;    // fp32 rows [img][pixel][out_ctot_f32] at column out_coff_f32
  long long out_img_stride = 0;
  int out_ctot_f32 = 0, out_coff_f32 = 0;
  const View* res = nullptr;
  const View* up = nullptr;
  const View* pre = nullptr;   // half-resolution pre-activation partial sum (see ConvParams::res_pre); excludes res
  int chain_cout = 0, chain_act = 1;   // > 0: a 1x1 conv with chain_cout outputs follows in the same kernel; `out` is ITS destination
};

// layer signature = key of the shipped per-layer kernel-variant table (csrc/conv_tune.inc)
struct ConvSig { int cin, cout, k, stride, H, W, flags; };   // flags: 1 residual, 2 upsampled copy, 4 f32 rows, 8 s2d output, 16 half-res pre-activation add, 32 chained 1x1 conv
ConvSig conv_signature(const ConvPlanArgs& a);
int conv_choose_variant(const ConvSig& s);

int conv_tc_init(gt_engine* e);  // resolves cuTensorMapEncodeTiled, sets kernel attributes
int conv_tc_plan(gt_engine* e, ConvOp* op, const ConvPlanArgs& a);  // tensor maps + launch geometry + weight storage
int conv_tc_pack_weights(gt_engine* e, ConvOp* op, const float* const* w, const float* const* b, const int* couts, int n);   // honours op->w_cin_total / w_cin_off / no_bias
int conv_tc_upload_packed(gt_engine* e, ConvOp* op, const uint16_t* packed, const float* bias);
int conv_tc_pack_chain(gt_engine* e, ConvOp* op, const float* w, const float* b);   // weights of the chained 1x1 conv (ConvOp::chain_src)
int conv_tc_launch(gt_engine* e, const ConvOp* op, int B, cudaStream_t st);
int conv_tc_launch_range(gt_engine* e, const ConvOp* op, int b0, int nb, cudaStream_t st);
// swapped-operand variant (conv_sw.cu); conv_tc_plan / conv_tc_launch* dispatch to it when the engine's swap_mode is on
typedef CUresult (*GtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
GtEncodeTiledFn conv_tc_encode();
struct View;
CUresult encode_s2d_out(CUtensorMap* tm, CUtensorMapDataType dt, const View* out, int Wo, int Ho, int B, int tw, int th);
int conv_tc_num_sms();
int conv_sw_init(gt_engine* e);
int conv_sw_plan(gt_engine* e, ConvOp* op, const ConvPlanArgs& a);
int conv_sw_launch_range(gt_engine* e, const ConvOp* op, int b0, int nb, cudaStream_t st);

// ---- small helpers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }  // MUFU.EX2 + MUFU.RCP
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
