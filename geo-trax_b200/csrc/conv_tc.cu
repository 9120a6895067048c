// bf16 implicit-GEMM convolution on tcgen05 / TMEM, operands staged by TMA (sm_100a).
//
// Replaces torch.nn.Conv2d + folded BatchNorm + SiLU of the fused YOLOv8s graph that the reference reaches through
// ultralytics (/root/reference/geotrax/extract.py:153; layer table SURVEY.md section 8a-4).
//
// Mapping: GEMM M = 128 output pixels (a tw x th spatial patch of one image), N = BN output channels (<= 256),
// K = taps x cin in blocks of 64 channels.  For each (tap, channel-block) the A tile is ONE 4-D TMA box
// {64 ch, tw*s, th*s, 1} of the NHWC input shifted by the tap offset -- out-of-image rows/columns are zero-filled by
// the TMA unit (the convolution's zero padding), the conv stride is the tensor map's elementStrides, and the box lands
// in shared memory as 128 rows x 128 B, which is exactly the K-major SWIZZLE_128B layout tcgen05.mma consumes.  No
// im2col buffer exists anywhere.  B tiles are {64, BN} boxes of the packed weights [cout][tap][cin_pad].
// Accumulators live in TMEM (128 lanes x BN fp32 columns); the epilogue reads them with tcgen05.ld and applies
// bias + SiLU (+ residual) and writes bf16 NHWC straight into a channel slice of the consumer's concat buffer
// (optionally also a 2x nearest-upsampled copy), or fp32 rows of the raw head tensor.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_idx % 4).
#include "engine.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

constexpr int kThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory operand descriptor (sm_100 UMMA format): start>>4 | SBO(1024 B)>>4 @32 |
// version 1 @46 | layout SWIZZLE_128B (2) @61.  LBO is unused for swizzled K-major operands.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = bn.
__device__ __forceinline__ uint32_t make_idesc(int bn, int fp16) {
  const uint32_t fmt = fp16 ? 0u : 1u;  // a/b format: 0 = F16, 1 = BF16
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// Epilogue for `ncols` (16 or 32) consecutive accumulator columns of one output pixel.
template <int NC>
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, const uint32_t* v, const float* s_bias, int col0, int n0,
                                               bool valid, int n, int y, int x) {
  if (!valid) return;
  const int gc0 = n0 + col0;  // first global output channel of this chunk
  if (gc0 >= p.cout) return;
  float f[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    float a = __uint_as_float(v[i]) + s_bias[col0 + i];
    f[i] = p.act ? silu_f(a) : a;
  }
  const long long pix = (long long)n * p.out_img_stride + (long long)y * p.W + x;
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + pix * p.out_ctot + p.out_coff + gc0;
    const int nv = min(NC, p.cout - gc0);
    for (int i = 0; i < nv; ++i) o[i] = f[i];
    return;
  }
  if (p.res) {
    const long long rp = ((long long)n * p.H + y) * p.W + x;
    const uint4* r = reinterpret_cast<const uint4*>(p.res + rp * p.res_ctot + p.res_coff + gc0);
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
      const uint4 u = __ldg(r + q);
      const float2 a0 = unpack2_act(u.x, p.fp16), a1 = unpack2_act(u.y, p.fp16), a2 = unpack2_act(u.z, p.fp16), a3 = unpack2_act(u.w, p.fp16);
      f[q * 8 + 0] += a0.x; f[q * 8 + 1] += a0.y; f[q * 8 + 2] += a1.x; f[q * 8 + 3] += a1.y;
      f[q * 8 + 4] += a2.x; f[q * 8 + 5] += a2.y; f[q * 8 + 6] += a3.x; f[q * 8 + 7] += a3.y;
    }
  }
  uint4 pk[NC / 8];
#pragma unroll
  for (int q = 0; q < NC / 8; ++q) {
    pk[q].x = pack2_act(f[q * 8 + 0], f[q * 8 + 1], p.fp16);
    pk[q].y = pack2_act(f[q * 8 + 2], f[q * 8 + 3], p.fp16);
    pk[q].z = pack2_act(f[q * 8 + 4], f[q * 8 + 5], p.fp16);
    pk[q].w = pack2_act(f[q * 8 + 6], f[q * 8 + 7], p.fp16);
  }
  uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + pix * p.out_ctot + p.out_coff + gc0);
#pragma unroll
  for (int q = 0; q < NC / 8; ++q) o[q] = pk[q];
  if (p.up) {
    const int W2 = p.W * 2;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const long long up = ((long long)n * (p.H * 2) + (y * 2 + (d >> 1))) * W2 + (x * 2 + (d & 1));
      uint4* u = reinterpret_cast<uint4*>(p.up + up * p.up_ctot + p.up_coff + gc0);
#pragma unroll
      for (int q = 0; q < NC / 8; ++q) u[q] = pk[q];
    }
  }
}

__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: 1024-aligned tiles first, then barriers / tmem pointer / bias
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.BN * 128;
  const int stage_bytes = kABytes + b_bytes;
  uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr_smem + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int m = blockIdx.x;
  const int tx = m % p.tiles_x;
  const int ty = (m / p.tiles_x) % p.tiles_y;
  const int n = m / (p.tiles_x * p.tiles_y);
  const int x0 = tx * p.tw, y0 = ty * p.th;
  const int n0 = blockIdx.y * p.BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.BN; i += 128) s_bias[i] = p.bias[n0 + i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)stage_bytes;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, tx_bytes);
        const int tap = kb / p.kc_blocks, kc = kb - tap * p.kc_blocks;
        const int dy = tap / p.ksize, dx = tap - dy * p.ksize;
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        tma_load_4d(smem_u32(sa), &tmA, fb, kc * 64, x0 * p.stride + dx - p.pad, y0 * p.stride + dy - p.pad, n);
        tma_load_2d(smem_u32(sa + kABytes), &tmB, fb, kb * 64, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.BN, p.fp16);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        const uint64_t adesc = make_sw128_desc(smem_u32(sa));
        const uint64_t bdesc = make_sw128_desc(smem_u32(sa + kABytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128-B swizzle atom: +2 in the (>>4) start-address field
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));  // frees this smem stage when the MMAs above retire
      }
      umma_commit(smem_u32(tmem_full_bar));    // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> bias/SiLU/residual -> global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;
    const int ly = row / p.tw, lx = row - ly * p.tw;
    const int y = y0 + ly, x = x0 + lx;
    const bool valid = (y < p.H) && (x < p.W) && (n < p.B);
    mbar_wait(smem_u32(tmem_full_bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    int c0 = 0;
    for (; c0 + 32 <= p.BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(trow + (uint32_t)c0, v);
      tmem_ld_wait();
      epilogue_chunk<32>(p, v, s_bias, c0, n0, valid, n, y, x);
    }
    if (c0 < p.BN) {
      uint32_t v[16];
      tmem_ld_x16(trow + (uint32_t)c0, v);
      tmem_ld_wait();
      epilogue_chunk<16>(p, v, s_bias, c0, n0, valid, n, y, x);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

size_t conv_smem_bytes(int stages, int bn) {
  return 1024 /*alignment slack*/ + (size_t)stages * (kABytes + bn * 128) + (2 * stages + 1) * 8 + 8 + (size_t)bn * 4 + 16;
}

}  // namespace

int conv_tc_init(gt_engine* e) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GT_CUDA(e, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    GT_CHECK(e, fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  GT_CUDA(e, cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return GT_OK;
}

static void pick_tile(int H, int W, int* tw, int* th) {
  const int cand[5][2] = {{16, 8}, {32, 4}, {8, 16}, {64, 2}, {128, 1}};
  double best = -1;
  for (int i = 0; i < 5; ++i) {
    const int w = cand[i][0], h = cand[i][1];
    const double cover = (double)ceil_div(W, w) * w * (double)ceil_div(H, h) * h;
    const double util = (double)H * W / cover;
    if (util > best + 1e-9) {
      best = util;
      *tw = w;
      *th = h;
    }
  }
}

int conv_tc_plan(gt_engine* e, ConvOp* op, const View& in, int Bmax, int cin, int cout_total, int k, int stride, int act,
                 const View* out, float* out_f32_ptr, long long out_img_stride, int out_ctot_f32, int out_coff_f32,
                 const View* res, const View* up) {
  GT_CHECK(e, g_encode != nullptr, "conv_tc_init not called");
  GT_CHECK(e, in.C == cin, "conv plan: input view has %d channels, conv expects %d", in.C, cin);
  GT_CHECK(e, (in.ctot % 8) == 0 && (in.coff % 8) == 0, "conv plan: input slice must be 16-byte aligned");
  GT_CHECK(e, k == 1 || k == 3, "conv plan: k=%d unsupported", k);
  GT_CHECK(e, stride == 1 || stride == 2, "conv plan: stride=%d unsupported", stride);
  ConvParams& p = op->p;
  memset(&p, 0, sizeof(p));
  const int pad = k / 2;
  const int Ho = (in.H + 2 * pad - k) / stride + 1, Wo = (in.W + 2 * pad - k) / stride + 1;
  op->cin = cin; op->cout = cout_total; op->k = k; op->stride = stride;
  p.B = Bmax; p.H = Ho; p.W = Wo;
  pick_tile(Ho, Wo, &p.tw, &p.th);
  p.tiles_x = ceil_div(Wo, p.tw); p.tiles_y = ceil_div(Ho, p.th);
  p.stride = stride; p.ksize = k; p.pad = pad;
  p.kc_blocks = ceil_div(cin, 64);
  op->cin_pad = p.kc_blocks * 64;
  p.num_kb = k * k * p.kc_blocks;
  const int cout16 = ceil_div(cout_total, 16) * 16;
  p.BN = cout16 <= 256 ? cout16 : 256;
  const int n_tiles = ceil_div(cout16, p.BN);
  op->cout_pad = n_tiles * p.BN;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.BN) p.tmem_cols *= 2;
  const int stage_bytes = kABytes + p.BN * 128;
  int stages = p.BN <= 128 ? (100 * 1024) / stage_bytes : (200 * 1024) / stage_bytes;  // <=128: two CTAs per SM
  if (stages > 8) stages = 8;
  if (stages > p.num_kb) stages = p.num_kb < 2 ? 2 : p.num_kb;
  p.stages = stages;
  p.cout = cout_total; p.act = act; p.fp16 = e->cfg.act_dtype == GT_ACT_FP16 ? 1 : 0;
  if (out_f32_ptr) {
    p.out_f32 = 1; p.out = out_f32_ptr; p.out_img_stride = out_img_stride; p.out_ctot = out_ctot_f32; p.out_coff = out_coff_f32;
  } else {
    GT_CHECK(e, out && out->H == Ho && out->W == Wo && out->C == cout_total, "conv plan: output view mismatch (%dx%dx%d vs %dx%dx%d)",
             out ? out->H : -1, out ? out->W : -1, out ? out->C : -1, Ho, Wo, cout_total);
    GT_CHECK(e, (out->ctot % 8) == 0 && (out->coff % 8) == 0 && (cout_total % 8) == 0, "conv plan: output slice must be 16-byte aligned");
    p.out_f32 = 0; p.out = out->ptr; p.out_img_stride = (long long)Ho * Wo; p.out_ctot = out->ctot; p.out_coff = out->coff;
  }
  if (res) {
    GT_CHECK(e, res->H == Ho && res->W == Wo && res->C == cout_total && !out_f32_ptr, "conv plan: residual view mismatch");
    p.res = res->ptr; p.res_ctot = res->ctot; p.res_coff = res->coff;
  }
  if (up) {
    GT_CHECK(e, up->H == 2 * Ho && up->W == 2 * Wo && up->C == cout_total && !out_f32_ptr, "conv plan: upsample view mismatch");
    p.up = up->ptr; p.up_ctot = up->ctot; p.up_coff = up->coff;
  }
  op->grid = dim3((unsigned)(p.tiles_x * p.tiles_y * Bmax), (unsigned)n_tiles, 1);
  op->smem = conv_smem_bytes(p.stages, p.BN);
  op->flops = 2.0 * Ho * Wo * (double)cout_total * cin * k * k;

  // weights + bias storage
  const size_t wn = (size_t)op->cout_pad * k * k * op->cin_pad;
  GT_TRY(e->dev_alloc((void**)&op->w_dev, wn * sizeof(bf16)));
  GT_TRY(e->dev_alloc((void**)&op->b_dev, (size_t)op->cout_pad * sizeof(float)));
  GT_CUDA(e, cudaMemset(op->w_dev, 0, wn * sizeof(bf16)));
  GT_CUDA(e, cudaMemset(op->b_dev, 0, (size_t)op->cout_pad * sizeof(float)));
  p.bias = op->b_dev;

  // A: NHWC input slice as a 4-D tensor {C, W, H, N}
  {
    cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)Bmax};
    cuuint64_t gstr[3] = {(cuuint64_t)in.ctot * 2, (cuuint64_t)in.W * in.ctot * 2, (cuuint64_t)in.H * in.W * in.ctot * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.tw * stride), (cuuint32_t)(p.th * stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode(&op->tmA, p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)(in.ptr + in.coff), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d ctot=%d box %dx%dx%d s=%d)", (int)r, cin,
             in.W, in.H, in.ctot, 64, p.tw * stride, p.th * stride, stride);
  }
  // B: packed weights as a 2-D tensor {Ktot, cout_pad}
  {
    const cuuint64_t ktot = (cuuint64_t)k * k * op->cin_pad;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)op->cout_pad};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p.BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&op->tmB, p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)op->w_dev, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d (ktot=%llu cout_pad=%d BN=%d)", (int)r,
             (unsigned long long)ktot, op->cout_pad, p.BN);
  }
  return GT_OK;
}

int conv_tc_pack_weights(gt_engine* e, ConvOp* op, const float* const* w, const float* const* b, const int* couts, int n) {
  const int taps = op->k * op->k;
  const size_t wn = (size_t)op->cout_pad * taps * op->cin_pad;
  const int fp16 = op->p.fp16;
  std::vector<uint16_t> hw(wn, 0);
  std::vector<float> hb(op->cout_pad, 0.f);
  int co0 = 0;
  for (int s = 0; s < n; ++s) {
    for (int co = 0; co < couts[s]; ++co) {
      for (int ci = 0; ci < op->cin; ++ci)
        for (int t = 0; t < taps; ++t)
          hw[((size_t)(co0 + co) * taps + t) * op->cin_pad + ci] = host_to_act(w[s][((size_t)co * op->cin + ci) * taps + t], fp16);
      hb[co0 + co] = b[s] ? b[s][co] : 0.f;
    }
    co0 += couts[s];
  }
  GT_CHECK(e, co0 == op->cout, "pack weights: cout mismatch %d vs %d", co0, op->cout);
  GT_CUDA(e, cudaMemcpy(op->w_dev, hw.data(), wn * sizeof(bf16), cudaMemcpyHostToDevice));
  GT_CUDA(e, cudaMemcpy(op->b_dev, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
  return GT_OK;
}

int conv_tc_launch(gt_engine* e, const ConvOp* op, int B, cudaStream_t st) {
  ConvParams p = op->p;
  p.B = B;
  dim3 grid((unsigned)(p.tiles_x * p.tiles_y * B), op->grid.y, 1);
  conv_tc_kernel<<<grid, kThreads, op->smem, st>>>(op->tmA, op->tmB, p);
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}
