// 16-bit implicit-GEMM convolution on tcgen05 / TMEM, operands staged by TMA (sm_100a).  Persistent, warp-specialised.
//
// Replaces torch.nn.Conv2d + folded BatchNorm + SiLU of the fused YOLOv8s graph that the reference reaches through
// ultralytics (/root/reference/geotrax/extract.py:153; layer table SURVEY.md section 8a-4).
//
// Mapping: GEMM M = 128 output pixels (a tw x th spatial patch of one image), N = BN output channels (<= 256),
// K = taps x cin in k-blocks of `kb_elems` channels (64 -> 128-byte rows, SWIZZLE_128B; 16 -> 32-byte rows, SWIZZLE_32B,
// used by layer 0 on its space-to-depth input).  For each (tap, channel-block) the A tile is ONE 4-D TMA box
// {kb_elems ch, tw*s, th*s, 1} of the NHWC input shifted by the tap offset -- out-of-image rows/columns are zero-filled
// by the TMA unit (the convolution's zero padding), the conv stride is the tensor map's elementStrides, and the box
// lands in shared memory as 128 rows of one k-block, which is exactly the K-major swizzled layout tcgen05.mma consumes.
// No im2col buffer exists anywhere.  B tiles are {kb_elems, BN} boxes of the packed weights [cout][tap][cin_pad].
//
// One CTA per SM loops over its tiles (tile = blockIdx.x, += gridDim.x).  Three pipelines run concurrently:
//   warp 0      TMA producer    : shared-memory ring (full/empty mbarriers) that keeps flowing across tile boundaries
//   warp 1      MMA issuer      : one thread issues tcgen05.mma into one of TWO TMEM accumulators (tmem full/empty mbarriers)
//   warps 2..9  epilogue        : tcgen05.ld the other accumulator, scale + bias + SiLU (+ residual), 16-bit NHWC stores
//                                 straight into a channel slice of the consumer's concat buffer (optionally also a 2x
//                                 nearest-upsampled copy), or fp32 rows of the raw head tensor
// so the epilogue of tile i overlaps the loads and MMAs of tile i+1 and the per-CTA prologue (barrier init, TMEM
// allocation, descriptor prefetch) is paid once per SM instead of once per tile.
#include "engine.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = GT_NUM_SMS;

constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;  // 320
constexpr int kMaxStages = 12;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// K-major swizzled shared-memory operand descriptor (sm_100 UMMA format): start>>4 | SBO>>4 @32 | version 1 @46 |
// layout type @61 (2 = SWIZZLE_128B with SBO 1024 B, 6 = SWIZZLE_32B with SBO 256 B: 8 rows of one swizzle atom).
// LBO is unused for swizzled K-major operands.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16 or f16, both K-major, M = 128, N = bn.
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

// Epilogue for NC (16 or 32) consecutive accumulator columns of one output pixel.
template <int NC>
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, const uint32_t* v, const float* s_bias, int gc0, bool valid, int n,
                                               int y, int x) {
  if (!valid || gc0 >= p.cout) return;
  float f[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const float a = fmaf(__uint_as_float(v[i]), p.scale, s_bias[gc0 + i]);
    f[i] = p.act ? silu_f(a) : a;
  }
  const long long pix = (long long)n * p.out_img_stride + (long long)y * p.W + x;
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + pix * p.out_ctot + p.out_coff + gc0;
    const int nv = p.cout - gc0;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (i < nv) o[i] = f[i];
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

struct TileCoord { int n, y0, x0, n0; };
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
  TileCoord t;
  const int nt = tile % p.n_tiles;
  const int m = tile / p.n_tiles;
  const int tx = m % p.tiles_x;
  const int r = m / p.tiles_x;
  const int ty = r % p.tiles_y;
  t.n = r / p.tiles_y + p.img0;
  t.x0 = tx * p.tw;
  t.y0 = ty * p.th;
  t.n0 = nt * p.BN;
  return t;
}

__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: 1024-aligned operand ring first, then barriers / tmem pointer / bias
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int row_bytes = p.kb_elems * 2;
  const int a_bytes = 128 * row_bytes;
  const int b_bytes = p.BN * row_bytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] accumulator drained
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr_smem + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.total_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    const int nb = p.n_tiles * p.BN;
    for (int i = threadIdx.x - 64; i < nb; i += kEpiWarps * 32) s_bias[i] = p.bias[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)stage_bytes;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int cx = t.x0 * p.stride - p.pad, cy = t.y0 * p.stride - p.pad;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const uint32_t s = it % (uint32_t)p.stages;
          const uint32_t ph = (it / (uint32_t)p.stages) & 1u;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, tx_bytes);
          const int tap = kb / p.kc_blocks, kc = kb - tap * p.kc_blocks;
          const int dy = tap / p.ksize, dx = tap - dy * p.ksize;
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          tma_load_4d(smem_u32(sa), &tmA, fb, kc * p.kb_elems, cx + dx, cy + dy, t.n);
          tma_load_2d(smem_u32(sa + a_bytes), &tmB, fb, kb * p.kb_elems, t.n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.BN, p.fp16);
      const bool sw128 = p.kb_elems == 64;
      const uint32_t sbo = sw128 ? 1024u : 256u, layout = sw128 ? 2u : 6u;
      const int mma_per_kb = p.kb_elems >> 4;
      uint32_t it = 0, li = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
        const uint32_t as = li & 1u;
        mbar_wait(smem_u32(&tempty_bar[as]), ((li >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + as * (uint32_t)p.acc_stride;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const uint32_t s = it % (uint32_t)p.stages;
          const uint32_t ph = (it / (uint32_t)p.stages) & 1u;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          const uint64_t adesc = make_desc(smem_u32(sa), sbo, layout);
          const uint64_t bdesc = make_desc(smem_u32(sa + a_bytes), sbo, layout);
          for (int k = 0; k < mma_per_kb; ++k) {
            // advance 16 elements = 32 B along K inside the 128-B swizzle atom: +2 in the (>>4) start-address field
            umma_f16(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[s]));  // frees this smem stage when the MMAs above retire
        }
        umma_commit(smem_u32(&tfull_bar[as]));   // accumulator complete
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> scale/bias/SiLU/residual -> global =====
    const int q = warp & 3;             // TMEM lane quarter this warp may touch (warp id % 4)
    const int half = (warp - 2) >> 2;   // two warps share a quarter and alternate 32-column chunks
    const int row = q * 32 + lane;
    const int ly = row / p.tw, lx = row - ly * p.tw;
    uint32_t li = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
      const uint32_t as = li & 1u;
      const TileCoord t = decode_tile(p, tile);
      const int y = t.y0 + ly, x = t.x0 + lx;
      const bool valid = (y < p.H) && (x < p.W);
      mbar_wait(smem_u32(&tfull_bar[as]), (li >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)p.acc_stride;
      int ci = 0;
      for (int c0 = 0; c0 < p.BN; c0 += 32, ++ci) {
        if ((ci & 1) != half) continue;
        if (c0 + 32 <= p.BN) {
          uint32_t v[32];
          tmem_ld_x32(trow + (uint32_t)c0, v);
          tmem_ld_wait();
          epilogue_chunk<32>(p, v, s_bias, t.n0 + c0, valid, t.n, y, x);
        } else {
          uint32_t v[16];
          tmem_ld_x16(trow + (uint32_t)c0, v);
          tmem_ld_wait();
          epilogue_chunk<16>(p, v, s_bias, t.n0 + c0, valid, t.n, y, x);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

size_t conv_smem_bytes(int stages, int stage_bytes, int bias_floats) {
  return 1024 /*alignment slack*/ + (size_t)stages * stage_bytes + (2 * kMaxStages + 4) * 8 + 8 + (size_t)bias_floats * 4 + 16;
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
  cudaDeviceProp prop;
  GT_CUDA(e, cudaGetDeviceProperties(&prop, e->device));
  g_num_sms = prop.multiProcessorCount;
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

int conv_tc_plan(gt_engine* e, ConvOp* op, const ConvPlanArgs& a) {
  GT_CHECK(e, g_encode != nullptr, "conv_tc_init not called");
  const View& in = a.in;
  const int cin = a.cin, k = a.k, stride = a.stride, kbe = a.kb_elems;
  GT_CHECK(e, in.C == cin, "conv plan: input view has %d channels, conv expects %d", in.C, cin);
  GT_CHECK(e, (in.ctot % 8) == 0 && (in.coff % 8) == 0, "conv plan: input slice must be 16-byte aligned");
  GT_CHECK(e, k >= 1 && k <= 3, "conv plan: k=%d unsupported", k);
  GT_CHECK(e, stride == 1 || stride == 2, "conv plan: stride=%d unsupported", stride);
  GT_CHECK(e, kbe == 64 || kbe == 16, "conv plan: kb_elems=%d unsupported", kbe);
  ConvParams& p = op->p;
  memset(&p, 0, sizeof(p));
  const int pad = a.pad >= 0 ? a.pad : k / 2;
  const int Ho = a.Ho > 0 ? a.Ho : (in.H + 2 * pad - k) / stride + 1, Wo = a.Wo > 0 ? a.Wo : (in.W + 2 * pad - k) / stride + 1;
  const int cout_total = a.cout;
  op->cin = cin; op->cout = cout_total; op->k = k; op->stride = stride;
  p.B = a.Bmax; p.H = Ho; p.W = Wo;
  pick_tile(Ho, Wo, &p.tw, &p.th);
  p.tiles_x = ceil_div(Wo, p.tw); p.tiles_y = ceil_div(Ho, p.th);
  p.stride = stride; p.ksize = k; p.pad = pad;
  p.kb_elems = kbe;
  p.kc_blocks = ceil_div(cin, kbe);
  op->cin_pad = p.kc_blocks * kbe;
  p.num_kb = k * k * p.kc_blocks;
  const int cout16 = ceil_div(cout_total, 16) * 16;
  p.BN = cout16 <= 256 ? cout16 : 256;
  p.n_tiles = ceil_div(cout16, p.BN);
  op->cout_pad = p.n_tiles * p.BN;
  // two accumulators of BN fp32 columns each; the allocation is a power of two >= 32 columns
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.BN) p.tmem_cols *= 2;
  p.acc_stride = p.tmem_cols / 2;
  const int stage_bytes = (128 + p.BN) * kbe * 2;
  const size_t fixed = conv_smem_bytes(0, stage_bytes, op->cout_pad);
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  GT_CHECK(e, stages >= 2, "conv plan: tile does not fit shared memory (BN=%d)", p.BN);
  p.stages = stages;
  p.cout = cout_total; p.act = a.act; p.fp16 = e->cfg.act_dtype == GT_ACT_FP16 ? 1 : 0;
  p.scale = a.scale;
  if (a.out_f32) {
    p.out_f32 = 1; p.out = a.out_f32; p.out_img_stride = a.out_img_stride; p.out_ctot = a.out_ctot_f32; p.out_coff = a.out_coff_f32;
  } else {
    const View* out = a.out;
    GT_CHECK(e, out && out->H == Ho && out->W == Wo && out->C == cout_total, "conv plan: output view mismatch (%dx%dx%d vs %dx%dx%d)",
             out ? out->H : -1, out ? out->W : -1, out ? out->C : -1, Ho, Wo, cout_total);
    GT_CHECK(e, (out->ctot % 8) == 0 && (out->coff % 8) == 0 && (cout_total % 8) == 0, "conv plan: output slice must be 16-byte aligned");
    p.out_f32 = 0; p.out = out->ptr; p.out_img_stride = (long long)Ho * Wo; p.out_ctot = out->ctot; p.out_coff = out->coff;
  }
  if (a.res) {
    GT_CHECK(e, a.res->H == Ho && a.res->W == Wo && a.res->C == cout_total && !a.out_f32, "conv plan: residual view mismatch");
    p.res = a.res->ptr; p.res_ctot = a.res->ctot; p.res_coff = a.res->coff;
  }
  if (a.up) {
    GT_CHECK(e, a.up->H == 2 * Ho && a.up->W == 2 * Wo && a.up->C == cout_total && !a.out_f32, "conv plan: upsample view mismatch");
    p.up = a.up->ptr; p.up_ctot = a.up->ctot; p.up_coff = a.up->coff;
  }
  op->smem = conv_smem_bytes(p.stages, stage_bytes, op->cout_pad);
  op->flops = 2.0 * Ho * Wo * (double)cout_total * cin * k * k;
  // algorithmic HBM bytes per image: input slice + output (+ residual, + upsampled copy) + weights (once per launch, ignored)
  op->bytes = (double)in.H * in.W * cin * 2 + (double)Ho * Wo * cout_total * (a.out_f32 ? 4 : 2) * (a.up ? 5 : 1) +
              (a.res ? (double)Ho * Wo * cout_total * 2 : 0.0);

  // weights + bias storage
  const size_t wn = (size_t)op->cout_pad * k * k * op->cin_pad;
  GT_TRY(e->dev_alloc((void**)&op->w_dev, wn * sizeof(bf16)));
  GT_TRY(e->dev_alloc((void**)&op->b_dev, (size_t)op->cout_pad * sizeof(float)));
  GT_CUDA(e, cudaMemset(op->w_dev, 0, wn * sizeof(bf16)));
  GT_CUDA(e, cudaMemset(op->b_dev, 0, (size_t)op->cout_pad * sizeof(float)));
  p.bias = op->b_dev;

  const CUtensorMapDataType dt = p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = kbe == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  // A: NHWC input slice as a 4-D tensor {C, W, H, N}
  {
    cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)a.Bmax};
    cuuint64_t gstr[3] = {(cuuint64_t)in.ctot * 2, (cuuint64_t)in.W * in.ctot * 2, (cuuint64_t)in.H * in.W * in.ctot * 2};
    cuuint32_t box[4] = {(cuuint32_t)kbe, (cuuint32_t)(p.tw * stride), (cuuint32_t)(p.th * stride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode(&op->tmA, dt, 4, (void*)(in.ptr + in.coff), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d ctot=%d box %dx%dx%d s=%d)", (int)r, cin,
             in.W, in.H, in.ctot, kbe, p.tw * stride, p.th * stride, stride);
  }
  // B: packed weights as a 2-D tensor {Ktot, cout_pad}
  {
    const cuuint64_t ktot = (cuuint64_t)k * k * op->cin_pad;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)op->cout_pad};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kbe, (cuuint32_t)p.BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&op->tmB, dt, 2, (void*)op->w_dev, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d (ktot=%llu cout_pad=%d BN=%d)", (int)r,
             (unsigned long long)ktot, op->cout_pad, p.BN);
  }
  return GT_OK;
}

// w[s]: f32 [couts[s]][cin][k][k] (PyTorch layout); several convs reading the same input are stacked along cout
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
  return conv_tc_upload_packed(e, op, hw.data(), hb.data());
}

// packed: 16-bit [cout_pad][taps][cin_pad] already in the activation format; bias f32 [cout_pad]
int conv_tc_upload_packed(gt_engine* e, ConvOp* op, const uint16_t* packed, const float* bias) {
  const size_t wn = (size_t)op->cout_pad * op->k * op->k * op->cin_pad;
  GT_CUDA(e, cudaMemcpy(op->w_dev, packed, wn * sizeof(bf16), cudaMemcpyHostToDevice));
  GT_CUDA(e, cudaMemcpy(op->b_dev, bias, (size_t)op->cout_pad * sizeof(float), cudaMemcpyHostToDevice));
  return GT_OK;
}

int conv_tc_launch(gt_engine* e, const ConvOp* op, int B, cudaStream_t st) { return conv_tc_launch_range(e, op, 0, B, st); }

// images [b0, b0 + nb) of the batch the op was planned for (tensor maps cover Bmax images; tiles are offset by b0)
int conv_tc_launch_range(gt_engine* e, const ConvOp* op, int b0, int nb, cudaStream_t st) {
  ConvParams p = op->p;
  p.B = nb;
  p.img0 = b0;
  p.total_tiles = p.tiles_x * p.tiles_y * nb * p.n_tiles;
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  conv_tc_kernel<<<grid, kThreads, op->smem, st>>>(op->tmA, op->tmB, p);
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}
