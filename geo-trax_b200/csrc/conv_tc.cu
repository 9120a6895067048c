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
// Variants (chosen per layer by detector_autotune): OCC = 2 compiles the kernel for two CTAs per SM (layers whose tiles carry little MMA
// work are bound by the latency of the per-tile epilogue chain); p.halo stages one (8 + k - 1) x (16 + k - 1) pixel box per k-block and
// reads the k*k taps through row-shifted UMMA descriptors (the per-tap path is bound by TMA box rows, ~3.4 cycles per row per SM).
#include <algorithm>

#include "engine.cuh"
#include "tc_ptx.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = GT_NUM_SMS;

constexpr int kThreads = 64 + kEpiWarps * 32;  // 320
constexpr int kMaxStages = 12;
constexpr int kMaxHaloStages = 8;

// Epilogue of one accumulator: slabs of 128 rows x SLAB_BYTES (128 or 64) are staged in shared memory in the TMA swizzle
// layout and written with one bulk tensor store each (plus four for the 2x nearest-upsampled copy).  CW = accumulator
// columns per warp per slab; two warps share a TMEM lane quarter and take the two halves of a slab's columns.
template <int CW, bool F32>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, const TileCoord& t, uint32_t trow, uint32_t stage_base, uint32_t s_bias,
                                              const CUtensorMap* tmOut, const CUtensorMap* tmUp, int half, int row, int y, int x, bool valid,
                                              uint32_t& slab_ctr, uint32_t tempty_bar, int lane, bool leader_warp, uint32_t tfull_bar, uint32_t tfull_parity) {
  constexpr int ELEM = F32 ? 4 : 2;
  constexpr int ROW_BYTES = 2 * CW * ELEM;            // slab row: both halves
  constexpr int CHUNKS = CW * ELEM / 16;              // 16-byte chunks this thread writes per slab row
  constexpr uint32_t SWZ_MASK = ROW_BYTES == 128 ? 7u : (ROW_BYTES == 64 ? 3u : 1u);
  const int n_slabs = p.BN / (2 * CW);
  // The residual of a slab is fetched one step ahead -- slab 0's before the wait for the accumulator, slab j+1's while slab j is
  // processed -- so that its global-memory latency is off the per-tile epilogue chain (32->32 @272x480 with residual: 135 -> 80 us).
  uint4 rres[F32 ? 1 : CHUNKS];
  const bool has_res = !F32 && p.res != nullptr && valid;
  const uint4* rbase = nullptr;
  if constexpr (!F32) {
    if (has_res) {
      const long long rp = p.res_pre ? ((long long)t.n * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)   // half-resolution partial sum
                                     : ((long long)t.n * p.H + y) * p.W + x;
      rbase = reinterpret_cast<const uint4*>(p.res + rp * p.res_ctot + p.res_coff + t.n0 + half * CW);
      if (t.n0 + half * CW < p.cout) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) rres[c] = __ldg(rbase + c);
      }
    }
  }
  if (lane == 0) mbar_wait(tfull_bar, tfull_parity);   // one lane polls; the warp reconverges below
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int j = 0; j < n_slabs; ++j, ++slab_ctr) {
    const uint32_t buf = stage_base + (slab_ctr & 1u) * (128 * 128);
    // the bulk store that last read this buffer (two slabs ago) must have finished reading shared memory
    if (leader_warp && slab_ctr >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    epi_bar();
    uint32_t v[CW];
    const int c0 = j * 2 * CW + half * CW;            // first accumulator column of this warp's part
    tmem_ld<CW>(trow + (uint32_t)c0, v);
    tmem_ld_wait();
    if (j == n_slabs - 1) {                           // accumulator fully read: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
    const int gc0 = t.n0 + c0;
    float f[CW];
    const uint32_t bias_addr = s_bias + (uint32_t)gc0 * 4u;
#pragma unroll
    for (int i = 0; i < CW; i += 4) {
      const float4 bv = lds_f4(bias_addr + (uint32_t)i * 4u);
      f[i] = fmaf(__uint_as_float(v[i]), p.scale, bv.x);
      f[i + 1] = fmaf(__uint_as_float(v[i + 1]), p.scale, bv.y);
      f[i + 2] = fmaf(__uint_as_float(v[i + 2]), p.scale, bv.z);
      f[i + 3] = fmaf(__uint_as_float(v[i + 3]), p.scale, bv.w);
    }
    if constexpr (!F32) {
      if (p.res_pre && has_res && gc0 < p.cout) {   // pre-activation partial sum of the upsampled branch (warp-uniform branch)
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const uint4 u = rres[c];
          const float2 a0 = unpack2_act(u.x, p.fp16), a1 = unpack2_act(u.y, p.fp16), a2 = unpack2_act(u.z, p.fp16), a3 = unpack2_act(u.w, p.fp16);
          f[c * 8 + 0] += a0.x; f[c * 8 + 1] += a0.y; f[c * 8 + 2] += a1.x; f[c * 8 + 3] += a1.y;
          f[c * 8 + 4] += a2.x; f[c * 8 + 5] += a2.y; f[c * 8 + 6] += a3.x; f[c * 8 + 7] += a3.y;
        }
      }
    }
    if (p.act == 1) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = silu_fast(f[i]);
    } else if (p.act == 2) {   // one-MUFU form (MUFU-bound layers)
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = silu_tanh(f[i]);
    }
    uint4 pk[CHUNKS];
    if constexpr (F32) {
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c)
        pk[c] = make_uint4(__float_as_uint(f[c * 4]), __float_as_uint(f[c * 4 + 1]), __float_as_uint(f[c * 4 + 2]), __float_as_uint(f[c * 4 + 3]));
    } else {
      if (has_res && !p.res_pre && gc0 < p.cout) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const uint4 u = rres[c];
          const float2 a0 = unpack2_act(u.x, p.fp16), a1 = unpack2_act(u.y, p.fp16), a2 = unpack2_act(u.z, p.fp16), a3 = unpack2_act(u.w, p.fp16);
          f[c * 8 + 0] += a0.x; f[c * 8 + 1] += a0.y; f[c * 8 + 2] += a1.x; f[c * 8 + 3] += a1.y;
          f[c * 8 + 4] += a2.x; f[c * 8 + 5] += a2.y; f[c * 8 + 6] += a3.x; f[c * 8 + 7] += a3.y;
        }
      }
      if (has_res && j + 1 < n_slabs && gc0 + 2 * CW < p.cout) {   // next slab's residual: in flight during this slab's stores
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) rres[c] = __ldg(rbase + (size_t)(j + 1) * (2 * CW * ELEM / 16) + c);
      }
      if (p.fp16) {   // warp-uniform: keeps the two encodings out of each other's instruction stream
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          pk[c].x = pack2_act(f[c * 8 + 0], f[c * 8 + 1], 1); pk[c].y = pack2_act(f[c * 8 + 2], f[c * 8 + 3], 1);
          pk[c].z = pack2_act(f[c * 8 + 4], f[c * 8 + 5], 1); pk[c].w = pack2_act(f[c * 8 + 6], f[c * 8 + 7], 1);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          pk[c].x = pack2_act(f[c * 8 + 0], f[c * 8 + 1], 0); pk[c].y = pack2_act(f[c * 8 + 2], f[c * 8 + 3], 0);
          pk[c].z = pack2_act(f[c * 8 + 4], f[c * 8 + 5], 0); pk[c].w = pack2_act(f[c * 8 + 6], f[c * 8 + 7], 0);
        }
      }
    }
    // swizzled staging write: byte address bits [4:6] ^= bits [7:9] (masked to the swizzle span), as the TMA unit expects
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const uint32_t logical = (uint32_t)row * ROW_BYTES + (uint32_t)(half * CHUNKS + c) * 16u;
      const uint32_t phys = logical ^ (((logical >> 7) & SWZ_MASK) << 4);
      sts_u4(buf + phys, pk[c]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    epi_bar();
    if (leader_warp && lane == 0) {
      const int cch = (t.n0 + j * 2 * CW);            // first channel of this slab inside the destination slice
      if (p.out_s2d) tma_store_4d(tmOut, buf, 0, cch >> 6, t.x0, t.n * p.out_rows + t.y0);
      else tma_store_4d(tmOut, buf, cch, t.x0, t.y0, t.n);
      if (p.up) {
#pragma unroll
        for (int d = 0; d < 4; ++d) tma_store_4d(tmUp + d, buf, cch, t.x0, t.y0, t.n);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
}

// Epilogue of a tile whose conv is followed by a CHAINED 1x1 conv (ConvParams::chain_n): the activated 16-bit slabs of the main conv
// (128 pixels x 64 channels each, written in the 128-byte swizzle the TMA store would want) are at the same time canonical K-major
// SWIZZLE_128B A-operand tiles.  So instead of storing them, the leader warp issues a second GEMM D2[128 px][chain_n] = slab . W2^T
// against the resident chain weights into a third TMEM accumulator, every warp then runs a second epilogue on D2, and only that
// result is stored.  The intermediate tensor (model.1's output: 16.7 MB per 4K frame) is never written to or read from HBM.
__device__ __forceinline__ void epilogue_chain_tile(const ConvParams& p, const TileCoord& t, uint32_t trow, uint32_t trow2, uint32_t tacc2, uint32_t stage_base,
                                                    uint32_t s_bias, uint32_t s_bias2, const CUtensorMap* tmOut, int half, int row, uint32_t tempty_bar, int lane,
                                                    bool leader_warp, uint32_t tfull_bar, uint32_t tfull_parity, uint32_t chain_bar, uint32_t chain_parity,
                                                    uint32_t cres_lo, uint32_t idesc2, bool first_tile) {
  constexpr int CW = 32, CHUNKS = 4;
  const int n1 = p.BN >> 6, n2 = p.chain_n >> 6;
  const uint32_t hi = desc_hi(1024u, 2u);
  if (lane == 0) mbar_wait(tfull_bar, tfull_parity);
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // both staging buffers are free once the previous tile's bulk stores have read them
  if (leader_warp && lane == 0 && !first_tile) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  epi_bar();
  auto finish_slab = [&](const uint32_t* v, uint32_t bias_addr, int act, float scale, uint32_t buf) {
    float f[CW];
#pragma unroll
    for (int i = 0; i < CW; i += 4) {
      const float4 bv = lds_f4(bias_addr + (uint32_t)i * 4u);
      f[i] = fmaf(__uint_as_float(v[i]), scale, bv.x); f[i + 1] = fmaf(__uint_as_float(v[i + 1]), scale, bv.y);
      f[i + 2] = fmaf(__uint_as_float(v[i + 2]), scale, bv.z); f[i + 3] = fmaf(__uint_as_float(v[i + 3]), scale, bv.w);
    }
    act_inplace<CW>(f, act);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      uint4 pk;
      pk.x = pack2_act(f[c * 8 + 0], f[c * 8 + 1], p.fp16); pk.y = pack2_act(f[c * 8 + 2], f[c * 8 + 3], p.fp16);
      pk.z = pack2_act(f[c * 8 + 4], f[c * 8 + 5], p.fp16); pk.w = pack2_act(f[c * 8 + 6], f[c * 8 + 7], p.fp16);
      const uint32_t logical = (uint32_t)row * 128u + (uint32_t)(half * CHUNKS + c) * 16u;
      sts_u4(buf + (logical ^ (((logical >> 7) & 7u) << 4)), pk);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  };
  for (int j = 0; j < n1; ++j) {
    const uint32_t buf = stage_base + (uint32_t)j * (128 * 128);
    uint32_t v[CW];
    tmem_ld<CW>(trow + (uint32_t)(j * 64 + half * CW), v);
    tmem_ld_wait();
    if (j == n1 - 1) {                                // main accumulator fully read: hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
    finish_slab(v, s_bias + (uint32_t)(j * 64 + half * CW) * 4u, p.act, p.scale, buf);
    epi_bar();
    if (leader_warp) {                                // k-block j of the chained GEMM: A = the slab just written, B = W2[:, 64 j .. 64 j + 63]
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a_lo = (buf & 0x3FFFFu) >> 4, b_lo = cres_lo + (uint32_t)j * ((uint32_t)p.chain_n * 128u >> 4);
        issue_kb<4>(tacc2, a_lo, hi, b_lo, hi, idesc2, j ? 1u : 0u);
        if (j == n1 - 1) umma_commit(chain_bar);
      }
      __syncwarp();
    }
  }
  if (lane == 0) mbar_wait(chain_bar, chain_parity);
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int j = 0; j < n2; ++j) {
    const uint32_t buf = stage_base + (uint32_t)j * (128 * 128);   // (the chained MMAs have consumed both buffers)
    uint32_t v[CW];
    tmem_ld<CW>(trow2 + (uint32_t)(j * 64 + half * CW), v);
    tmem_ld_wait();
    if (j == n2 - 1) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // D2 is free for the next tile after the barrier below
    finish_slab(v, s_bias2 + (uint32_t)(j * 64 + half * CW) * 4u, p.chain_act, 1.0f, buf);
    epi_bar();
    if (leader_warp && lane == 0) {
      tma_store_4d(tmOut, buf, j * 64, t.x0, t.y0, t.n);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
}

// OCC = CTAs per SM the instantiation is compiled for: 2 (register cap 96, <= 113 KB shared memory, <= 256 TMEM columns per CTA)
// doubles the number of epilogue chains in flight for layers whose tiles carry little MMA work (small cout / small K).
template <int OCC>
__global__ void __launch_bounds__(kThreads, OCC) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                              const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvUpMaps tmUp,
                                                              const __grid_constant__ CUtensorMap tmC, const ConvParams p) {
  pdl_launch_dependents();   // the next layer's CTAs may start their prologue (barriers, TMEM, resident weights) as SMs free up
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve (all 1024-aligned): operand ring | resident weights (optional) | 2 output staging slabs | barriers / tmem pointer / bias
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int row_bytes = p.kb_elems * 2;
  const int a_bytes = 128 * row_bytes;
  const int b_bytes = p.BN * row_bytes;
  const int stage_bytes = p.halo ? b_bytes : (p.b_resident ? a_bytes : a_bytes + b_bytes);
  uint8_t* halo_ring = smem + (size_t)p.stages * stage_bytes;
  uint8_t* b_res = halo_ring + (size_t)p.a_stages * p.halo_bytes;
  uint8_t* c_res = b_res + (p.b_resident ? (size_t)p.num_kb * b_bytes : 0);   // chained conv: weights [BN / 64][chain_n rows][128 B]
  uint8_t* out_stage = c_res + (size_t)(p.BN >> 6) * p.chain_n * 128;
  uint8_t* tail = out_stage + 2 * 128 * 128;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* afull_bar = empty_bar + kMaxStages;   // halo ring
  uint64_t* aempty_bar = afull_bar + kMaxHaloStages;
  uint64_t* tfull_bar = aempty_bar + kMaxHaloStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] accumulator drained
  uint64_t* bres_bar = tempty_bar + 2;            // resident weights landed
  uint64_t* cres_bar = bres_bar + 1;              // chain weights landed
  uint64_t* chain_bar = cres_bar + 1;             // chained GEMM of the current tile complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(chain_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr_smem + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.total_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(smem_u32(&afull_bar[s]), 1);
      mbar_init(smem_u32(&aempty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), kEpiWarps);
    }
    mbar_init(smem_u32(bres_bar), 1);
    mbar_init(smem_u32(cres_bar), 1);
    mbar_init(smem_u32(chain_bar), 1);
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
    for (int i = threadIdx.x - 64; i < p.chain_n; i += kEpiWarps * 32) s_bias[nb + i] = p.chain_bias[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the loop (converged), one elected lane issues =====
    // (issuing from inside an `if (lane == 0)` region makes the compiler wrap every uniform-register operand of UTMALDG /
    // UTCHMMA in an ELECT / R2UR.BROADCAST waterfall loop -- measured ~190 cycles per MMA)
    if (p.b_resident) {   // all weights of this layer stay in shared memory for the CTA's lifetime
      if (elect_one()) {
        const uint32_t bb = smem_u32(bres_bar);
        mbar_expect_tx(bb, (uint32_t)(p.num_kb * b_bytes));
        for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(smem_u32(b_res + (size_t)kb * b_bytes), &tmB, bb, kb * p.kb_elems, 0);
      }
      __syncwarp();
    }
    if (p.chain_n) {
      if (elect_one()) {
        const uint32_t cb = smem_u32(cres_bar), kb_bytes = (uint32_t)p.chain_n * 128u;
        mbar_expect_tx(cb, (uint32_t)(p.BN >> 6) * kb_bytes);
        for (int kb = 0; kb < (p.BN >> 6); ++kb) tma_load_2d(smem_u32(c_res) + (uint32_t)kb * kb_bytes, &tmC, cb, kb * 64, 0);
      }
      __syncwarp();
    }
    pdl_wait();   // everything above touched only this layer's constants; activations of the previous layer are read below
    const uint32_t tx_bytes = (uint32_t)stage_bytes;
    uint32_t s = 0, ph = 0;
    TileIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    if (p.halo) {
      uint32_t hs = 0, hph = 0;
      const int taps = p.ksize * p.ksize;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ti.next(p)) {
        const TileCoord t = ti.coord(p);
        for (int kc = 0; kc < p.kc_blocks; ++kc) {
          mbar_wait(smem_u32(&aempty_bar[hs]), hph ^ 1u);
          if (elect_one()) {
            const uint32_t ab = smem_u32(&afull_bar[hs]);
            mbar_expect_tx(ab, (uint32_t)p.halo_tx);
            tma_load_4d(smem_u32(halo_ring + (size_t)hs * p.halo_bytes), &tmA, ab, kc * p.kb_elems, t.x0 - p.pad, t.y0 - p.pad, t.n);
          }
          __syncwarp();
          if (++hs == (uint32_t)p.a_stages) { hs = 0; hph ^= 1u; }
          if (!p.b_resident) {
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
              if (elect_one()) {
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_expect_tx(fb, tx_bytes);
                tma_load_2d(smem_u32(smem + (size_t)s * stage_bytes), &tmB, fb, (tap * p.kc_blocks + kc) * p.kb_elems, t.n0);
              }
              __syncwarp();
              if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
            }
          }
        }
      }
    } else {
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ti.next(p)) {
        const TileCoord t = ti.coord(p);
        const int cx = t.x0 * p.stride - p.pad, cy = t.y0 * p.stride - p.pad;
        int kb = 0;
        for (int dy = 0; dy < p.ksize; ++dy)
          for (int dx = 0; dx < p.ksize; ++dx)
            for (int kc = 0; kc < p.kc_blocks; ++kc, ++kb) {
              mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
              if (elect_one()) {
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_expect_tx(fb, tx_bytes);
                uint8_t* sa = smem + (size_t)s * stage_bytes;
                tma_load_4d(smem_u32(sa), &tmA, fb, kc * p.kb_elems, cx + dx, cy + dy, t.n);
                if (!p.b_resident) tma_load_2d(smem_u32(sa + a_bytes), &tmB, fb, kb * p.kb_elems, t.n0);
              }
              __syncwarp();
              if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
            }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, one elected lane issues tcgen05.mma / tcgen05.commit =====
    const uint32_t idesc = make_idesc(p.BN, p.fp16);
    // swizzle span = one k-block row: 128 B (64 ch), 64 B (32 ch) or 32 B (16 ch); SBO = 8 rows of it
    const uint32_t sbo = 8u * (uint32_t)row_bytes, layout = p.kb_elems == 64 ? 2u : (p.kb_elems == 32 ? 4u : 6u);
    const int mma_per_kb = p.kb_elems >> 4;
    if (p.b_resident) mbar_wait(smem_u32(bres_bar), 0);
    uint32_t s = 0, ph = 0, li = 0, hs = 0, hph = 0;
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4, bres_lo = (smem_u32(b_res) & 0x3FFFFu) >> 4;
    const uint32_t stage16 = (uint32_t)stage_bytes >> 4, a16 = (uint32_t)a_bytes >> 4, b16 = (uint32_t)b_bytes >> 4;
    const uint32_t hi_std = desc_hi(sbo, layout), full_bar_a = smem_u32(full_bar), empty_bar_a = smem_u32(empty_bar);
    const int pitch = p.tw + p.ksize - 1;                                       // halo row pitch in pixels
    const uint32_t hi_halo = desc_hi((uint32_t)(pitch * row_bytes), layout), row16 = (uint32_t)row_bytes >> 4;
    const uint32_t halo_lo = (smem_u32(halo_ring) & 0x3FFFFu) >> 4, halo16 = (uint32_t)p.halo_bytes >> 4;
    const uint32_t afull_bar_a = smem_u32(afull_bar), aempty_bar_a = smem_u32(aempty_bar);
    uint32_t s_lo = smem_lo;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
      const uint32_t as = li & 1u;
      mbar_wait(smem_u32(&tempty_bar[as]), ((li >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + as * (uint32_t)p.acc_stride;
      if (p.halo) {
        // one pixel box per k-block; tap (dy, dx) = the same box read from row dy * pitch + dx, 8-row groups one pitch apart.
        // Resident weights: the k*k taps of a k-block are issued back to back under one election.
        const uint32_t h_lo = halo_lo + hs * halo16;
        for (int kc = 0; kc < p.kc_blocks; ++kc) {
          mbar_wait(afull_bar_a + hs * 8u, hph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t hb_lo = halo_lo + hs * halo16;
          if (p.b_resident) {
            if (elect_one()) {
              uint32_t w_lo = bres_lo + (uint32_t)kc * b16;
              for (int dy = 0; dy < p.ksize; ++dy)
                for (int dx = 0; dx < p.ksize; ++dx) {
                  const uint32_t al = hb_lo + (uint32_t)(dy * pitch + dx) * row16;
                  const uint32_t acc = (kc | dy | dx) ? 1u : 0u;
                  if (mma_per_kb == 4) issue_kb<4>(tacc, al, hi_halo, w_lo, hi_std, idesc, acc);
                  else if (mma_per_kb == 2) issue_kb<2>(tacc, al, hi_halo, w_lo, hi_std, idesc, acc);
                  else issue_kb<1>(tacc, al, hi_halo, w_lo, hi_std, idesc, acc);
                  w_lo += b16 * (uint32_t)p.kc_blocks;
                }
              umma_commit(aempty_bar_a + hs * 8u);   // halo stage free once its taps have retired
            }
            __syncwarp();
          } else {
            for (int dy = 0; dy < p.ksize; ++dy)
              for (int dx = 0; dx < p.ksize; ++dx) {
                mbar_wait(full_bar_a + s * 8u, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                  const uint32_t al = hb_lo + (uint32_t)(dy * pitch + dx) * row16;
                  const uint32_t acc = (kc | dy | dx) ? 1u : 0u;
                  if (mma_per_kb == 4) issue_kb<4>(tacc, al, hi_halo, s_lo, hi_std, idesc, acc);
                  else if (mma_per_kb == 2) issue_kb<2>(tacc, al, hi_halo, s_lo, hi_std, idesc, acc);
                  else issue_kb<1>(tacc, al, hi_halo, s_lo, hi_std, idesc, acc);
                  umma_commit(empty_bar_a + s * 8u);
                  if (dy == p.ksize - 1 && dx == p.ksize - 1) umma_commit(aempty_bar_a + hs * 8u);
                }
                __syncwarp();
                s_lo += stage16;
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; s_lo = smem_lo; }
              }
          }
          if (++hs == (uint32_t)p.a_stages) { hs = 0; hph ^= 1u; }
        }
        (void)h_lo;
      } else {
        // lean issue: running low words, MMAs of a k-block unrolled at compile time (tc_ptx.cuh)
        uint32_t b_lo = bres_lo;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar_a + s * 8u, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t bl = p.b_resident ? b_lo : s_lo + a16;
            if (mma_per_kb == 4) issue_kb<4>(tacc, s_lo, hi_std, bl, hi_std, idesc, kb ? 1u : 0u);
            else if (mma_per_kb == 2) issue_kb<2>(tacc, s_lo, hi_std, bl, hi_std, idesc, kb ? 1u : 0u);
            else issue_kb<1>(tacc, s_lo, hi_std, bl, hi_std, idesc, kb ? 1u : 0u);
            umma_commit(empty_bar_a + s * 8u);  // frees this smem stage when the MMAs above retire
          }
          __syncwarp();
          b_lo += b16;
          s_lo += stage16;
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; s_lo = smem_lo; }
        }
      }
      if (elect_one()) umma_commit(smem_u32(&tfull_bar[as]));   // accumulator complete
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> scale/bias/SiLU/residual -> swizzled smem slab -> TMA store =====
    const int q = warp & 3;             // TMEM lane quarter this warp may touch (warp id % 4)
    const int half = (warp - 2) >> 2;   // two warps share a quarter and split each slab's columns
    const int row = q * 32 + lane;
    const int ly = row / p.tw, lx = row - ly * p.tw;
    const bool leader_warp = warp == 2;
    if (p.chain_n) {
      if (lane == 0) mbar_wait(smem_u32(cres_bar), 0);
      __syncwarp();
    }
    pdl_wait();   // residual reads and output stores below
    uint32_t li = 0, slab_ctr = 0;
    const uint32_t out_stage_a = smem_u32(out_stage), s_bias_a = smem_u32(s_bias);
    const uint32_t cres_lo = (smem_u32(c_res) & 0x3FFFFu) >> 4, idesc2 = make_idesc(p.chain_n ? p.chain_n : 64, p.fp16);
    TileIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li, ti.next(p)) {
      const uint32_t as = li & 1u;
      const TileCoord t = ti.coord(p);
      const int y = t.y0 + ly, x = t.x0 + lx;
      const bool valid = (y < p.H) && (x < p.W);
      const uint32_t tfb = smem_u32(&tfull_bar[as]), tfp = (li >> 1) & 1u;   // waited for inside epilogue_tile, after the residual prefetch
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)p.acc_stride;
      const uint32_t teb = smem_u32(&tempty_bar[as]);
#if defined(GT_TC_EXP) && GT_TC_EXP == 1   // debug experiment: no epilogue (accumulator released at once, nothing stored)
      if (lane == 0) mbar_wait(tfb, tfp);
      __syncwarp();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(teb);
      continue;
#endif
      if (p.chain_n) {
        epilogue_chain_tile(p, t, trow, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.chain_tmem, tmem_base + (uint32_t)p.chain_tmem, out_stage_a, s_bias_a,
                            s_bias_a + (uint32_t)(p.n_tiles * p.BN) * 4u, &tmOut, half, row, teb, lane, leader_warp, tfb, tfp, smem_u32(chain_bar), li & 1u, cres_lo,
                            idesc2, li == 0);
        continue;
      }
      switch (p.epi_mode) {
        case 0: epilogue_tile<32, false>(p, t, trow, out_stage_a, s_bias_a, &tmOut, tmUp.m, half, row, y, x, valid, slab_ctr, teb, lane, leader_warp, tfb, tfp); break;
        case 1: epilogue_tile<16, false>(p, t, trow, out_stage_a, s_bias_a, &tmOut, tmUp.m, half, row, y, x, valid, slab_ctr, teb, lane, leader_warp, tfb, tfp); break;
        case 2: epilogue_tile<16, true>(p, t, trow, out_stage_a, s_bias_a, &tmOut, tmUp.m, half, row, y, x, valid, slab_ctr, teb, lane, leader_warp, tfb, tfp); break;
        default: epilogue_tile<8, true>(p, t, trow, out_stage_a, s_bias_a, &tmOut, tmUp.m, half, row, y, x, valid, slab_ctr, teb, lane, leader_warp, tfb, tfp); break;
      }
    }
    if (leader_warp && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output stores complete before the CTA retires
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

size_t conv_smem_bytes(int stages, int stage_bytes, int halo_total, int bres_bytes, int bias_floats, int chain_bytes = 0) {
  return 1024 /*alignment slack*/ + (size_t)stages * stage_bytes + (size_t)halo_total + (size_t)bres_bytes + (size_t)chain_bytes + 2 * 128 * 128 /*output staging*/ +
         (2 * kMaxStages + 2 * kMaxHaloStages + 7) * 8 + 8 + (size_t)bias_floats * 4 + 16;
}

}  // namespace

GtEncodeTiledFn conv_tc_encode() { return g_encode; }

// Layer-0 output: the destination is the full-resolution NHWC tensor [B][2Ho][2Wo][32] (a dedicated 32-channel buffer); one
// staging slab holds, for every super-pixel (Y, X) of the tile, the two pixels (2Y + oy, 2X + {0,1}) x 32 channels = 128
// contiguous bytes.  4-D view of the destination: {64 = (ox, c), oy 2, X Wo, Y' = n * Ho + Y}; box {64, 1, tw, th}; tiles never
// straddle images (Ho % th == 0).
CUresult encode_s2d_out(CUtensorMap* tm, CUtensorMapDataType dt, const View* out, int Wo, int Ho, int B, int tw, int th) {
  const cuuint64_t ps = (cuuint64_t)out->ctot * 2;   // one full-resolution pixel (64 bytes)
  cuuint64_t gdim[4] = {64, 2, (cuuint64_t)Wo, (cuuint64_t)Ho * B};
  cuuint64_t gstr[3] = {(cuuint64_t)2 * Wo * ps, 2 * ps, (cuuint64_t)4 * Wo * ps};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)tw, (cuuint32_t)th};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return g_encode(tm, dt, 4, (void*)(out->ptr + out->coff), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
int conv_tc_num_sms() { return g_num_sms; }

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
  GT_CUDA(e, cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  GT_CUDA(e, cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
  return conv_sw_init(e);
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
  if ((e->plan_variant == 1 || e->plan_variant == 2 || e->plan_variant == 6) && !a.pre && !a.chain_cout) return conv_sw_plan(e, op, a);   // (the half-resolution pre-activation add and the chained 1x1 conv live in the pixel-major epilogue only)
  const View& in = a.in;
  const int cin = a.cin, k = a.k, stride = a.stride;
  const int kbe = (a.kb_elems == 64 && cin == 32) ? 32 : a.kb_elems;   // 32-channel inputs: 64-byte rows instead of half-empty 128-byte rows
  GT_CHECK(e, in.C == cin, "conv plan: input view has %d channels, conv expects %d", in.C, cin);
  GT_CHECK(e, (in.ctot % 8) == 0 && (in.coff % 8) == 0, "conv plan: input slice must be 16-byte aligned");
  GT_CHECK(e, k >= 1 && k <= 3, "conv plan: k=%d unsupported", k);
  GT_CHECK(e, stride == 1 || stride == 2, "conv plan: stride=%d unsupported", stride);
  GT_CHECK(e, kbe == 64 || kbe == 32 || kbe == 16, "conv plan: kb_elems=%d unsupported", kbe);
  ConvParams& p = op->p;
  memset(&p, 0, sizeof(p));
  const int pad = a.pad >= 0 ? a.pad : k / 2;
  const int Ho = a.Ho > 0 ? a.Ho : (in.H + 2 * pad - k) / stride + 1, Wo = a.Wo > 0 ? a.Wo : (in.W + 2 * pad - k) / stride + 1;
  const int cout_total = a.cout;
  op->cin = cin; op->cout = cout_total; op->k = k; op->stride = stride;
  p.B = a.Bmax; p.H = Ho; p.W = Wo;
  pick_tile(Ho, Wo, &p.tw, &p.th);
  // Halo mode (stride-1 k x k): 8 x 16 output tiles; one TMA box per k-block brings the (16-pitch) x (16 + k - 1) halo and the
  // k*k taps are shifted shared-memory descriptors -> k*k times fewer A bytes from L2 and k*k times fewer TMA issues.
  // variants 4 / 5 (or GT_HALO=1): halo staging for stride-1 k >= 2 layers; the autotuner decides where it pays
  int want_halo = ((e->plan_variant == 4 || e->plan_variant == 5 || e->halo_mode) && stride == 1 && k >= 2) ? 1 : 0;
  if (want_halo) { p.tw = 8; p.th = 16; }
  if (a.out_s2d && !(want_halo && (Ho % 16) == 0)) { p.tw = 16; p.th = 8; want_halo = 0; }   // Ho is a multiple of th: tiles never straddle images in the folded row index
  p.tiles_x = ceil_div(Wo, p.tw); p.tiles_y = ceil_div(Ho, p.th);
  p.stride = stride; p.ksize = k; p.pad = pad;
  p.kb_elems = kbe;
  p.kc_blocks = ceil_div(cin, kbe);
  op->cin_pad = p.kc_blocks * kbe;
  p.num_kb = k * k * p.kc_blocks;
  const int cout16 = ceil_div(cout_total, 16) * 16;
  // N tile: whole output slabs of 128 bytes per row (64 16-bit or 32 fp32 columns), or one 64-byte slab for narrow outputs
  const int elem = a.out_f32 ? 4 : 2;
  const int slab_cols = 128 / elem;
  if (cout16 * elem <= 64 || (!a.out_f32 && cout16 <= 32)) {
    p.BN = 64 / elem;                       // 32 (16-bit) or 16 (fp32) columns: one 64-byte slab
    p.epi_mode = a.out_f32 ? 3 : 1;
  } else {
    p.BN = std::min(256, ceil_div(cout16, slab_cols) * slab_cols);
    p.epi_mode = a.out_f32 ? 2 : 0;
  }
  p.n_tiles = ceil_div(cout16, p.BN);
  op->cout_pad = p.n_tiles * p.BN;
  // two accumulators of BN fp32 columns each; the allocation is a power of two >= 32 columns
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.BN) p.tmem_cols *= 2;
  p.acc_stride = p.tmem_cols / 2;
  int chain_bytes = 0;
  if (a.chain_cout) {
    GT_CHECK(e, !a.out_f32 && !a.out_s2d && !a.res && !a.up && !a.pre && p.epi_mode == 0 && p.n_tiles == 1 && p.BN == cout_total && p.BN <= 128 &&
                    (a.chain_cout % 64) == 0 && a.chain_cout <= 128,
             "conv plan: a chained 1x1 conv needs a plain 16-bit conv with 64 or 128 outputs (got %d -> %d)", cout_total, a.chain_cout);
    p.chain_n = a.chain_cout; p.chain_act = getenv("GT_DEBUG_NOACT") ? 0 : a.chain_act;
    p.acc_stride = p.BN; p.chain_tmem = 2 * p.BN;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.BN + p.chain_n) p.tmem_cols *= 2;
    chain_bytes = (p.BN >> 6) * p.chain_n * 128;
  }
  const int a_bytes = 128 * kbe * 2, b_bytes = p.BN * kbe * 2;
  const int bres_bytes = p.num_kb * b_bytes;
  p.b_resident = (p.n_tiles == 1 && bres_bytes <= 112 * 1024 && (size_t)bres_bytes + chain_bytes + 3 * a_bytes + 40 * 1024 <= (size_t)e->conv_smem_kb * 1024) ? 1 : 0;
  const bool occ2 = (e->plan_variant == 3 || e->plan_variant == 5) && p.tmem_cols <= 256;   // two CTAs per SM (BN <= 128); else the plain plan
  const size_t budget = occ2 ? (size_t)112 * 1024 : (size_t)e->conv_smem_kb * 1024;
  if (occ2) {
    p.b_resident = (p.n_tiles == 1 && (size_t)bres_bytes + chain_bytes + 3 * a_bytes + 36 * 1024 <= budget) ? 1 : 0;
  }
  op->occ2 = occ2 ? 1 : 0;
  int stage_bytes = p.b_resident ? a_bytes : a_bytes + b_bytes;
  int halo_total = 0;
  if (want_halo) {
    const int hrows = p.th + k - 1;
    p.halo_tx = (p.tw + k - 1) * hrows * kbe * 2;
    p.halo_bytes = (p.halo_tx + 1023) / 1024 * 1024;
    const size_t fixed = conv_smem_bytes(0, 0, 0, p.b_resident ? bres_bytes : 0, op->cout_pad + p.chain_n, chain_bytes);
    int a_st, b_st = 0;
    if (p.b_resident) a_st = (int)((budget - fixed) / p.halo_bytes);
    else {
      a_st = 2;
      b_st = (int)((budget - fixed - (size_t)a_st * p.halo_bytes) / b_bytes);
      if (b_st > kMaxStages) { b_st = kMaxStages; a_st = (int)((budget - fixed - (size_t)b_st * b_bytes) / p.halo_bytes); }
    }
    if (a_st > kMaxHaloStages) a_st = kMaxHaloStages;
    if (a_st >= 2 && (p.b_resident || b_st >= 3)) {
      p.halo = want_halo; p.a_stages = a_st; p.stages = b_st; stage_bytes = b_bytes;
      halo_total = a_st * p.halo_bytes;
    } else {
      pick_tile(Ho, Wo, &p.tw, &p.th);
      p.halo_tx = p.halo_bytes = 0;
    }
  }
  p.tiles_x = ceil_div(Wo, p.tw); p.tiles_y = ceil_div(Ho, p.th);
  if (!p.halo) {
    const size_t fixed = conv_smem_bytes(0, stage_bytes, 0, p.b_resident ? bres_bytes : 0, op->cout_pad + p.chain_n, chain_bytes);
    int stages = (int)((budget - fixed) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    GT_CHECK(e, stages >= 2, "conv plan: tile does not fit shared memory (BN=%d)", p.BN);
    p.stages = stages;
  }
  p.cout = cout_total; p.act = getenv("GT_DEBUG_NOACT") ? 0 : a.act; p.fp16 = e->cfg.act_dtype == GT_ACT_FP16 ? 1 : 0;
  p.scale = a.scale;
  if (a.out_f32) {
    p.out_f32 = 1; p.out = a.out_f32; p.out_img_stride = a.out_img_stride; p.out_ctot = a.out_ctot_f32; p.out_coff = a.out_coff_f32;
  } else {
    const View* out = a.out;
    if (a.out_s2d) GT_CHECK(e, out && out->H == 2 * Ho && out->W == 2 * Wo && out->C * 4 == cout_total && cout_total == 128 && (Ho % p.th) == 0 &&
                                   out->ctot == 32 && out->coff == 0,
                            "conv plan: s2d output view mismatch");
    else
    GT_CHECK(e, out && out->H == Ho && out->W == Wo && out->C == (a.chain_cout ? a.chain_cout : cout_total), "conv plan: output view mismatch (%dx%dx%d vs %dx%dx%d)",
             out ? out->H : -1, out ? out->W : -1, out ? out->C : -1, Ho, Wo, a.chain_cout ? a.chain_cout : cout_total);
    GT_CHECK(e, (out->ctot % 8) == 0 && (out->coff % 8) == 0 && (cout_total % 8) == 0, "conv plan: output slice must be 16-byte aligned");
    p.out_f32 = 0; p.out = out->ptr; p.out_img_stride = (long long)Ho * Wo; p.out_ctot = out->ctot; p.out_coff = out->coff;
    p.out_s2d = a.out_s2d ? 1 : 0; p.out_rows = Ho;
  }
  if (a.res) {
    GT_CHECK(e, a.res->H == Ho && a.res->W == Wo && a.res->C == cout_total && !a.out_f32, "conv plan: residual view mismatch");
    p.res = a.res->ptr; p.res_ctot = a.res->ctot; p.res_coff = a.res->coff;
  }
  if (a.up) {
    GT_CHECK(e, a.up->H == 2 * Ho && a.up->W == 2 * Wo && a.up->C == cout_total && !a.out_f32, "conv plan: upsample view mismatch");
    p.up = a.up->ptr; p.up_ctot = a.up->ctot; p.up_coff = a.up->coff;
  }
  if (a.pre) {
    GT_CHECK(e, !a.res && !a.out_f32 && (Ho % 2) == 0 && (Wo % 2) == 0 && a.pre->H * 2 == Ho && a.pre->W * 2 == Wo && a.pre->C == cout_total &&
                    (a.pre->ctot % 8) == 0 && (a.pre->coff % 8) == 0, "conv plan: half-resolution partial-sum view mismatch");
    p.res = a.pre->ptr; p.res_ctot = a.pre->ctot; p.res_coff = a.pre->coff; p.res_pre = 1;
  }
  op->smem = conv_smem_bytes(p.stages, stage_bytes, halo_total, p.b_resident ? bres_bytes : 0, op->cout_pad + p.chain_n, chain_bytes);
  op->flops = 2.0 * Ho * Wo * (double)cout_total * cin * k * k + 2.0 * Ho * Wo * (double)a.chain_cout * cout_total;
  // algorithmic HBM bytes per image: input slice + output (+ residual, + upsampled copy) + weights (once per launch, ignored)
  op->bytes = (double)in.H * in.W * cin * 2 + (double)Ho * Wo * (a.chain_cout ? a.chain_cout : cout_total) * (a.out_f32 ? 4 : 2) * (a.up ? 5 : 1) +
              (a.res ? (double)Ho * Wo * cout_total * 2 : 0.0) + (a.pre ? (double)Ho * Wo * cout_total * 2 / 4 : 0.0);

  // weights + bias storage
  const size_t wn = (size_t)op->cout_pad * k * k * op->cin_pad;
  GT_TRY(e->dev_alloc((void**)&op->w_dev, wn * sizeof(bf16)));
  GT_TRY(e->dev_alloc((void**)&op->b_dev, (size_t)op->cout_pad * sizeof(float)));
  GT_CUDA(e, cudaMemset(op->w_dev, 0, wn * sizeof(bf16)));
  GT_CUDA(e, cudaMemset(op->b_dev, 0, (size_t)op->cout_pad * sizeof(float)));
  p.bias = op->b_dev;
  if (a.chain_cout) {
    GT_TRY(e->dev_alloc((void**)&op->w2_dev, (size_t)p.chain_n * p.BN * sizeof(bf16)));
    GT_TRY(e->dev_alloc((void**)&op->b2_dev, (size_t)p.chain_n * sizeof(float)));
    GT_CUDA(e, cudaMemset(op->w2_dev, 0, (size_t)p.chain_n * p.BN * sizeof(bf16)));
    GT_CUDA(e, cudaMemset(op->b2_dev, 0, (size_t)p.chain_n * sizeof(float)));
    p.chain_bias = op->b2_dev;
  }

  const CUtensorMapDataType dt = p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = kbe == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kbe == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  // A: NHWC input slice as a 4-D tensor {C, W, H, N}
  {
    cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)a.Bmax};
    cuuint64_t gstr[3] = {(cuuint64_t)in.ctot * 2, (cuuint64_t)in.W * in.ctot * 2, (cuuint64_t)in.H * in.W * in.ctot * 2};
    cuuint32_t box[4] = {(cuuint32_t)kbe, (cuuint32_t)(p.tw * stride), (cuuint32_t)(p.th * stride), 1};
    if (p.halo) { box[1] = (cuuint32_t)(p.tw + k - 1); box[2] = (cuuint32_t)(p.th + k - 1); }
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    // L2 promotion no wider than a box row: a 32-channel slice (64 B) of a wider pixel promoted to 128 B pulls the neighbouring
    // slice's bytes out of DRAM as well (GT_L2PROMO=128 restores the old setting)
    const CUtensorMapL2promotion promo = (kbe * 2 >= 128 || e->l2promo_128) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                        : (kbe * 2 >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
    CUresult r = g_encode(&op->tmA, dt, 4, (void*)(in.ptr + in.coff), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  memset(&op->tmC, 0, sizeof(op->tmC));
  if (a.chain_cout) {   // chained conv weights [chain_n][BN] (K-major): one box = one 64-channel k-block of all chain_n rows
    cuuint64_t gdim[2] = {(cuuint64_t)p.BN, (cuuint64_t)p.chain_n};
    cuuint64_t gstr[1] = {(cuuint64_t)p.BN * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p.chain_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&op->tmC, dt, 2, (void*)op->w2_dev, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(chain weights) failed: %d", (int)r);
  }
  // Output: destination slice as a 4-D tensor {C, W, H, N}; one box = one staging slab (128 rows x 128 or 64 bytes).
  // Out-of-range rows / columns / channels of ragged tiles are clipped by the TMA unit.
  {
    const int box_c = (p.epi_mode == 0) ? 64 : (p.epi_mode == 2 ? 32 : p.BN);
    const CUtensorMapSwizzle osw = (p.epi_mode == 0 || p.epi_mode == 2) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUtensorMapDataType odt = a.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt;
    auto enc = [&](CUtensorMap* tm, void* base, cuuint64_t W_, cuuint64_t H_, cuuint64_t pix_stride_b, cuuint64_t row_stride_b,
                   cuuint64_t img_stride_b) -> CUresult {
      cuuint64_t gdim[4] = {(cuuint64_t)(a.chain_cout ? a.chain_cout : cout_total), W_, H_, (cuuint64_t)a.Bmax};
      cuuint64_t gstr[3] = {pix_stride_b, row_stride_b, img_stride_b};
      cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)p.tw, (cuuint32_t)p.th, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      return g_encode(tm, odt, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, osw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r;
    if (a.out_f32) {
      const cuuint64_t ps = (cuuint64_t)a.out_ctot_f32 * 4;
      GT_CHECK(e, (ps % 16) == 0 && (a.out_coff_f32 % 4) == 0, "conv plan: fp32 output rows must be 16-byte aligned (row %d floats, col %d)",
               a.out_ctot_f32, a.out_coff_f32);
      r = enc(&op->tmOut, (void*)(a.out_f32 + a.out_coff_f32), Wo, Ho, ps, (cuuint64_t)Wo * ps, (cuuint64_t)a.out_img_stride * ps);
    } else if (a.out_s2d) {
      r = encode_s2d_out(&op->tmOut, odt, a.out, Wo, Ho, a.Bmax, p.tw, p.th);
    } else {
      const cuuint64_t ps = (cuuint64_t)a.out->ctot * 2;
      r = enc(&op->tmOut, (void*)(a.out->ptr + a.out->coff), Wo, Ho, ps, (cuuint64_t)Wo * ps, (cuuint64_t)Ho * Wo * ps);
    }
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(out) failed: %d (cout=%d box_c=%d)", (int)r, cout_total, box_c);
    memset(&op->tmUp, 0, sizeof(op->tmUp));
    if (a.up) {
      const cuuint64_t ps = (cuuint64_t)a.up->ctot * 2, W2 = (cuuint64_t)2 * Wo, H2 = (cuuint64_t)2 * Ho;
      for (int d = 0; d < 4; ++d) {
        bf16* base = a.up->ptr + a.up->coff + ((size_t)(d >> 1) * W2 + (d & 1)) * a.up->ctot;
        r = enc(&op->tmUp.m[d], (void*)base, Wo, Ho, 2 * ps, 2 * W2 * ps, H2 * W2 * ps);
        GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(up %d) failed: %d", d, (int)r);
      }
    }
  }
  return GT_OK;
}

// w[s]: f32 [couts[s]][cin][k][k] (PyTorch layout); several convs reading the same input are stacked along cout
int conv_tc_pack_weights(gt_engine* e, ConvOp* op, const float* const* w, const float* const* b, const int* couts, int n) {
  const int taps = op->k * op->k;
  const size_t wn = (size_t)op->cout_pad * taps * op->cin_pad;
  const int fp16 = op->p.fp16;
  const int cin_total = op->w_cin_total > 0 ? op->w_cin_total : op->cin, cin_off = op->w_cin_total > 0 ? op->w_cin_off : 0;   // input-channel slice of the canonical conv
  std::vector<uint16_t> hw(wn, 0);
  std::vector<float> hb(op->cout_pad, 0.f);
  int co0 = 0;
  for (int s = 0; s < n; ++s) {
    for (int co = 0; co < couts[s]; ++co) {
      for (int ci = 0; ci < op->cin; ++ci)
        for (int t = 0; t < taps; ++t)
          hw[((size_t)(co0 + co) * taps + t) * op->cin_pad + ci] = host_to_act(w[s][((size_t)co * cin_total + cin_off + ci) * taps + t], fp16);
      hb[co0 + co] = (b[s] && !op->no_bias) ? b[s][co] : 0.f;
    }
    co0 += couts[s];
  }
  GT_CHECK(e, co0 == op->cout, "pack weights: cout mismatch %d vs %d", co0, op->cout);
  return conv_tc_upload_packed(e, op, hw.data(), hb.data());
}

// chained 1x1 conv: w f32 [chain_n][cin2 = this conv's cout] (PyTorch [cout][cin][1][1]) -> 16-bit [chain_n][BN]; bias f32 [chain_n]
int conv_tc_pack_chain(gt_engine* e, ConvOp* op, const float* w, const float* b) {
  const int n = op->p.chain_n, K = op->p.BN;
  GT_CHECK(e, n > 0 && op->w2_dev && K == op->cout, "pack chain weights: the op has no chained conv");
  std::vector<uint16_t> hw((size_t)n * K, 0);
  std::vector<float> hb(n, 0.f);
  for (int co = 0; co < n; ++co) {
    for (int ci = 0; ci < K; ++ci) hw[(size_t)co * K + ci] = host_to_act(w[(size_t)co * K + ci], op->p.fp16);
    hb[co] = b ? b[co] : 0.f;
  }
  GT_CUDA(e, cudaMemcpy(op->w2_dev, hw.data(), hw.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  GT_CUDA(e, cudaMemcpy(op->b2_dev, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
  return GT_OK;
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
  if (op->swapped) return conv_sw_launch_range(e, op, b0, nb, st);
  ConvParams p = op->p;
  p.B = nb;
  p.img0 = b0;
  p.total_tiles = p.tiles_x * p.tiles_y * nb * p.n_tiles;
  const int slots = g_num_sms * (op->occ2 ? 2 : 1);
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = op->smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = e->pdl ? 1 : 0;
  if (op->occ2) GT_CUDA(e, cudaLaunchKernelEx(&cfg, conv_tc_kernel<2>, op->tmA, op->tmB, op->tmOut, op->tmUp, op->tmC, p));
  else GT_CUDA(e, cudaLaunchKernelEx(&cfg, conv_tc_kernel<1>, op->tmA, op->tmB, op->tmOut, op->tmUp, op->tmC, p));
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}
