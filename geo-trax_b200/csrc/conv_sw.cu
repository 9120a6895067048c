// Swapped-operand tcgen05 implicit-GEMM convolution (sm_100a): D[cout][pixels] = W[cout][K] . X[pixels][K]^T.
//
// Why: a 256-pixel tile as the GEMM N dimension halves the instruction count and the weight re-reads of the 3x3 layers with
// cout >= 128 (one M128 x N256 x K16 tcgen05.mma = 135 cycles = 95 % of the pipe's nominal rate, tools/mma_bench.cu); the
// WEIGHTS are the A operand (M = 128 output channels, zero rows above cout), the pixels the B operand; cout > 128 runs
// ceil(cout / 128) M tiles.  (The first motivation written here -- "an MMA costs 130-150 cycles whatever N" -- was a measurement of
// the old issue path, not of the pipe: profiles/round1_summary.md section 3.)  The per-layer autotuner picks this kernel for 17-24 of the
// 60 launches, the pixel-major kernel (one or two CTAs per SM, with or without halo staging) for the rest.
//
// Same skeleton as conv_tc.cu: persistent CTA per SM, warps 0 and 10 TMA producers, warp 1 MMA issuer (elect.sync), warps 2..9
// epilogue, shared-memory operand ring across tiles, two TMEM accumulators (2 x 256 columns = all of TMEM), weights resident
// in shared memory when they fit.  The accumulator arrives transposed (TMEM lane = output channel, column = pixel), so the
// epilogue thread owns ONE channel: bias is a register, and each value is written as a 2-byte (or 4-byte, fp32 head rows)
// element into a [pixel][channel] staging granule in the TMA swizzle layout, stored with bulk tensor stores:
//   16-bit: granule = 128 pixels x 64 channels (16 KB), filled by the two warps of a channel half; 4 granules per tile
//   fp32  : granule = 32 pixels x 32 channels (4 KB), private to a warp, double-buffered
//   cout <= 64: two-phase transposed epilogue (raw accumulators -> smem -> all warps finish them with thread = pixel), see below
// Variant 2 (p.halo): one (8 + k - 1) x (32 + k - 1) pixel box per k-block, the k*k taps are row-shifted descriptors into it.
#include <algorithm>

#include "engine.cuh"
#include "tc_ptx.cuh"

#ifdef GT_SW_TIMING
__device__ unsigned long long g_sw_dbg[16 * 8];
extern "C" int gt_debug_sw_timing(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_sw_dbg, sizeof(g_sw_dbg)); unsigned long long z[128] = {}; cudaMemcpyToSymbol(g_sw_dbg, z, sizeof(z)); return 0; }
__device__ __forceinline__ long long clk_now() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
#define TCK(v) const long long v = clk_now()
#else
#define TCK(v)
#endif
namespace {

constexpr int kSwThreads = 64 + kEpiWarps * 32 + 32;  // 352: warp 0 + warp 10 TMA producers, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int kSwProd2 = kSwThreads / 32 - 1;         // the second producer warp (per-tap path only): boxes of alternate k-blocks
constexpr int kSwMaxStages = 8;
constexpr int kPx = 256;        // pixels per tile (GEMM N)
constexpr int kCo = 128;        // output channels per tile (GEMM M)
constexpr int kStagingBytes = 64 * 1024;
constexpr int kSwWStages = 4;   // weight ring depth of the halo variant

__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ uint32_t to_act16(float f, int fp16) {
  if (fp16) { __half h = __float2half_rn(f); return (uint32_t)*reinterpret_cast<unsigned short*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  return (uint32_t)*reinterpret_cast<unsigned short*>(&h);
}
__device__ __forceinline__ float from_act16(unsigned short u, int fp16) {
  if (fp16) return __half2float(*reinterpret_cast<__half*>(&u));
  return __uint_as_float((uint32_t)u << 16);
}

// ---- MMA issue helpers (see the MMA role in the kernel) -----------------------------------------------------------------------
struct MmaCtx {
  uint32_t full_bar, empty_bar, tfull_bar, tempty_bar, wfull_bar, wempty_bar;   // shared addresses of the barrier arrays
  uint32_t smem_lo, wres_lo;        // descriptor low words (address >> 4) of ring stage 0 and of the weight region
  uint32_t stage16, w16, x16, row16;  // byte sizes >> 4
  uint32_t hi_std, hi_halo;         // descriptor high words: SBO = 8 rows / SBO = halo row pitch
  uint32_t tmem_base, idesc;
};
template <int MPK, int PAIR>
__device__ __forceinline__ void mma_role(const ConvParams& p, const MmaCtx& c) {
  uint32_t s = 0, ph = 0, li = 0, s_lo = c.smem_lo;
  const int num_kb = p.num_kb;
  const bool res = p.b_resident != 0;
  const int first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  for (int tile = first; tile < p.total_tiles; tile += step, ++li) {
    const uint32_t as = li & 1u;
    mbar_wait(c.tempty_bar + as * 8u, ((li >> 1) & 1u) ^ 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tacc = c.tmem_base + as * (uint32_t)kPx;
    uint32_t w_lo = c.wres_lo;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(c.full_bar + s * 8u, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        if constexpr (PAIR) {   // one M = 256 MMA over both CTAs' operands; the stage is released in both CTAs
          issue_kb_pair<MPK>(tacc, s_lo + c.x16, c.hi_std, s_lo, c.hi_std, c.idesc, kb ? 1u : 0u);
          umma_commit_pair(c.empty_bar + s * 8u);
        } else {
          issue_kb<MPK>(tacc, res ? w_lo : s_lo + c.x16, c.hi_std, s_lo, c.hi_std, c.idesc, kb ? 1u : 0u);
          umma_commit(c.empty_bar + s * 8u);
        }
      }
      __syncwarp();
      w_lo += c.w16;
      s_lo += c.stage16;
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; s_lo = c.smem_lo; }
    }
    if (elect_one()) {
      if constexpr (PAIR) umma_commit_pair(c.tfull_bar + as * 8u);
      else umma_commit(c.tfull_bar + as * 8u);
    }
    __syncwarp();
  }
}

// halo variant: one pixel box per (tile, k-block); tap (dy, dx) = the same box read from row dy * halo_w + dx, its 8-row groups
// halo_w rows apart (hi_halo); weights resident ([tap][kc] k-blocks) or through the weight ring (one stage per tap)
template <int KS, int MPK>
__device__ __forceinline__ void mma_role_halo(const ConvParams& p, const MmaCtx& c) {
  uint32_t s = 0, ph = 0, li = 0, ws = 0, wph = 0, s_lo = c.smem_lo;
  const uint32_t halo_w16 = (uint32_t)(p.tw + KS - 1) * c.row16;
  const bool res = p.b_resident != 0;
  const uint32_t wtap16 = c.w16 * (uint32_t)p.kc_blocks;   // resident weights: distance between the taps of one kc
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li) {
    const uint32_t as = li & 1u;
    mbar_wait(c.tempty_bar + as * 8u, ((li >> 1) & 1u) ^ 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tacc = c.tmem_base + as * (uint32_t)kPx;
    for (int kc = 0; kc < p.kc_blocks; ++kc) {
      mbar_wait(c.full_bar + s * 8u, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (res) {
        if (elect_one()) {
          uint32_t w_lo = c.wres_lo + (uint32_t)kc * c.w16;
#pragma unroll
          for (int dy = 0; dy < KS; ++dy)
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) {
              issue_kb<MPK>(tacc, w_lo, c.hi_std, s_lo + (uint32_t)dy * halo_w16 + (uint32_t)dx * c.row16, c.hi_halo, c.idesc, (dy | dx | kc) ? 1u : 0u);
              w_lo += wtap16;
            }
          umma_commit(c.empty_bar + s * 8u);
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int dy = 0; dy < KS; ++dy)
#pragma unroll
          for (int dx = 0; dx < KS; ++dx) {
            mbar_wait(c.wfull_bar + ws * 8u, wph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
              issue_kb<MPK>(tacc, c.wres_lo + ws * c.w16, c.hi_std, s_lo + (uint32_t)dy * halo_w16 + (uint32_t)dx * c.row16, c.hi_halo, c.idesc, (dy | dx | kc) ? 1u : 0u);
              umma_commit(c.wempty_bar + ws * 8u);
              if (dy == KS - 1 && dx == KS - 1) umma_commit(c.empty_bar + s * 8u);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.a_stages) { ws = 0; wph ^= 1u; }
          }
      }
      s_lo += c.stage16;
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; s_lo = c.smem_lo; }
    }
    if (elect_one()) umma_commit(c.tfull_bar + as * 8u);
    __syncwarp();
  }
}

// PAIR = 1 (variant 6, cout a multiple of 256, per-tap path): the kernel runs as clusters of two CTAs on one TPC.  A pair owns a
// 256-channel x 256-pixel tile: each CTA stages its own 128 weight rows and its own half (th / 2 image rows) of the pixel tile, the
// leader (cluster rank 0) issues M = 256 cta_group::2 MMAs that read both CTAs' shared memory, and each CTA's TMEM receives the
// accumulator rows of its own 128 channels for all 256 pixels -- so the epilogue is the single-CTA one.  A stage is 32 KB instead of
// 48 KB for the same MMA time: five or six ring stages, i.e. ~1.5x the load latency covered.
template <int PAIR>
__global__ void __launch_bounds__(kSwThreads, 1) conv_sw_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                                                                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvUpMaps tmUp,
                                                                const __grid_constant__ CUtensorMap tmRes, const ConvParams p) {
  pdl_launch_dependents();   // the next layer's CTAs may start their prologue as SMs free up
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int row_bytes = p.kb_elems * 2;
  const int w_bytes = kCo * row_bytes;          // A: 128 weight rows of one k-block
  const int x_bytes = (PAIR ? kPx / 2 : kPx) * row_bytes;   // B: 256 pixel rows of one k-block (pair: this CTA's 128)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int tile_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kCoTile = PAIR ? 2 * kCo : kCo;  // output channels per (pair) tile
  // halo mode (3x3 / 2x2 stride-1 layers, 8 x 32 pixel tiles): a ring stage holds ONE (tw + k - 1) x (th + k - 1) pixel box of a
  // k-block; the k*k taps are row-shifted UMMA descriptors into it (the swizzle XOR acts on absolute smem address bits, so any
  // whole-row shift and any row-multiple SBO address the bytes TMA wrote).  Weights are resident or flow through their own ring.
  const int stage_bytes = p.halo ? p.halo_bytes : (p.b_resident ? x_bytes : x_bytes + w_bytes);
  uint8_t* w_res = smem + (size_t)p.stages * stage_bytes;   // resident weights, or (halo, not resident) the weight ring
  uint8_t* staging = w_res + (p.b_resident ? (size_t)p.num_kb * w_bytes : (p.halo ? (size_t)p.a_stages * w_bytes : 0));
  uint8_t* tail = staging + kStagingBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kSwMaxStages;
  uint64_t* tfull_bar = empty_bar + kSwMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* wres_bar = tempty_bar + 2;
  uint64_t* res_bar = wres_bar + 1;               // [4] residual granule landed (one per 16-bit staging granule)
  uint64_t* wfull_bar = res_bar + 4;              // [kSwWStages] halo mode, weights not resident: weight ring
  uint64_t* wempty_bar = wfull_bar + kSwWStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(wempty_bar + kSwWStages);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem + 2) + 15) & ~(uintptr_t)15);   // 16-byte aligned (lds.128)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.total_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), PAIR ? 2 * kEpiWarps : kEpiWarps);   // pair: the leader's barrier collects both CTAs' epilogue warps
    }
    mbar_init(smem_u32(wres_bar), 1);
    for (int s = 0; s < 4; ++s) mbar_init(smem_u32(&res_bar[s]), 1);
    for (int s = 0; s < kSwWStages; ++s) {
      mbar_init(smem_u32(&wfull_bar[s]), 1);
      mbar_init(smem_u32(&wempty_bar[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {   // warp 1 of both CTAs, same shared-memory offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (warp >= 2 && warp < kSwProd2) {
    const int nb = p.n_tiles * kCo;   // (pair: s_bias keeps only this CTA's 128 channels of every 256-channel tile)
    for (int i = threadIdx.x - 64; i < nb; i += kEpiWarps * 32) s_bias[i] = PAIR ? p.bias[(i >> 7) * kCoTile + (int)rank * kCo + (i & 127)] : p.bias[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (PAIR) { __syncthreads(); cluster_sync_all(); }   // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0 || (warp == kSwProd2 && !p.halo)) {
    // ===== TMA producers (converged warps, elected lane issues).  The per-tap path is bound by how fast boxes are issued (~3.4 cycles
    // per box row with one issuing warp, fewer with two CTAs), so two warps issue the boxes of alternate k-blocks. =====
    const uint32_t pid = warp == 0 ? 0u : 1u;     // (halo path: warp 0 only)
    if (p.b_resident && pid == 0u) {
      if (elect_one()) {
        const uint32_t bb = smem_u32(wres_bar);
        mbar_expect_tx(bb, (uint32_t)(p.num_kb * w_bytes));
        for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(smem_u32(w_res + (size_t)kb * w_bytes), &tmW, bb, kb * p.kb_elems, 0);
      }
      __syncwarp();
    }
    pdl_wait();   // everything above touched only this layer's constants; activations of the previous layer are read below
    const uint32_t tx_bytes = (uint32_t)stage_bytes;
    uint32_t s = 0, ph = 0, ws = 0, wph = 0, gk = 0;
    TileIter ti;
    ti.init(p, tile_first, tile_step);
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ti.next(p)) {
      const TileCoord t = ti.coord(p);
      const int m0 = ti.nt * kCoTile + (int)rank * kCo;
      const int cx = t.x0 * p.stride - p.pad, cy = (t.y0 + (PAIR ? (int)rank * (p.th >> 1) : 0)) * p.stride - p.pad;
      if (p.halo) {
        const int taps = p.ksize * p.ksize;
        for (int kc = 0; kc < p.kc_blocks; ++kc) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          if (elect_one()) {
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, (uint32_t)p.halo_tx);
            tma_load_4d(smem_u32(smem + (size_t)s * stage_bytes), &tmX, fb, kc * p.kb_elems, cx, cy, t.n);
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
          if (!p.b_resident) {
            for (int tap = 0; tap < taps; ++tap) {
              mbar_wait(smem_u32(&wempty_bar[ws]), wph ^ 1u);
              if (elect_one()) {
                const uint32_t fb = smem_u32(&wfull_bar[ws]);
                mbar_expect_tx(fb, (uint32_t)w_bytes);
                tma_load_2d(smem_u32(w_res + (size_t)ws * w_bytes), &tmW, fb, (tap * p.kc_blocks + kc) * p.kb_elems, m0);
              }
              __syncwarp();
              if (++ws == (uint32_t)p.a_stages) { ws = 0; wph ^= 1u; }
            }
          }
        }
        continue;
      }
      int kb = 0;
      for (int dy = 0; dy < p.ksize; ++dy)
        for (int dx = 0; dx < p.ksize; ++dx)
          for (int kc = 0; kc < p.kc_blocks; ++kc, ++kb, ++gk) {
            if ((gk & 1u) == pid) {
              mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
              if (elect_one()) {
                const uint32_t fb = smem_u32(&full_bar[s]);
                uint8_t* sx = smem + (size_t)s * stage_bytes;
                if constexpr (PAIR) {   // both CTAs' boxes are counted on the leader's barrier (the MMA reads both CTAs' stage s)
                  const uint32_t fb0 = mapa_rank(fb, 0u);
                  if (rank == 0u) mbar_expect_tx(fb, 2u * tx_bytes);
                  tma_load_4d_pair(smem_u32(sx), &tmX, fb0, kc * p.kb_elems, cx + dx, cy + dy, t.n);
                  tma_load_2d_pair(smem_u32(sx + x_bytes), &tmW, fb0, kb * p.kb_elems, m0);
                } else {
                  mbar_expect_tx(fb, tx_bytes);
                  tma_load_4d(smem_u32(sx), &tmX, fb, kc * p.kb_elems, cx + dx, cy + dy, t.n);
                  if (!p.b_resident) tma_load_2d(smem_u32(sx + x_bytes), &tmW, fb, kb * p.kb_elems, m0);
                }
              }
              __syncwarp();
            }
            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
          }
    }
  } else if (warp == kSwProd2) {
    // halo path: the second producer warp has nothing to do
  } else if (warp == 1 && rank != 0u) {
    // pair: only the leader CTA issues MMAs
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The issue loop is written for the single issuing lane's latency: descriptors are (constant high word, running 32-bit low
    // word) pairs, the MMAs of a k-block are unrolled at compile time (MPK = k-block elements / 16), and nothing but two adds
    // separates consecutive tcgen05.mma.  (With general 64-bit descriptor code the issue block cost ~360 cycles per k-block --
    // more than the 2 x 135 cycles its MMAs take on the tensor pipe -- and the pipe idled half of the time.)
    const uint32_t fmt = p.fp16 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kPx >> 3) << 17) | ((uint32_t)(kCoTile >> 4) << 24);
    if (p.b_resident) mbar_wait(smem_u32(wres_bar), 0);
    MmaCtx c;
    c.full_bar = smem_u32(full_bar); c.empty_bar = smem_u32(empty_bar); c.tfull_bar = smem_u32(tfull_bar); c.tempty_bar = smem_u32(tempty_bar);
    c.wfull_bar = smem_u32(wfull_bar); c.wempty_bar = smem_u32(wempty_bar);
    c.smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4; c.wres_lo = (smem_u32(w_res) & 0x3FFFFu) >> 4;
    c.stage16 = (uint32_t)stage_bytes >> 4; c.w16 = (uint32_t)w_bytes >> 4; c.x16 = (uint32_t)x_bytes >> 4; c.row16 = (uint32_t)row_bytes >> 4;
    c.tmem_base = tmem_base; c.idesc = idesc;
    const uint32_t layout = p.kb_elems == 64 ? 2u : (p.kb_elems == 32 ? 4u : 6u);
    c.hi_std = desc_hi(8u * (uint32_t)row_bytes, layout);
    c.hi_halo = desc_hi((uint32_t)((p.tw + p.ksize - 1) * row_bytes), layout);
    const int mpk = p.kb_elems >> 4;
    if constexpr (PAIR) {
      if (mpk == 4) mma_role<4, 1>(p, c); else if (mpk == 2) mma_role<2, 1>(p, c); else mma_role<1, 1>(p, c);
    } else if (p.halo) {
      if (p.ksize == 3) { if (mpk == 4) mma_role_halo<3, 4>(p, c); else if (mpk == 2) mma_role_halo<3, 2>(p, c); else mma_role_halo<3, 1>(p, c); }
      else { if (mpk == 4) mma_role_halo<2, 4>(p, c); else if (mpk == 2) mma_role_halo<2, 2>(p, c); else mma_role_halo<2, 1>(p, c); }
    } else {
      if (mpk == 4) mma_role<4, 0>(p, c); else if (mpk == 2) mma_role<2, 0>(p, c); else mma_role<1, 0>(p, c);
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;             // TMEM lane quarter = channels [32q, 32q + 32) of the cout tile
    const int h = (warp - 2) >> 2;      // pixel half: accumulator columns [128h, 128h + 128)
    const int g = q >> 1;               // 16-bit staging granule (channel half) this warp fills together with its neighbour quarter
    const int cg = (q & 1) * 32 + lane; // channel inside the granule
    const bool gran_leader = ((q & 1) == 0) && lane == 0;
    const int half_rows = p.th >> 1;    // image rows per pixel half
    const int tw_shift = __ffs(p.tw) - 1, tw_mask = p.tw - 1;
    const uint32_t stg = smem_u32(staging);
    pdl_wait();   // residual reads and output stores below
    uint32_t li = 0, nstore = 0, res_uses = 0;
    // accumulator `as` has been read by this warp (pair: counted on the leader's barrier, the leader's MMAs overwrite both CTAs' TMEM)
    auto release_acc = [&](uint32_t as) {
      if constexpr (PAIR) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty_bar[as]), 0u));
      else mbar_arrive(smem_u32(&tempty_bar[as]));
    };
    TileIter ti;
    ti.init(p, tile_first, tile_step);
#ifdef GT_SW_TIMING
    long long d_pre = 0, d_wait = 0, d_a = 0, d_b1 = 0, d_b = 0, d_b2 = 0;
#endif
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++li, ti.next(p)) {
      TCK(c0);
      const uint32_t as = li & 1u;
      const TileCoord t = ti.coord(p);
      const int m0 = ti.nt * kCoTile + (int)rank * kCo;
      const int ch = m0 + q * 32 + lane;                 // this thread's output channel
      const bool warp_active = (m0 + q * 32) < p.cout;   // any valid channel in this lane quarter
      const float bias = s_bias[PAIR ? ti.nt * kCo + q * 32 + lane : min(ch, p.n_tiles * kCo - 1)];
      const int ybase = t.y0 + h * half_rows;            // first image row of this pixel half
      const bool xpose = !p.out_f32 && p.cout <= 64;   // small cout: two-phase epilogue through a shared-memory transpose (below)
      const bool gran_active = !p.out_f32 && !xpose && (m0 + g * 64) < p.cout;
      const uint32_t buf = stg + (uint32_t)((h * 2 + g) * 16384);
      if (xpose && q == 0 && lane == 0) {
        if (li > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (p.res) {
          const uint32_t rb = smem_u32(&res_bar[h * 2]);
          mbar_expect_tx(rb, 16384u);
          tma_load_4d(stg + (uint32_t)(h * 2 * 16384), &tmRes, rb, 0, t.x0, ybase, t.n);
        }
      }
      if (gran_active) {
        // the granule is free once last tile's bulk stores have read it; then (residual layers) the residual tile is fetched
        // straight into the granule while the MMAs of this tile are still running
        if (gran_leader) {
          if (li > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (p.res) {
            const uint32_t rb = smem_u32(&res_bar[h * 2 + g]);
            mbar_expect_tx(rb, 16384u);
            tma_load_4d(buf, &tmRes, rb, m0 + g * 64, t.x0, ybase, t.n);
          }
        }
      }
      TCK(c1);
      if (lane == 0) mbar_wait(smem_u32(&tfull_bar[as]), (li >> 1) & 1u);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      TCK(c2);
#ifdef GT_SW_TIMING
      d_pre += c1 - c0; d_wait += c2 - c1;
#endif
#if defined(GT_SW_EXP) && GT_SW_EXP == 1   // debug experiment: no epilogue at all (accumulator released at once)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) release_acc(as);
      continue;
#endif
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)kPx + (uint32_t)(h * 128);
      if (xpose) {
        // ---- 16-bit, cout <= 64: only cout / 32 lane quarters of the accumulator hold channels, and a TMEM lane quarter can only
        // be read by the warps of ONE scheduler (warp % 4), so a direct epilogue does all the SiLU work (2 MUFU per value, 4 lanes /
        // clk / scheduler) on one or two schedulers.  Instead the owning warps only move raw fp32 accumulators to shared memory
        // (T[channel][pixel], 16-byte chunks XOR-swizzled by channel), and all four warps of the pixel half then run bias + SiLU
        // (+ residual) with thread = pixel, packing 8 channels per 16-byte store into the TMA staging granule (h, 0).
        // T lives in granule (h, 1), which these layers never store from: 16 KB = 128 pixels x 32 channels or 64 x 64.
        const int steps = p.cout <= 32 ? 1 : 2;
        const int px = 128 / steps;                        // pixels per step
        const uint32_t G = stg + (uint32_t)(h * 2 * 16384);
        const uint32_t T = G + 16384u;
        const uint32_t trow_b = (uint32_t)px * 4u;         // bytes of one channel row of T
        const int team_tid = q * 32 + lane;
        const int pp = team_tid & (px - 1);                // this thread's pixel inside the step
        const int cb = (team_tid / px) * 32;               // ... and its 32 channels
#pragma unroll 1
        for (int st = 0; st < steps; ++st) {
          if (warp_active) {
            const uint32_t tw_row = T + (uint32_t)(q * 32 + lane) * trow_b;
            const int nch = px >> 5;
#pragma unroll 1
            for (int c = 0; c < nch; ++c) {
              uint32_t v[32];
              tmem_ld_x32(trow + (uint32_t)(st * px + c * 32), v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts_u4(tw_row + (uint32_t)((((c * 8 + j) ^ (lane & 7))) << 4), make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            }
          }
          if (st == steps - 1) {                            // accumulator fully read by this warp (or never needed)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) release_acc(as);
          }
          TCK(c3);
          named_bar(6 + h, 128);                            // T complete; the leader's wait_group.read (granule free) is visible
          TCK(c4);
          if (st == 0 && p.res) { mbar_wait(smem_u32(&res_bar[h * 2]), res_uses & 1u); ++res_uses; }
          const int r = st * px + pp;                       // granule row = pixel of this half
          const uint32_t trd = T + (uint32_t)(((pp >> 2) << 4) + (pp & 3) * 4);
          // all 32 values of the thread are loaded first and run through SiLU as independent chains (the epilogue is bound by
          // MUFU / fixed-latency stalls, so instruction-level parallelism across the four 8-channel groups matters)
          float f[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            uint32_t u;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"((trd ^ (uint32_t)((k & 7) << 4)) + (uint32_t)(cb + k) * trow_b) : "memory");
            f[k] = __uint_as_float(u);
          }
#pragma unroll
          for (int g8 = 0; g8 < 8; ++g8) {
            const float4 b4 = lds_f4(smem_u32(s_bias + cb + g8 * 4));
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) f[g8 * 4 + k] = fmaf(f[g8 * 4 + k], p.scale, bb[k]);
          }
          act_inplace<32>(f, p.act);
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int c0 = cb + g8 * 8;
            const uint32_t ga = G + (uint32_t)r * 128u + (uint32_t)((((c0 >> 3) ^ (r & 7))) << 4);
            float* fg = f + g8 * 8;
            if (p.res) {
              uint4 rv;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w) : "r"(ga));
              const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 d = unpack2_act(rr[k], p.fp16);
                fg[2 * k] += d.x; fg[2 * k + 1] += d.y;
              }
            }
            uint4 o;
            o.x = pack2_act(fg[0], fg[1], p.fp16); o.y = pack2_act(fg[2], fg[3], p.fp16);
            o.z = pack2_act(fg[4], fg[5], p.fp16); o.w = pack2_act(fg[6], fg[7], p.fp16);
            sts_u4(ga, o);
          }
          TCK(c5);
          if (st == steps - 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          named_bar(6 + h, 128);                            // T may be overwritten; the granule is complete after the last step
          TCK(c6);
#ifdef GT_SW_TIMING
          d_a += c3 - (st == 0 ? c2 : c2); d_b1 += c4 - c3; d_b += c5 - c4; d_b2 += c6 - c5;
#endif
        }
        if (q == 0 && lane == 0) {
          tma_store_4d(&tmOut, G, 0, t.x0, ybase, t.n);
          if (p.up) {
#pragma unroll
            for (int d = 0; d < 4; ++d) tma_store_4d(tmUp.m + d, G, 0, t.x0, ybase, t.n);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (!p.out_f32) {
        // ---- 16-bit: granule (h, g) = 128 pixels x 64 channels, shared by quarters 2g and 2g+1 ----
        if (gran_active) {
          named_bar(2 + h * 2 + g, 64);                  // leader's wait_group.read is visible to the neighbour warp
          if (p.res) { mbar_wait(smem_u32(&res_bar[h * 2 + g]), res_uses & 1u); ++res_uses; }
        }
        // staging address = row * 128 + ((chunk ^ (row & 7)) << 4) + (cg & 7) * 2; rows advance by compile-time steps
        uint32_t base_k[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) base_k[k] = buf + (uint32_t)((((cg >> 3) ^ k) << 4) + (cg & 7) * 2);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          if (warp_active) {
            tmem_ld_x32(trow + (uint32_t)(c * 32), v);
            tmem_ld_wait();
          }
          if (c == 3) {                                   // accumulator fully read by this warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) release_acc(as);
          }
          if (!warp_active) continue;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaf(__uint_as_float(v[i]), p.scale, bias);
          act_inplace<32>(f, p.act);
          const uint32_t rowoff = (uint32_t)(c * 32) * 128u;
          if (p.res) {   // residual values sit in the granule at the very addresses this thread is about to overwrite
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              unsigned short rv;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(rv) : "r"(base_k[i & 7] + rowoff + (uint32_t)i * 128u));
              f[i] += from_act16(rv, p.fp16);
            }
          }
          if (p.fp16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sts_u16(base_k[i & 7] + rowoff + (uint32_t)i * 128u, to_act16(f[i], 1));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) sts_u16(base_k[i & 7] + rowoff + (uint32_t)i * 128u, to_act16(f[i], 0));
          }
        }
        if (gran_active) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          named_bar(2 + h * 2 + g, 64);
          if (gran_leader && !(p.out_s2d && ybase >= p.H)) {   // (s2d rows are folded over images: a fully out-of-range granule must not wrap)
            const int cch = m0 + g * 64;
            if (p.out_s2d) tma_store_4d(&tmOut, buf, 0, g, t.x0, t.n * p.out_rows + ybase);
            else tma_store_4d(&tmOut, buf, cch, t.x0, ybase, t.n);
            if (p.up) {
#pragma unroll
              for (int d = 0; d < 4; ++d) tma_store_4d(tmUp.m + d, buf, cch, t.x0, ybase, t.n);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
        // ---- fp32 head rows: per-warp granules of 32 pixels x 32 channels (128-byte rows), double-buffered ----
        const int rows_per_chunk = 32 / p.tw;             // image rows covered by 32 pixels (tw <= 32 for fp32 outputs)
        const uint32_t wbuf = stg + (uint32_t)((warp - 2) * 8192);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          if (warp_active) {
            tmem_ld_x32(trow + (uint32_t)(c * 32), v);
            tmem_ld_wait();
          }
          if (c == 3) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) release_acc(as);
          }
          if (!warp_active) continue;
          const uint32_t buf = wbuf + (nstore & 1u) * 4096u;
          if (lane == 0 && nstore >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          // row i (pixel), 4-byte element `lane`: chunk = lane >> 2, swizzled with row & 7
#pragma unroll
          float fo[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) fo[i] = fmaf(__uint_as_float(v[i]), p.scale, bias);
          act_inplace<32>(fo, p.act);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            sts_u32(buf + (uint32_t)i * 128u + (uint32_t)((((lane >> 2) ^ (i & 7)) << 4) + (lane & 3) * 4), __float_as_uint(fo[i]));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmOut, buf, m0 + q * 32, t.x0, ybase + c * rows_per_chunk, t.n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++nstore;
        }
      }
    }
#ifdef GT_SW_TIMING
    if (lane == 0) {
      unsigned long long* d = g_sw_dbg + (warp) * 8;
      atomicAdd(d + 0, (unsigned long long)d_pre); atomicAdd(d + 1, (unsigned long long)d_wait); atomicAdd(d + 2, (unsigned long long)d_a);
      atomicAdd(d + 3, (unsigned long long)d_b1); atomicAdd(d + 4, (unsigned long long)d_b); atomicAdd(d + 5, (unsigned long long)d_b2);
      atomicAdd(d + 6, (unsigned long long)li);
    }
#endif
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

size_t sw_smem_bytes(int stages, int stage_bytes, int wres_bytes, int bias_floats) {
  return 1024 + (size_t)stages * stage_bytes + (size_t)wres_bytes + kStagingBytes + (2 * kSwMaxStages + 9 + 2 * kSwWStages) * 8 + 8 + (size_t)bias_floats * 4 + 16;
}

void pick_tile256(int H, int W, bool f32, int* tw, int* th) {
  const int cand[5][2] = {{16, 16}, {32, 8}, {8, 32}, {64, 4}, {128, 2}};
  double best = -1;
  for (int i = 0; i < 5; ++i) {
    const int w = cand[i][0], h = cand[i][1];
    if (f32 && w > 32) continue;   // fp32 granules are 32 pixels = whole image rows of the tile
    const double util = (double)H * W / ((double)ceil_div(W, w) * w * (double)ceil_div(H, h) * h);
    if (util > best + 1e-9) { best = util; *tw = w; *th = h; }
  }
}

}  // namespace

int conv_sw_init(gt_engine* e) {
  GT_CUDA(e, cudaFuncSetAttribute(conv_sw_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  GT_CUDA(e, cudaFuncSetAttribute(conv_sw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return GT_OK;
}

int conv_sw_plan(gt_engine* e, ConvOp* op, const ConvPlanArgs& a) {
  const View& in = a.in;
  const int cin = a.cin, k = a.k, stride = a.stride;
  const int kbe = (a.kb_elems == 64 && cin == 32) ? 32 : a.kb_elems;
  GT_CHECK(e, in.C == cin, "conv plan: input view has %d channels, conv expects %d", in.C, cin);
  GT_CHECK(e, (in.ctot % 8) == 0 && (in.coff % 8) == 0, "conv plan: input slice must be 16-byte aligned");
  GT_CHECK(e, k >= 1 && k <= 3 && (stride == 1 || stride == 2), "conv plan: k=%d stride=%d unsupported", k, stride);
  GT_CHECK(e, a.pre == nullptr, "conv plan (swapped): the half-resolution pre-activation add is a pixel-major epilogue feature");
  ConvParams& p = op->p;
  memset(&p, 0, sizeof(p));
  op->swapped = 1;
  op->pair = 0;
  const int pad = a.pad >= 0 ? a.pad : k / 2;
  const int Ho = a.Ho > 0 ? a.Ho : (in.H + 2 * pad - k) / stride + 1, Wo = a.Wo > 0 ? a.Wo : (in.W + 2 * pad - k) / stride + 1;
  const int cout = a.cout;
  op->cin = cin; op->cout = cout; op->k = k; op->stride = stride;
  p.B = a.Bmax; p.H = Ho; p.W = Wo;
  pick_tile256(Ho, Wo, a.out_f32 != nullptr, &p.tw, &p.th);
  if (a.out_s2d) { p.tw = 16; p.th = 16; GT_CHECK(e, (Ho % 8) == 0, "conv plan: s2d output needs Ho %% 8 == 0"); }   // granules are 16 x 8: never straddle images
  // variant 2: halo staging for stride-1 k >= 2 layers with 16-bit NHWC outputs (8-pixel-wide tiles: one 8-row descriptor group per image row)
  bool halo = e->plan_variant == 2 && stride == 1 && k >= 2 && !a.out_f32 && !a.out_s2d;
  if (halo) { p.tw = 8; p.th = 32; }
  // variant 6: CTA pairs (one 256-channel x 256-pixel tile per pair, per-tap path); every tile shape has an even number of image rows
  const bool pair = e->plan_variant == 6 && (cout % (2 * kCo)) == 0 && !a.out_f32 && !a.out_s2d;
  GT_CHECK(e, p.tw * stride <= 256 && p.th * stride <= 256, "conv plan: TMA box too large");
  p.tiles_x = ceil_div(Wo, p.tw); p.tiles_y = ceil_div(Ho, p.th);
  p.stride = stride; p.ksize = k; p.pad = pad;
  p.kb_elems = kbe;
  p.kc_blocks = ceil_div(cin, kbe);
  op->cin_pad = p.kc_blocks * kbe;
  p.num_kb = k * k * p.kc_blocks;
  p.n_tiles = pair ? cout / (2 * kCo) : ceil_div(cout, kCo);
  op->cout_pad = p.n_tiles * (pair ? 2 * kCo : kCo);
  op->pair = pair ? 1 : 0;
  p.BN = kCo;                       // TileIter::coord's n0 = nt * BN = first output channel of the tile
  p.tmem_cols = 512; p.acc_stride = kPx;
  const int w_bytes = kCo * kbe * 2, x_bytes = (pair ? kPx / 2 : kPx) * kbe * 2;   // (pair: per CTA)
  const int wres_bytes = p.num_kb * w_bytes;
  const size_t budget = (size_t)e->conv_smem_kb * 1024;
  const int bias_floats = pair ? p.n_tiles * kCo : op->cout_pad;                    // (pair: a CTA keeps its own half of every tile)
  p.b_resident = (!pair && p.n_tiles == 1 && wres_bytes <= 80 * 1024) ? 1 : 0;
  int stage_bytes = 0, stages = 0;
  if (halo) {
    const int halo_rows = (p.tw + k - 1) * (p.th + k - 1);
    p.halo_tx = halo_rows * kbe * 2;
    p.halo_bytes = (p.halo_tx + 1023) / 1024 * 1024;
    p.a_stages = kSwWStages;
    const int wreg = p.b_resident ? wres_bytes : kSwWStages * w_bytes;
    const size_t fixed = sw_smem_bytes(0, 0, wreg, op->cout_pad);
    stages = fixed < budget ? (int)((budget - fixed) / p.halo_bytes) : 0;
    if (stages >= 2) {
      p.halo = 1;
      stage_bytes = p.halo_bytes;
      if (stages > kSwMaxStages) stages = kSwMaxStages;
    } else {   // does not fit: plain swapped plan
      halo = false;
      pick_tile256(Ho, Wo, false, &p.tw, &p.th);
      p.tiles_x = ceil_div(Wo, p.tw); p.tiles_y = ceil_div(Ho, p.th);
      p.halo_tx = p.halo_bytes = p.a_stages = 0;
    }
  }
  for (int attempt = 0; attempt < 2 && !halo; ++attempt) {   // resident weights only if at least 3 ring stages remain
    stage_bytes = p.b_resident ? x_bytes : x_bytes + w_bytes;
    const size_t fixed = sw_smem_bytes(0, stage_bytes, p.b_resident ? wres_bytes : 0, bias_floats);
    stages = fixed < budget ? (int)((budget - fixed) / stage_bytes) : 0;
    if (stages >= 3 || !p.b_resident) break;
    p.b_resident = 0;
  }
  if (stages > kSwMaxStages) stages = kSwMaxStages;
  GT_CHECK(e, stages >= 2, "conv plan (swapped): operands do not fit shared memory");
  p.stages = stages;
  p.cout = cout; p.act = getenv("GT_DEBUG_NOACT") ? 0 : a.act; p.fp16 = e->cfg.act_dtype == GT_ACT_FP16 ? 1 : 0;
  p.scale = a.scale;
  if (a.out_f32) {
    p.out_f32 = 1; p.out = a.out_f32; p.out_img_stride = a.out_img_stride; p.out_ctot = a.out_ctot_f32; p.out_coff = a.out_coff_f32;
  } else {
    const View* out = a.out;
    if (a.out_s2d) GT_CHECK(e, out && out->H == 2 * Ho && out->W == 2 * Wo && out->C * 4 == cout && cout == 128, "conv plan: s2d output view mismatch");
    else GT_CHECK(e, out && out->H == Ho && out->W == Wo && out->C == cout, "conv plan: output view mismatch");
    GT_CHECK(e, (out->ctot % 8) == 0 && (out->coff % 8) == 0 && (cout % 8) == 0, "conv plan: output slice must be 16-byte aligned");
    p.out_f32 = 0; p.out = out->ptr; p.out_img_stride = (long long)Ho * Wo; p.out_ctot = out->ctot; p.out_coff = out->coff;
    p.out_s2d = a.out_s2d ? 1 : 0; p.out_rows = Ho;
  }
  if (a.res) {
    GT_CHECK(e, a.res->H == Ho && a.res->W == Wo && a.res->C == cout && !a.out_f32, "conv plan: residual view mismatch");
    p.res = a.res->ptr; p.res_ctot = a.res->ctot; p.res_coff = a.res->coff;
  }
  if (a.up) {
    GT_CHECK(e, a.up->H == 2 * Ho && a.up->W == 2 * Wo && a.up->C == cout && !a.out_f32, "conv plan: upsample view mismatch");
    p.up = a.up->ptr; p.up_ctot = a.up->ctot; p.up_coff = a.up->coff;
  }
  op->smem = sw_smem_bytes(p.stages, stage_bytes, p.b_resident ? wres_bytes : (p.halo ? kSwWStages * w_bytes : 0), bias_floats);
  op->flops = 2.0 * Ho * Wo * (double)cout * cin * k * k;
  op->bytes = (double)in.H * in.W * cin * 2 + (double)Ho * Wo * cout * (a.out_f32 ? 4 : 2) * (a.up ? 5 : 1) + (a.res ? (double)Ho * Wo * cout * 2 : 0.0);

  const size_t wn = (size_t)op->cout_pad * k * k * op->cin_pad;
  GT_TRY(e->dev_alloc((void**)&op->w_dev, wn * sizeof(bf16)));
  GT_TRY(e->dev_alloc((void**)&op->b_dev, (size_t)op->cout_pad * sizeof(float)));
  GT_CUDA(e, cudaMemset(op->w_dev, 0, wn * sizeof(bf16)));
  GT_CUDA(e, cudaMemset(op->b_dev, 0, (size_t)op->cout_pad * sizeof(float)));
  p.bias = op->b_dev;

  const CUtensorMapDataType dt = p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = kbe == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kbe == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  {  // X: NHWC input slice {C, W, H, N}, box = one 256-pixel tile of one k-block (op->tmB keeps the "activation" map)
    cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)a.Bmax};
    cuuint64_t gstr[3] = {(cuuint64_t)in.ctot * 2, (cuuint64_t)in.W * in.ctot * 2, (cuuint64_t)in.H * in.W * in.ctot * 2};
    cuuint32_t box[4] = {(cuuint32_t)kbe, (cuuint32_t)(p.halo ? p.tw + k - 1 : p.tw * stride),
                         (cuuint32_t)(p.halo ? p.th + k - 1 : (pair ? p.th / 2 : p.th) * stride), 1};   // (pair: each CTA loads half of the tile's image rows)
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = conv_tc_encode()(&op->tmA, dt, 4, (void*)(in.ptr + in.coff), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(X) failed: %d", (int)r);
  }
  {  // W: packed weights {Ktot, cout_pad}, box = 128 rows of one k-block
    const cuuint64_t ktot = (cuuint64_t)k * k * op->cin_pad;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)op->cout_pad};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kbe, (cuuint32_t)kCo};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = conv_tc_encode()(&op->tmB, dt, 2, (void*)op->w_dev, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }
  {  // output granules
    auto enc = [&](CUtensorMap* tm, CUtensorMapDataType odt, void* base, int box_c, int box_w, int box_h, cuuint64_t W_, cuuint64_t H_,
                   cuuint64_t pix_b, cuuint64_t row_b, cuuint64_t img_b) -> CUresult {
      cuuint64_t gdim[4] = {(cuuint64_t)cout, W_, H_, (cuuint64_t)a.Bmax};
      cuuint64_t gstr[3] = {pix_b, row_b, img_b};
      cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      return conv_tc_encode()(tm, odt, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r;
    memset(&op->tmUp, 0, sizeof(op->tmUp));
    memset(&op->tmRes, 0, sizeof(op->tmRes));
    if (a.out_f32) {
      const cuuint64_t ps = (cuuint64_t)a.out_ctot_f32 * 4;
      GT_CHECK(e, (ps % 16) == 0 && (a.out_coff_f32 % 4) == 0, "conv plan: fp32 output rows must be 16-byte aligned");
      r = enc(&op->tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (void*)(a.out_f32 + a.out_coff_f32), 32, p.tw, 32 / p.tw, Wo, Ho, ps, (cuuint64_t)Wo * ps,
              (cuuint64_t)a.out_img_stride * ps);
    } else {
      const cuuint64_t ps = (cuuint64_t)a.out->ctot * 2;
      if (a.out_s2d) r = encode_s2d_out(&op->tmOut, dt, a.out, Wo, Ho, a.Bmax, p.tw, p.th / 2);
      else r = enc(&op->tmOut, dt, (void*)(a.out->ptr + a.out->coff), 64, p.tw, p.th / 2, Wo, Ho, ps, (cuuint64_t)Wo * ps, (cuuint64_t)Ho * Wo * ps);
      if (r == CUDA_SUCCESS && a.res) {
        const cuuint64_t rs = (cuuint64_t)a.res->ctot * 2;
        r = enc(&op->tmRes, dt, (void*)(a.res->ptr + a.res->coff), 64, p.tw, p.th / 2, Wo, Ho, rs, (cuuint64_t)Wo * rs, (cuuint64_t)Ho * Wo * rs);
      }
      if (r == CUDA_SUCCESS && a.up) {
        const cuuint64_t us = (cuuint64_t)a.up->ctot * 2, W2 = (cuuint64_t)2 * Wo, H2 = (cuuint64_t)2 * Ho;
        for (int d = 0; d < 4 && r == CUDA_SUCCESS; ++d) {
          bf16* base = a.up->ptr + a.up->coff + ((size_t)(d >> 1) * W2 + (d & 1)) * a.up->ctot;
          r = enc(&op->tmUp.m[d], dt, (void*)base, 64, p.tw, p.th / 2, Wo, Ho, 2 * us, 2 * W2 * us, H2 * W2 * us);
        }
      }
    }
    GT_CHECK(e, r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(out, swapped) failed: %d (cout=%d)", (int)r, cout);
  }
  return GT_OK;
}

int conv_sw_launch_range(gt_engine* e, const ConvOp* op, int b0, int nb, cudaStream_t st) {
  ConvParams p = op->p;
  p.B = nb;
  p.img0 = b0;
  p.total_tiles = p.tiles_x * p.tiles_y * nb * p.n_tiles;
  const int units = op->pair ? conv_tc_num_sms() / 2 : conv_tc_num_sms();   // persistent CTAs, or CTA pairs (one per TPC)
  const int grid = (p.total_tiles < units ? p.total_tiles : units) * (op->pair ? 2 : 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kSwThreads); cfg.dynamicSmemBytes = op->smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (e->pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (op->pair) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  if (op->pair) GT_CUDA(e, cudaLaunchKernelEx(&cfg, conv_sw_kernel<1>, op->tmB, op->tmA, op->tmOut, op->tmUp, op->tmRes, p));
  else GT_CUDA(e, cudaLaunchKernelEx(&cfg, conv_sw_kernel<0>, op->tmB, op->tmA, op->tmOut, op->tmUp, op->tmRes, p));   // (weights, activations, ...)
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}
