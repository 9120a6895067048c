// Hamming 2-NN on the tensor cores (tcgen05, sm_100a).
//
// Restates cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) as stabilo calls it (/root/reference/geotrax/extract.py:181;
// SURVEY.md 8a-11) -- bit-exact, ties to the lower train index -- as a GEMM:
//
//     hamming(q, t) = popc(q) + popc(t) - 2 * <q, t>          (q, t as 256-vectors of 0 / 1)
//
// The POPC form (match_kernel, match_ransac.cu) is bound by the POPC pipe (16 / clk / SM): 2,000 x 4,000 x 16 frames x 8 words =
// 228 us per step.  The same inner products are 65.5 GFLOP of 8-bit MMA: descriptors expanded to E4M3 bytes (0x38 = 1.0, 0x00),
// `tcgen05.mma.kind::f8f6f4` with f32 accumulation -- every product and every partial sum is an integer <= 256, so the result is exact.
//
//   desc_expand_kernel   bits -> the shared-memory IMAGE of a K-major SWIZZLE_128B operand tile, group of 128 descriptors by group
//                        ([group][k-block][128 rows][128 B], 16-byte chunks XOR-swizzled by row & 7), so the matcher stages operands
//                        with plain bulk copies (cp.async.bulk, no tensor map); plus c[j] = popc(t_j) * 8192 + j as a float
//                        (FLT_MAX-like for the rows of the last group beyond the count).
//   match_tc_kernel      one CTA = QT query groups (M = 128 each) against all train groups (N = 128 per tile): warp 0 bulk-copy
//                        producer, warp 1 MMA issuer, warps 2-9 epilogue (thread = query row): key = c[j] - 16384 * dot is
//                        (popc(t) - 2 dot) * 8192 + j, exact in f32 (|key| < 2^22); running two smallest keys per row; adding
//                        popc(q) * 8192 at the end gives dist * 8192 + index.  Two TMEM accumulator sets: the epilogue of tile i
//                        overlaps the MMAs of tile i + 1.
//
// GT_MATCH = 2 (default) E4M3 operands | 1 fp16 operands (kind::f16; diagnostic: same pipeline, the conv kernels' MMA kind) | 0 POPC kernel.
//
// The same pipeline with fp16 operands of 128 elements is the brute-force L2 matcher of the registration path (SURVEY.md 8f-3:
// /root/reference/geotrax/utils/registration.py:57-93 runs stabilo with RootSIFT descriptors, cv2.BFMatcher(NORM_L2).knnMatch(k=2)):
// key = |t|^2 - 2 <q, t> ranks the train rows of a query exactly like the L2 distance; the tensor-core pass keeps the FOUR best
// candidates per query on fp16-rounded operands, l2_rerank_kernel recomputes their distances in fp32 on the original descriptors and
// returns the two nearest (ties to the lower index) -- the fp16 pass only has to get the true two nearest into its top four.
#include "engine.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kGroups = GT_MAX_KP / 128;   // descriptor groups per slot
constexpr int kChunkBytes = 16384;         // one k-block of one group: 128 rows x 128 B
constexpr float kBig = 3.0e38f;
constexpr int kMtThreads = 320;

enum { MT_E4M3 = 0, MT_F16 = 1, MT_L2 = 2 };          // Hamming on E4M3 bytes | Hamming on fp16 | L2 on fp16 (128-element float descriptors)
template <int MODE>
struct MT {
  static constexpr bool FP8 = MODE == MT_E4M3;
  static constexpr int KB = MODE == MT_F16 ? 4 : 2; // 128-byte k-blocks per descriptor (256 E4M3 / 256 fp16 / 128 fp16 elements)
  static constexpr int QT = MODE == MT_F16 ? 1 : 2; // query groups per CTA
  static constexpr int STAGES = MODE == MT_F16 ? 2 : 3;   // train-group ring
  static constexpr int GROUP_BYTES = KB * kChunkBytes;
  static constexpr int A_BYTES = QT * GROUP_BYTES;
  static constexpr int TAIL = 4096;                 // barriers, TMEM pointer, c[] slots
  static constexpr int SMEM = A_BYTES + STAGES * GROUP_BYTES + TAIL + 1024;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <bool FP8, int ACC>
__device__ __forceinline__ void umma_kind(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc = 1u) {
  if constexpr (FP8) {
    if constexpr (ACC == 2)
      asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                   "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
                   : "memory");
    else
      asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                   "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
                   : "memory");
  } else {
    umma_lohi<ACC>(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
  }
}

// ---- bits -> operand image ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread_e4m3(uint32_t nib) {   // 4 bits -> 4 bytes of 0x38 (1.0) / 0x00
  return ((nib * 0x00204081u) & 0x01010101u) * 0x38u;
}
__device__ __forceinline__ uint32_t spread_f16(uint32_t two) {    // 2 bits -> 2 halves of 0x3C00 (1.0) / 0
  return (two & 1u) * 0x3C00u + (two >> 1) * 0x3C000000u;
}

template <int MODE>
__global__ void __launch_bounds__(256) desc_expand_kernel(const uint8_t* __restrict__ desc_all, const int* __restrict__ kp_count, int slot0, int B,
                                                          int ref_slot, uint8_t* __restrict__ desc_x, float* __restrict__ desc_c) {
  using C = MT<MODE>;
  constexpr bool FP8 = C::FP8;
  const int slot = (int)blockIdx.y < B ? slot0 + (int)blockIdx.y : ref_slot;
  const int n = min(kp_count[slot], GT_MAX_KP);
  const int g = blockIdx.x;
  if (g * 128 >= n) return;
  const uint8_t* D = desc_all + (size_t)slot * GT_MAX_KP * 32;
  uint8_t* X = desc_x + ((size_t)slot * kGroups + g) * C::GROUP_BYTES;
  constexpr int CPR = FP8 ? 16 : 32;   // 16-byte chunks per expanded row
  for (int i = threadIdx.x; i < 128 * CPR; i += 256) {
    const int r = i / CPR, c = i % CPR, row = g * 128 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < n) {
      if constexpr (FP8) {
        const uint32_t bits = *reinterpret_cast<const unsigned short*>(D + (size_t)row * 32 + c * 2);
        v.x = spread_e4m3(bits & 15u); v.y = spread_e4m3((bits >> 4) & 15u); v.z = spread_e4m3((bits >> 8) & 15u); v.w = spread_e4m3(bits >> 12);
      } else {
        const uint32_t bits = D[(size_t)row * 32 + c];
        v.x = spread_f16(bits & 3u); v.y = spread_f16((bits >> 2) & 3u); v.z = spread_f16((bits >> 4) & 3u); v.w = spread_f16(bits >> 6);
      }
    }
    const int kb = c >> 3, cc = c & 7;
    *reinterpret_cast<uint4*>(X + (size_t)kb * kChunkBytes + r * 128 + ((cc ^ (r & 7)) << 4)) = v;
  }
  if (threadIdx.x < 128) {
    const int row = g * 128 + threadIdx.x;
    float cj = kBig;
    if (row < n) {
      const uint4 a = *reinterpret_cast<const uint4*>(D + (size_t)row * 32), b = *reinterpret_cast<const uint4*>(D + (size_t)row * 32 + 16);
      const int pc = __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
      cj = (float)(pc * 8192 + row);
    }
    desc_c[(size_t)slot * GT_MAX_KP + row] = cj;
  }
}

// float descriptors [n][128] -> fp16 operand image (two 64-element k-blocks per row) + c[row] = |d|^2 (fp32, of the original values);
// one block per group of 128 rows, rows beyond n are zero operands with a huge constant
__global__ void __launch_bounds__(256) desc_expand_f32_kernel(const float* __restrict__ D, int n, uint8_t* __restrict__ X, float* __restrict__ cn) {
  constexpr int GB = MT<MT_L2>::GROUP_BYTES;
  const int g = blockIdx.x;
  uint8_t* Xg = X + (size_t)g * GB;
  for (int i = threadIdx.x; i < 128 * 16; i += 256) {   // 16 chunks of 8 halves per row
    const int r = i >> 4, c = i & 15, row = g * 128 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(D + (size_t)row * 128 + c * 8)), b = __ldg(reinterpret_cast<const float4*>(D + (size_t)row * 128 + c * 8) + 1);
      __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
      v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1); v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
    }
    const int kb = c >> 3, cc = c & 7;
    *reinterpret_cast<uint4*>(Xg + (size_t)kb * kChunkBytes + r * 128 + ((cc ^ (r & 7)) << 4)) = v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < 128; r += 8) {
    const int row = g * 128 + r;
    float s = 0.f;
    if (row < n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(D + (size_t)row * 128) + lane);
      s = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) cn[row] = row < n ? s : kBig;
  }
}

// ---- the matcher --------------------------------------------------------------------------------------------------------------
struct MatchArgs {
  const uint8_t* xq; const uint8_t* xt;        // operand images of the query / train rows
  long long xq_bstride, xt_bstride;            // bytes per batch index (0: shared by every batch index)
  const float* cq; const float* ct;            // per-row constants
  long long cq_bstride, ct_bstride;            // floats per batch index
  const int* nq; const int* nt;                // row counts (device)
  int nq_step, nt_step, cap;                   // count stride per batch index; row capacity
  int* out_a; int* out_b;                      // Hamming: index / distance [b][cap][2]; L2: candidate indices [row][4], out_b unused
  long long out_bstride;
};

__device__ __forceinline__ void top2(float& m1, float& m2, float k) {
  m2 = fminf(m2, fmaxf(m1, k));
  m1 = fminf(m1, k);
}
struct Top4 {            // four smallest (key, index), ascending; strict compares keep the earlier (lower) index on equal keys
  float v0, v1, v2, v3; int i0, i1, i2, i3;
  __device__ __forceinline__ void init() { v0 = v1 = v2 = v3 = kBig; i0 = i1 = i2 = i3 = -1; }
  __device__ __forceinline__ void insert(float k, int idx) {
    if (!(k < v3)) return;
    if (k < v2) {
      v3 = v2; i3 = i2;
      if (k < v1) {
        v2 = v1; i2 = i1;
        if (k < v0) { v1 = v0; i1 = i0; v0 = k; i0 = idx; } else { v1 = k; i1 = idx; }
      } else { v2 = k; i2 = idx; }
    } else { v3 = k; i3 = idx; }
  }
};

template <int MODE>
__global__ void __launch_bounds__(kMtThreads, 1) match_tc_kernel(const MatchArgs a) {
  using C = MT<MODE>;
  constexpr bool FP8 = C::FP8;
  constexpr int QT = C::QT, KB = C::KB, STAGES = C::STAGES;
  const int b = blockIdx.y;
  const int NQ = min(a.nq[b * a.nq_step], a.cap), NT = min(a.nt[b * a.nt_step], a.cap);
  const int q0 = blockIdx.x * 128 * QT;
  if (q0 >= NQ) return;
  int* oi = a.out_a + (size_t)b * a.out_bstride;
  int* od = a.out_b + (size_t)b * a.out_bstride;
  if (NT == 0) {
    if constexpr (MODE == MT_L2) {
      for (int i = threadIdx.x; i < 512 * QT; i += kMtThreads)
        if (q0 + (i >> 2) < NQ) oi[(size_t)q0 * 4 + i] = -1;
    } else {
      for (int i = threadIdx.x; i < 256 * QT; i += kMtThreads)
        if (q0 + (i >> 1) < NQ) { oi[(size_t)q0 * 2 + i] = -1; od[(size_t)q0 * 2 + i] = -1; }
    }
    return;
  }
  const int ntiles = (NT + 127) >> 7;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                  // [QT][KB][128 rows][128 B]
  uint8_t* sB = smem + C::A_BYTES;                     // [STAGES][KB][128 rows][128 B]
  uint8_t* tail = sB + (size_t)STAGES * C::GROUP_BYTES;
  uint64_t* a_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full_bar = a_bar + 1;
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_c = reinterpret_cast<float*>(tail + 256);   // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(smem_u32(a_bar), 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tfull_bar[s]), 1); mbar_init(smem_u32(&tempty_bar[s]), 4 * QT); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  const uint8_t* Xq = a.xq + (size_t)b * a.xq_bstride + (size_t)(q0 >> 7) * C::GROUP_BYTES;
  const uint8_t* Xt = a.xt + (size_t)b * a.xt_bstride;

  if (warp == 0) {
    // ===== producer: the query groups once, then the train groups through the ring =====
    if (elect_one()) {
      mbar_expect_tx(smem_u32(a_bar), (uint32_t)C::A_BYTES);
      bulk_g2s(smem_u32(sA), Xq, (uint32_t)C::A_BYTES, smem_u32(a_bar));
    }
    __syncwarp();
    uint32_t s = 0, ph = 0;
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, (uint32_t)C::GROUP_BYTES);
        bulk_g2s(smem_u32(sB + (size_t)s * C::GROUP_BYTES), Xt + (size_t)t * C::GROUP_BYTES, (uint32_t)C::GROUP_BYTES, fb);
      }
      __syncwarp();
      if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D[128 queries][128 train] (+)= A . B^T, KB k-blocks x 4 MMAs of 32 bytes of K =====
    const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // D f32; A, B format 0 (E4M3 / F16), K-major
    const uint32_t hi = desc_hi(1024u, 2u);
    const uint32_t a_lo0 = (smem_u32(sA) & 0x3FFFFu) >> 4, b_lo0 = (smem_u32(sB) & 0x3FFFFu) >> 4;
    mbar_wait(smem_u32(a_bar), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t s = 0, ph = 0;
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t as = (uint32_t)t & 1u;
      mbar_wait(smem_u32(&tempty_bar[as]), (((uint32_t)t >> 1) & 1u) ^ 1u);
      mbar_wait(smem_u32(&full_bar[s]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t b_lo = b_lo0 + s * (uint32_t)(C::GROUP_BYTES >> 4);
#pragma unroll
        for (int qt = 0; qt < QT; ++qt) {
          const uint32_t tacc = tmem_base + (as * (uint32_t)QT + (uint32_t)qt) * 128u;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t al = a_lo0 + (uint32_t)((qt * KB + kb) * (kChunkBytes >> 4)), bl = b_lo + (uint32_t)(kb * (kChunkBytes >> 4));
            if (kb == 0) umma_kind<FP8, 0>(tacc, al, hi, bl, hi, idesc);
            else umma_kind<FP8, 1>(tacc, al, hi, bl, hi, idesc);
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_kind<FP8, 1>(tacc, al + 2u * k, hi, bl + 2u * k, hi, idesc);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
        umma_commit(smem_u32(&tfull_bar[as]));
      }
      __syncwarp();
      if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1u; }
    }
  } else if ((warp - 2) >> 2 < QT) {
    // ===== epilogue: thread = query row, running smallest keys over all train columns =====
    const int q = warp & 3, qt = (warp - 2) >> 2;
    const int e = (warp - 2) * 32 + lane;                 // index inside the epilogue group (QT x 128 threads)
    const float* ct = a.ct + (size_t)b * a.ct_bstride;
    float m1a = kBig, m2a = kBig, m1b = kBig, m2b = kBig;
    Top4 t4; t4.init();
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t as = (uint32_t)t & 1u;
      if (e < 128) s_c[as * 128 + e] = __ldg(ct + t * 128 + e);
      asm volatile("bar.sync 1, %0;" ::"n"(QT * 128) : "memory");
      if (lane == 0) mbar_wait(smem_u32(&tfull_bar[as]), ((uint32_t)t >> 1) & 1u);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (as * (uint32_t)QT + (uint32_t)qt) * 128u;
      const uint32_t cbase = smem_u32(s_c + as * 128);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld_x32(trow + (uint32_t)(ch * 32), v);
        tmem_ld_wait();
        if constexpr (MODE == MT_L2) {
          // key = |t|^2 - 2 <q, t>; almost every group of four fails the cut against the current fourth-best
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 c4 = lds_f4(cbase + (uint32_t)((ch * 32 + j * 4) * 4));
            const float k0 = fmaf(__uint_as_float(v[4 * j + 0]), -2.0f, c4.x), k1 = fmaf(__uint_as_float(v[4 * j + 1]), -2.0f, c4.y);
            const float k2 = fmaf(__uint_as_float(v[4 * j + 2]), -2.0f, c4.z), k3 = fmaf(__uint_as_float(v[4 * j + 3]), -2.0f, c4.w);
            if (fminf(fminf(k0, k1), fminf(k2, k3)) < t4.v3) {
              const int col = t * 128 + ch * 32 + j * 4;
              t4.insert(k0, col); t4.insert(k1, col + 1); t4.insert(k2, col + 2); t4.insert(k3, col + 3);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 c4 = lds_f4(cbase + (uint32_t)((ch * 32 + j * 4) * 4));
            top2(m1a, m2a, fmaf(__uint_as_float(v[4 * j + 0]), -16384.0f, c4.x));
            top2(m1b, m2b, fmaf(__uint_as_float(v[4 * j + 1]), -16384.0f, c4.y));
            top2(m1a, m2a, fmaf(__uint_as_float(v[4 * j + 2]), -16384.0f, c4.z));
            top2(m1b, m2b, fmaf(__uint_as_float(v[4 * j + 3]), -16384.0f, c4.w));
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
    }
    const int row = q0 + qt * 128 + q * 32 + lane;
    if constexpr (MODE == MT_L2) {
      if (row < NQ) *reinterpret_cast<int4*>(oi + (size_t)row * 4) = make_int4(t4.i0, t4.i1, t4.i2, t4.i3);   // keys at or above kBig / 2 never enter (columns beyond NT)
    } else {
      // merge the two interleaved accumulators (keys are distinct: the index is part of the key)
      const float m1 = fminf(m1a, m1b);
      const float m2 = fminf(fmaxf(m1a, m1b), fminf(m2a, m2b));
      if (row < NQ) {
        const float pa = __ldg(a.cq + (size_t)b * a.cq_bstride + row) - (float)row;   // popc(q) * 8192
        const int k1 = m1 < 1e30f ? (int)(m1 + pa) : -1, k2 = m2 < 1e30f ? (int)(m2 + pa) : -1;
        oi[(size_t)row * 2 + 0] = k1 < 0 ? -1 : (k1 & 8191); od[(size_t)row * 2 + 0] = k1 < 0 ? -1 : (k1 >> 13);
        oi[(size_t)row * 2 + 1] = k2 < 0 ? -1 : (k2 & 8191); od[(size_t)row * 2 + 1] = k2 < 0 ? -1 : (k2 >> 13);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// exact fp32 re-rank of the four tensor-core candidates of a query (one warp per query): squared L2 distance on the original
// descriptors, the two nearest by (distance, index); distances are returned as sqrt like cv2.BFMatcher(NORM_L2)
__global__ void __launch_bounds__(256) l2_rerank_kernel(const float* __restrict__ Q, const float* __restrict__ T, int nq, const int* __restrict__ cand,
                                                        int* __restrict__ out_idx, float* __restrict__ out_dist) {
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= nq) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(Q + (size_t)q * 128) + lane);
  float d[4]; int id[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    id[c] = cand[(size_t)q * 4 + c];
    float s = kBig;
    if (id[c] >= 0) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(T + (size_t)id[c] * 128) + lane);
      const float dx = a.x - t.x, dy = a.y - t.y, dz = a.z - t.z, dw = a.w - t.w;
      s = dx * dx + dy * dy + dz * dz + dw * dw;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    d[c] = s;
  }
  if (lane == 0) {
    int b1 = -1, b2 = -1; float d1 = kBig, d2 = kBig;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (id[c] < 0) continue;
      if (d[c] < d1 || (d[c] == d1 && id[c] < b1)) { d2 = d1; b2 = b1; d1 = d[c]; b1 = id[c]; }
      else if (d[c] < d2 || (d[c] == d2 && id[c] < b2)) { d2 = d[c]; b2 = id[c]; }
    }
    out_idx[(size_t)q * 2] = b1; out_idx[(size_t)q * 2 + 1] = b2;
    out_dist[(size_t)q * 2] = b1 >= 0 ? sqrtf(d1) : -1.f; out_dist[(size_t)q * 2 + 1] = b2 >= 0 ? sqrtf(d2) : -1.f;
  }
}

template <int MODE>
int launch(gt_engine* e, int q_slot0, int q_step, int nq_cap, int t_slot0, int t_step, int batch, cudaStream_t st) {
  using C = MT<MODE>;
  const int R = e->cfg.max_batch;
  const long long slot_bytes = (long long)kGroups * C::GROUP_BYTES;
  // the slots involved: [0, batch) and the reference slot (gt_match uses slot 0 and the reference slot as well)
  desc_expand_kernel<MODE><<<dim3(kGroups, (unsigned)batch + 1), 256, 0, st>>>(e->desc_all, e->kp_count, 0, batch, R, e->desc_x, e->desc_c);
  MatchArgs a;
  a.xq = e->desc_x + (size_t)q_slot0 * slot_bytes; a.xq_bstride = q_step * slot_bytes;
  a.xt = e->desc_x + (size_t)t_slot0 * slot_bytes; a.xt_bstride = t_step * slot_bytes;
  a.cq = e->desc_c + (size_t)q_slot0 * GT_MAX_KP; a.cq_bstride = (long long)q_step * GT_MAX_KP;
  a.ct = e->desc_c + (size_t)t_slot0 * GT_MAX_KP; a.ct_bstride = (long long)t_step * GT_MAX_KP;
  a.nq = e->kp_count + q_slot0; a.nq_step = q_step; a.nt = e->kp_count + t_slot0; a.nt_step = t_step; a.cap = GT_MAX_KP;
  a.out_a = e->match_idx; a.out_b = e->match_dist; a.out_bstride = (long long)GT_MAX_KP * 2;
  match_tc_kernel<MODE><<<dim3((unsigned)ceil_div(nq_cap, 128 * C::QT), (unsigned)batch), kMtThreads, C::SMEM, st>>>(a);
  e->launches += 2;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

}  // namespace

int match_tc_build(gt_engine* e) {
  GT_CUDA(e, cudaFuncSetAttribute(match_tc_kernel<MT_E4M3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT<MT_E4M3>::SMEM));
  GT_CUDA(e, cudaFuncSetAttribute(match_tc_kernel<MT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT<MT_F16>::SMEM));
  GT_CUDA(e, cudaFuncSetAttribute(match_tc_kernel<MT_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MT<MT_L2>::SMEM));
  if (e->match_mode == 0) return GT_OK;
  const size_t S = (size_t)e->cfg.max_batch + 1;
  const size_t gb = e->match_mode == 2 ? MT<MT_E4M3>::GROUP_BYTES : MT<MT_F16>::GROUP_BYTES;
  GT_TRY(e->dev_alloc((void**)&e->desc_x, S * kGroups * gb));
  GT_TRY(e->dev_alloc((void**)&e->desc_c, S * GT_MAX_KP * sizeof(float)));
  GT_CUDA(e, cudaMemset(e->desc_x, 0, S * kGroups * gb));   // rows never written stay valid operands (0.0)
  return GT_OK;
}

// Queries = slots q_slot0 + b * q_step, train = slots t_slot0 + b * t_step of desc_all, b < batch (all within [0, batch) or the reference
// slot); results in match_idx / match_dist [b][GT_MAX_KP][2].  nq_cap bounds the query count (grid size; rows beyond the count exit).
int match_tc_run(gt_engine* e, int q_slot0, int q_step, int nq_cap, int t_slot0, int t_step, int batch, cudaStream_t st) {
  if (e->match_mode == 2) return launch<MT_E4M3>(e, q_slot0, q_step, nq_cap, t_slot0, t_step, batch, st);
  return launch<MT_F16>(e, q_slot0, q_step, nq_cap, t_slot0, t_step, batch, st);
}

// Brute-force L2 2-NN of float descriptors [n][128] (device pointers): tensor-core candidate pass + exact fp32 re-rank.  The work
// buffers live in the engine and grow on demand (registration matches up to 250,000 x 250,000 descriptors once or twice per video).
int match_l2_run(gt_engine* e, const float* q_dev, int nq, const float* t_dev, int nt, int* out_idx_dev, float* out_dist_dev, cudaStream_t st) {
  using C = MT<MT_L2>;
  const int need = std::max(ceil_div(std::max(nq, nt), 256) * 256, 256);
  if (need > e->reg_cap) {
    for (void* p : {(void*)e->reg_xq, (void*)e->reg_xt, (void*)e->reg_cq, (void*)e->reg_ct, (void*)e->reg_cand, (void*)e->reg_n})
      if (p) cudaFree(p);
    e->reg_xq = e->reg_xt = nullptr; e->reg_cq = e->reg_ct = nullptr; e->reg_cand = nullptr; e->reg_n = nullptr; e->reg_cap = 0;
    const size_t xb = (size_t)need / 128 * C::GROUP_BYTES;
    GT_CUDA(e, cudaMalloc((void**)&e->reg_xq, xb)); GT_CUDA(e, cudaMalloc((void**)&e->reg_xt, xb));
    GT_CUDA(e, cudaMalloc((void**)&e->reg_cq, (size_t)need * sizeof(float))); GT_CUDA(e, cudaMalloc((void**)&e->reg_ct, (size_t)need * sizeof(float)));
    GT_CUDA(e, cudaMalloc((void**)&e->reg_cand, (size_t)need * 4 * sizeof(int))); GT_CUDA(e, cudaMalloc((void**)&e->reg_n, 2 * sizeof(int)));
    e->reg_cap = need;
  }
  const int h_n[2] = {nq, nt};
  GT_CUDA(e, cudaMemcpyAsync(e->reg_n, h_n, sizeof(h_n), cudaMemcpyHostToDevice, st));
  GT_CUDA(e, cudaStreamSynchronize(st));   // (h_n is a stack buffer)
  desc_expand_f32_kernel<<<ceil_div(nq, 256) * 2, 256, 0, st>>>(q_dev, nq, e->reg_xq, e->reg_cq);   // whole pairs of groups: a CTA stages two query groups
  desc_expand_f32_kernel<<<ceil_div(nt, 128), 256, 0, st>>>(t_dev, nt, e->reg_xt, e->reg_ct);
  MatchArgs a;
  a.xq = e->reg_xq; a.xt = e->reg_xt; a.xq_bstride = a.xt_bstride = 0;
  a.cq = e->reg_cq; a.ct = e->reg_ct; a.cq_bstride = a.ct_bstride = 0;
  a.nq = e->reg_n; a.nt = e->reg_n + 1; a.nq_step = a.nt_step = 0; a.cap = e->reg_cap;
  a.out_a = e->reg_cand; a.out_b = nullptr; a.out_bstride = 0;
  match_tc_kernel<MT_L2><<<dim3((unsigned)ceil_div(nq, 128 * C::QT), 1), kMtThreads, C::SMEM, st>>>(a);
  l2_rerank_kernel<<<ceil_div(nq, 8), 256, 0, st>>>(q_dev, t_dev, nq, e->reg_cand, out_idx_dev, out_dist_dev);
  e->launches += 4;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}
