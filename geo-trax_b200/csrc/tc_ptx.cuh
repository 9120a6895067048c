// PTX wrappers shared by the tcgen05 convolution kernels (sm_100a): mbarrier, TMA loads/stores, UMMA descriptors + issue,
// TMEM loads, elect.sync, and the division-free persistent tile iterator.
#pragma once
#include "common.cuh"

constexpr int kEpiWarps = 8;

// Programmatic dependent launch: consecutive conv layers are launched with programmaticStreamSerialization, so a layer's CTAs can
// run their prologue while the previous layer drains; pdl_wait() blocks until the previous grid has completed and flushed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
// Halo taps use the same descriptor with a start address that is a whole number of rows into a swizzle atom.  The hardware
// applies the swizzle XOR on absolute shared-memory address bits (measured: base-offset field 0 reproduces torch.conv2d
// bit-for-bit-equivalently for every shift, setting it to (addr >> 7) & 7 does not), so no base offset is encoded.

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
// ---- lean issue helpers: descriptors as (running 32-bit low word, constant high word) ----------------------------------------
// The single issuing lane is latency bound: with general 64-bit descriptor arithmetic the issue block of one k-block cost ~360
// cycles (measured with clock64), more than the 2-4 MMAs of the block take on the tensor pipe (135 cycles for M128 N256 K16,
// 71 for N128, 55 for N <= 64: tools/mma_bench.cu), so the pipe idled.  High word = SBO | version | layout.
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) { return (sbo_bytes >> 4) | (1u << 14) | (layout << 29); }
// D (+)= A . B^T with descriptors given as (low word, high word); ACC = 0: overwrite, 1: accumulate, 2: runtime flag `acc`
template <int ACC>
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc = 1u) {
  if constexpr (ACC == 2) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
        : "memory");
  }
}
// one k-block: MPK MMAs of K = 16 (32 bytes = 2 descriptor units apart); the first one accumulates iff `acc_first`
template <int MPK>
__device__ __forceinline__ void issue_kb(uint32_t tacc, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc_first) {
  umma_lohi<2>(tacc, a_lo, a_hi, b_lo, b_hi, idesc, acc_first);
#pragma unroll
  for (int k = 1; k < MPK; ++k) umma_lohi<1>(tacc, a_lo + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc);
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA-pair helpers (cta_group::2: the two CTAs of a cluster on one TPC run ONE M = 256 MMA; each CTA stages its own 128 rows of A
// and its own half of B's N rows, the accumulator rows of a CTA's A half land in that CTA's TMEM) --------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of both CTAs, warps converged
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads whose completion bytes are counted on a barrier that may live in the peer CTA (`cluster_bar` = shared::cluster address)
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t cluster_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
template <int ACC>
__device__ __forceinline__ void umma_lohi_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc = 1u) {
  if constexpr (ACC == 2) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
        : "memory");
  }
}
template <int MPK>
__device__ __forceinline__ void issue_kb_pair(uint32_t tacc, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc_first) {
  umma_lohi_pair<2>(tacc, a_lo, a_hi, b_lo, b_hi, idesc, acc_first);
#pragma unroll
  for (int k = 1; k < MPK; ++k) umma_lohi_pair<1>(tacc, a_lo + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc);
}
// arrives (once all MMAs issued so far have completed) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((unsigned short)3)
               : "memory");
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
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int CW>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* v) {
  if constexpr (CW == 32) tmem_ld_x32(taddr, v);
  else if constexpr (CW == 16) tmem_ld_x16(taddr, v);
  else tmem_ld_x8(taddr, v);
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {   // true in exactly one lane of the (converged) warp
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// SiLU in 5 instructions: FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL (flush-to-zero approximations, ~1e-6 relative error)
__device__ __forceinline__ float silu_fast(float a) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return a * r;
}
// SiLU with ONE MUFU op: a * (0.5 + 0.5 * tanh(a / 2)) -- FMUL, MUFU.TANH, FFMA, FMUL.  tanh.approx.f32 has an absolute error of 2^-11, so
// sigma is off by <= 2.4e-4 absolute: comparable to the fp16 rounding of the stored output (4.9e-4 relative) for positive a, larger in
// relative terms on the negative tail.  Used where the epilogue is MUFU bound (ConvParams::act == 2: the 544x960 / 272x480 layers,
// 2 MUFU x 50 M outputs per frame = 345 us per step at 16 / clk / SM) and accounted for in the raw-head gate (tests + DESIGN.md 4.2).
__device__ __forceinline__ float silu_tanh(float a) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(a * 0.5f));
  return a * fmaf(0.5f, t, 0.5f);
}
// activation of N register values; the (warp-uniform) mode is tested ONCE, outside the unrolled loops -- a per-element `mode == 1 ? .. : ..`
// makes the compiler evaluate both forms and select (3 MUFU per value: the swapped kernels got 10-70 % slower that way)
template <int N>
__device__ __forceinline__ void act_inplace(float* f, int mode) {
  if (mode == 2) {
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = silu_tanh(f[i]);
  } else if (mode == 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = silu_fast(f[i]);
  }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

struct TileCoord { int n, y0, x0, n0; };
// Walks tile = blockIdx.x, += gridDim.x over the (image, tile row, tile column, N tile) space without per-tile divisions.
struct TileIter {
  int nt, tx, ty, n;       // current coordinates
  int d_nt, d_tx, d_ty, d_n;  // gridDim.x decomposed in the same mixed radix
  __device__ __forceinline__ void init(const ConvParams& p, int first, int step) {
    nt = first % p.n_tiles; int m = first / p.n_tiles;
    tx = m % p.tiles_x; m /= p.tiles_x;
    ty = m % p.tiles_y; n = m / p.tiles_y;
    d_nt = step % p.n_tiles; m = step / p.n_tiles;
    d_tx = m % p.tiles_x; m /= p.tiles_x;
    d_ty = m % p.tiles_y; d_n = m / p.tiles_y;
  }
  __device__ __forceinline__ void next(const ConvParams& p) {
    nt += d_nt; if (nt >= p.n_tiles) { nt -= p.n_tiles; ++tx; }
    tx += d_tx; if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
    ty += d_ty; if (ty >= p.tiles_y) { ty -= p.tiles_y; ++n; }
    n += d_n;
  }
  __device__ __forceinline__ TileCoord coord(const ConvParams& p) const {
    TileCoord t;
    t.n = n + p.img0; t.x0 = tx * p.tw; t.y0 = ty * p.th; t.n0 = nt * p.BN;
    return t;
  }
};

