// Microbenchmark: back-to-back tcgen05.mma (cta_group::1, kind::f16, SS operands) issue rate on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I geo-trax_b200/csrc -o gpurun_out/mma_bench tools/mma_bench.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

__global__ void __launch_bounds__(128, 1) bench(int M, int N, int iters, int per_commit, int sbo_rows, int row_bytes, int layout, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;
  if (threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
    uint32_t ph = 0;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
      t0 = clock64();
      for (int it = 0; it < iters; it += per_commit) {
        if (elect_one()) {
          for (int j = 0; j < per_commit; ++j) {
            // rotate over 4 k-slices and 4 stages like a real pipeline
            const uint32_t st = (uint32_t)((it + j) >> 2) & 3u, ks = (uint32_t)(it + j) & 3u;
            const uint64_t ad = make_desc(a_addr + st * (uint32_t)(M * row_bytes), 8u * row_bytes, layout) + ks * 2;
            const uint64_t bd = make_desc(b_addr + st * (uint32_t)(N / 8 * sbo_rows * row_bytes), (uint32_t)(sbo_rows * row_bytes), layout) + ks * 2;
            umma_f16(tmem_base, ad, bd, idesc, 1u);
          }
          umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), ph);
        ph ^= 1u;
      }
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int M, N, per_commit, sbo_rows, row_bytes, layout; const char* name; };
  const Cfg cfgs[] = {
      {128, 256, 64, 8, 128, 2, "M128 N256 sw128 commit/64"}, {128, 256, 4, 8, 128, 2, "M128 N256 sw128 commit/4"},
      {128, 128, 64, 8, 128, 2, "M128 N128 sw128 commit/64"}, {128, 64, 64, 8, 128, 2, "M128 N64  sw128 commit/64"},
      {128, 32, 64, 8, 128, 2, "M128 N32  sw128 commit/64"},  {64, 256, 64, 8, 128, 2, "M64  N256 sw128 commit/64"},
      {128, 256, 64, 10, 128, 2, "M128 N256 sw128 SBO=10 rows"}, {128, 256, 64, 8, 64, 4, "M128 N256 sw64"},
      {128, 256, 64, 10, 64, 4, "M128 N256 sw64 SBO=10 rows"}, {128, 192, 64, 8, 128, 2, "M128 N192 sw128"},
  };
  for (int grid : {1, 148}) {
    for (const Cfg& c : cfgs) {
      const int iters = 4096;
      bench<<<grid, 128, 200 * 1024>>>(c.M, c.N, iters, c.per_commit, c.sbo_rows, c.row_bytes, c.layout, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("grid %3d  %-30s  %7.1f cycles/MMA  (%s)  -> %.0f FLOP/clk/SM\n", grid, c.name, (double)mx / iters, cudaGetErrorString(e),
             2.0 * c.M * c.N * 16 * iters / (double)mx);
    }
  }
  return 0;
}
