// Per-SM ingest probe (B200): how many bytes per clock can ONE SM pull from L2 into shared memory
//   mode 0: TMA bulk copies (cp.async.bulk, one elected thread, 4 x 16 KB in flight)
//   mode 1: LDGSTS (cp.async.cg 16 B per thread, 256 threads, 4 groups in flight)
//   mode 2: both at the same time (warp 0 drives the bulk copies, warps 1-8 the cp.async stream)
// Question behind it: the tcgen05 GEMMs of this repo are bound by ~38 B/clk/SM of TMA traffic -- would moving one operand to the
// LSU path (cp.async) add bandwidth, or do both paths share one L2 -> SM port?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/ingest_probe.bin tools/exp/ingest_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int CHUNK = 16384, STAGES = 4;

__global__ void __launch_bounds__(288, 1) k_probe(const uint8_t* __restrict__ src, size_t per_cta, int iters, int mode, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tma_buf = smem;                              // STAGES x CHUNK
  uint8_t* lsu_buf = smem + STAGES * CHUNK;             // STAGES x CHUNK
  uint64_t* bars = (uint64_t*)(smem + 2 * STAGES * CHUNK);
  const uint8_t* mine = src + (size_t)blockIdx.x * per_cta;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    for (int i = 0; i < 2 * STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0 || (warp == 1 && mode == 3)) {
    if ((mode == 0 || mode == 2 || mode == 3) && (threadIdx.x & 31) == 0) {
      if (warp == 1) { tma_buf = lsu_buf; bars += STAGES; mine += per_cta / 2; }
      uint32_t phase[STAGES] = {0, 0, 0, 0};
      for (int it = 0; it < iters + STAGES; ++it) {
        const int st = it % STAGES;
        if (it >= STAGES) {                             // wait for the copy issued STAGES iterations ago
          uint32_t ok = 0;
          while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bars[st])), "r"(phase[st]) : "memory");
          phase[st] ^= 1;
        }
        if (it < iters) {
          const size_t off = ((size_t)it * CHUNK) % per_cta;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[st])), "r"(CHUNK) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(tma_buf + st * CHUNK)), "l"(mine + off), "r"(CHUNK), "r"(smem_u32(&bars[st])) : "memory");
        }
      }
    }
  } else if ((mode == 1 || mode == 2) && warp >= 1) {
    const int t = threadIdx.x - 32;                     // 256 loader threads: 4 KB per round, 4 rounds per 16 KB chunk
    const size_t half = per_cta / 2;
    for (int it = 0; it < iters; ++it) {
      const int st = it % STAGES;
      const size_t off = (half + (size_t)it * CHUNK) % per_cta;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(lsu_buf + st * CHUNK + r * 4096 + t * 16)),
                     "l"(mine + off + r * 4096 + t * 16) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 3;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  const long long t_mine = clock64() - t0;             // (per-thread: the loaders and the TMA thread finish at different times)
  __shared__ unsigned long long s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  atomicMax(&s_max, (unsigned long long)t_mine);
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = s_max;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t per_cta = 512 << 10;                    // 512 KB per CTA, 74 MB in all: L2-resident after the first pass
  uint8_t* src;
  unsigned long long* cyc;
  cudaMalloc(&src, per_cta * sms);
  cudaMemset(src, 1, per_cta * sms);
  cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  const int smem = 2 * STAGES * CHUNK + 256;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  for (int grid : {sms, sms / 4}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        k_probe<<<grid, 288, smem>>>(src, per_cta, iters, mode, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long h[256];
      cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)iters * CHUNK * (mode >= 2 ? 2 : 1);
      printf("grid %3d mode %d (%s): %.1f B/clk/SM (slowest CTA %.0f cycles)\n", grid, mode,
             mode == 0 ? "TMA bulk" : mode == 1 ? "cp.async" : mode == 2 ? "both" : "TMA bulk from two issuing threads", bytes / mx, mx);
    }
  }
  return 0;
}
