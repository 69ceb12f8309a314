// Microbenchmark: what stretches a TS-mode tcgen05.mma (M=128, N=256, K=16, bf16; 128 cycles ideal) inside the trunk
// kernel?  One CTA per SM, 18 warps:
//   warp 0      issues the MMAs back to back (A from TMEM, B from shared memory) and times them
//   warp 1      (mode & 1) streams 32 KB cp.async.bulk copies from an L2-resident buffer into a 4-stage ring
//   warps 2-17  (mode & 2) epilogue-like TMEM traffic: tcgen05.ld 16 columns -> a few ALU ops -> tcgen05.st
//               (mode & 4) plus one 16-bit st.shared per 16 columns (the sign-bit words)
// Prints cycles per MMA and the achieved side traffic for every mode.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  return ok != 0;
}
constexpr int NT = 18 * 32;
__global__ void __launch_bounds__(NT, 1) k(int mode, int iters, const uint8_t* src, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* buf = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);   // [B tile 32 KB][ring 4 x 32 KB]
  __shared__ uint64_t bar, ring_bar[4];
  __shared__ uint32_t tbase;
  __shared__ volatile int done;
  __shared__ uint16_t masks[16][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;           // (mode & 8): random bf16 pairs in (-2, 2)
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    ((uint32_t*)buf)[i] = (mode & 8) ? ((h & 0x807F807Fu) | 0x3F003F00u) : 0u;
  }
  if (threadIdx.x == 0) {
    done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&ring_bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tbase;
  if ((mode & 8) && warp >= 2 && warp < 6) {                                   // A operand columns 0..63 <- random bf16 pairs
    uint32_t h = threadIdx.x * 2654435761u + 12345u;
    for (int c = 0; c < 64; c += 8) {
      uint32_t r[8];
      for (int i = 0; i < 8; ++i) { h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; r[i] = (h & 0x807F807Fu) | 0x3F003F00u; }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   ::"r"(tm + ((uint32_t)((warp - 2) * 32) << 16) + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t bd = desc(s32(buf));
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t a_col = (i & 7) * 8;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tm + 256), "r"(tm + a_col), "l"(bd + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1));
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)));
      while (!try_wait(&bar, 0)) {}
      long long t1 = clock64();
      out[blockIdx.x * 4 + 0] = t1 - t0;
      done = 1;
    }
  } else if (warp == 1) {
    if (lane == 0 && (mode & 1)) {
      long long n = 0;
      uint32_t phase[4] = {0, 0, 0, 0};
      const long long t0 = clock64();
      while (!done) {
        const int s = (int)(n & 3);
        if (n >= 4) { while (!try_wait(&ring_bar[s], phase[s])) {} phase[s] ^= 1; }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&ring_bar[s])), "r"(32768) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(buf + 32768 + s * 32768)), "l"(src + (size_t)(n % 28) * 32768), "r"(32768), "r"(s32(&ring_bar[s])) : "memory");
        ++n;
      }
      const long long t1 = clock64();
      for (long long j = (n >= 4 ? n - 4 : 0); j < n; ++j) { const int s = (int)(j & 3); while (!try_wait(&ring_bar[s], phase[s])) {} phase[s] ^= 1; }
      out[blockIdx.x * 4 + 1] = n * 32768;
      out[blockIdx.x * 4 + 2] = t1 - t0;
    }
  } else if (mode & 2) {
    const int e = warp - 2, q = e & 3, hq = e >> 2;
    const uint32_t addr = tm + ((uint32_t)(q * 32) << 16) + 64 + hq * 48;
    long long n = 0;
    while (!__shfl_sync(0xffffffffu, done, 0)) {
#pragma unroll 1
      for (int kb = 0; kb < 3; ++kb) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(addr + kb * 16) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) { float v = __uint_as_float(r[i]) + 1.0f; m |= (v > 0.f ? 1u : 0u) << i; r[i] = __float_as_uint(fmaxf(v, 0.f)); }
        if (mode & 4) masks[e][q * 32 + lane] = (uint16_t)m;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(addr + kb * 16), "r"(r[0] ^ m), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      ++n;
    }
    if (e == 0 && lane == 0) out[blockIdx.x * 4 + 3] = n * 3 * 16 * 16 * 128 * 4;   // TMEM bytes loaded by all 16 warps
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 4 * 8);
  uint8_t* src; cudaMalloc(&src, 28 * 32768); cudaMemset(src, 0, 28 * 32768);
  const int smem = 1024 + 5 * 32768;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 200000;
  const char* names[16] = {"MMA alone", "+ bulk ring", "+ TMEM ld/st", "+ bulk + TMEM", "", "", "+ TMEM + masks", "+ bulk + TMEM + masks", "random data: MMA alone", "random: + bulk", "random: + TMEM", "random: + bulk + TMEM", "", "", "", "random: + bulk + TMEM + masks"};
  for (int mode : {0, 1, 2, 3, 6, 7, 8, 9, 10, 11, 15}) {
    for (int rep = 0; rep < 2; ++rep) { cudaMemset(d, 0, 148 * 4 * 8); k<<<148, NT, smem>>>(mode, iters, src, d); cudaDeviceSynchronize(); }
    long long h[148 * 4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mma = 0, bytes = 0, cyc = 0, tm = 0;
    for (int i = 0; i < 148; ++i) { mma += h[i * 4]; bytes += h[i * 4 + 1]; cyc += h[i * 4 + 2]; tm += h[i * 4 + 3]; }
    printf("%-24s %.1f cycles/MMA | bulk %.1f B/clk/SM | TMEM ld %.1f B/clk/SM | %s\n", names[mode], mma / 148 / iters,
           cyc > 0 ? bytes / cyc : 0.0, tm / (mma / 148) / 148, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
