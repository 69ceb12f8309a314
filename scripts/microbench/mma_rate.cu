// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for N = 256 / 128 / 64, one CTA per SM: TS mode (A from
// TMEM), SS mode (A and B SWIZZLE_128B) and SS mode with the no-swizzle, tap-shifted A operand conv_tc_kernel uses.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_nosw(uint32_t a) {   // K-major, no swizzle, LBO 2112 B (a 132-row plane), SBO 128 B
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(2112 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128, 1) k(int n, int ss, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* buf = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)buf)[i] = 0;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tbase;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t bd = desc(s32(buf)), ad = ss == 2 ? desc_nosw(s32(buf + 32768)) : desc(s32(buf + 32768));
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t a_col = (i & 7) * 8;
      if (ss) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm + 256), "l"(ad + (uint64_t)(ss == 2 ? (i & 3) * 264 + (i & 4 ? 1 : 0) : (i & 3) * 2)), "l"(bd + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1));
      else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm + 256), "r"(tm + a_col), "l"(bd + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)));
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)), "r"(0));
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  int iters = 20000;
  for (int ss = 0; ss < 3; ++ss) for (int n : {256, 128, 64}) {
    for (int rep = 0; rep < 2; ++rep) { k<<<148, 128, 70000>>>(n, ss, iters, d); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%s N=%d: %.1f cycles/MMA (ideal %d) err=%s\n", ss == 2 ? "SS (A no-swizzle, tap-shifted)" : ss ? "SS" : "TS", n, avg / iters, n / 2, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
