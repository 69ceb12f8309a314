// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM with 4, 8, 16 warps.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(512, 1) k(int iters, int mode, long long* out) {
  __shared__ uint32_t tbase;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = tbase;
  int warp = threadIdx.x >> 5;
  uint32_t addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
  uint32_t r[32];
  for (int i = 0; i < 32; ++i) r[i] = i;
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(addr + (uint32_t)((i & 3) * 128)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[i & 31];
    } else {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(addr + (uint32_t)((i & 3) * 128)), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = (t1 - t0) + (acc == 0xdeadbeef);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 16 * 8);
  int iters = 4000;
  for (int mode = 0; mode < 2; ++mode) for (int nw : {4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) { k<<<148, nw * 32>>>(iters, mode, d); cudaDeviceSynchronize(); }
    long long h[148 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
    double bytes = (double)iters * nw * 32 * (mode == 0 ? 32 : 16) * 4;
    printf("%s warps=%d: %.1f cycles/instr/warp, %.1f B/clk/SM  (%s)\n", mode ? "STTM.x16" : "LDTM.x32+wait", nw, mx / iters, bytes / mx, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
