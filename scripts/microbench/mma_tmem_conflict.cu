// Microbenchmark (round 2): does epilogue-style tcgen05.ld / tcgen05.st traffic slow a TS-mode tcgen05.mma down, and
// does it depend on WHICH TMEM columns the two touch and on the MMA's N?  The trunk timeline showed N = 128 MMAs at
// 62-64 cycles while the epilogue warps were idle and ~110 cycles while they were converting the other half of the same
// 256-column accumulator region.
//   warp 0      one lane issues M=128 x N x K=16 bf16 MMAs back to back: A from TMEM columns [a_col, a_col+64),
//               D at d_col (optionally alternating with d_col+128 every 4 MMAs), B from one resident smem tile
//   warps 2..   n_epi epilogue warps: tcgen05.ld 16 columns -> ALU -> tcgen05.st 8 (+8) columns, wait::st, over the
//               128 columns starting at epi_col (warp (q, hq): lanes 32q.., columns epi_col + 32hq + 16kb, kb < 2)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tmem_conflict mma_tmem_conflict.cu && ./mma_tmem_conflict
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
struct Cfg { int d_col, d_alt, a_col, epi_col, n_epi, st_cols; };
template <int N>
__global__ void __launch_bounds__(NT, 1) k(Cfg c, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* buf = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)buf)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
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
  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t bd = desc(s32(buf));
      const uint32_t d0 = (uint32_t)c.d_col, a0 = (uint32_t)c.a_col, alt = (uint32_t)c.d_alt;
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t a_col = a0 + (uint32_t)(i & 7) * 8u;
        const uint32_t d = d0 + (alt ? (uint32_t)((i >> 2) & 1) * 128u : 0u);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d), "r"(a_col), "l"(bd + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1));
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)));
      while (!try_wait(&bar, 0)) {}
      out[blockIdx.x * 2 + 0] = clock64() - t0;
      done = 1;
    }
  } else if (warp >= 2 && warp - 2 < c.n_epi) {
    const int e = warp - 2, q = e & 3, hq = e >> 2;
    const uint32_t addr = ((uint32_t)(q * 32) << 16) + (uint32_t)c.epi_col + hq * 32;
    long long n = 0;
    while (!__shfl_sync(0xffffffffu, done, 0)) {
#pragma unroll 1
      for (int kb = 0; kb < 2; ++kb) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(addr + kb * 16) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]) + 1.0f, 0.f));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(addr + kb * 16), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        if (c.st_cols == 16)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                       ::"r"(addr + kb * 16 + 8), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      ++n;
    }
    if (e == 0 && lane == 0) out[blockIdx.x * 2 + 1] = n * 2 * 16 * 128 * 4 * c.n_epi / 4;   // TMEM bytes loaded (all epilogue warps)
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 2 * 8);
  const int smem = 1024 + 32768;
  cudaFuncSetAttribute(k<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 100000;
  struct Case { const char* name; int n; Cfg c; };
  const Case cases[] = {
      {"N=256 D=R1         A=R0[0,64)    no epilogue", 256, {256, 0, 0, 128, 0, 8}},
      {"N=256 D=R1         A=R0[0,64)    epi 16w on R0[128,256)", 256, {256, 0, 0, 128, 16, 8}},
      {"N=256 D=R1         A=R0[0,64)    epi 16w on R0[128,256) st16", 256, {256, 0, 0, 128, 16, 16}},
      {"N=128 D=R1 alt     A=R0[0,64)    no epilogue", 128, {256, 1, 0, 128, 0, 8}},
      {"N=128 D=R1 alt     A=R0[0,64)    epi 16w on R0[128,256)", 128, {256, 1, 0, 128, 16, 8}},
      {"N=128 D=R1[128,256) A=R0[0,64)   epi 16w on R1[0,128)   (slots 6,7)", 128, {384, 0, 0, 256, 16, 8}},
      {"N=128 D=R1[128,256) A=R0[0,64)   epi 16w on R1[0,128) st16 (slots 6,7 fp32-grade)", 128, {384, 0, 0, 256, 16, 16}},
      {"N=128 D=R1[128,256) A=R0[128,192) epi 16w on R1[0,128) st16", 128, {384, 0, 128, 256, 16, 16}},
      {"N=128 D=R0[0,128)  A=R1[0,64)    epi 16w on R1[128,256) st16 (next slot 0)", 128, {0, 0, 256, 384, 16, 16}},
      {"N=128 D=R1[128,256) A=R0[0,64)   epi  8w on R1[0,128) st16", 128, {384, 0, 0, 256, 8, 16}},
      {"N=128 D=R1[128,256) A=R0[0,64)   epi  4w on R1[0,128) st16", 128, {384, 0, 0, 256, 4, 16}},
      {"N=256 D=R1         A=R0[0,64)    epi  8w on R0[128,256) st16", 256, {256, 0, 0, 128, 8, 16}},
  };
  for (const Case& cs : cases) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d, 0, 148 * 2 * 8);
      if (cs.n == 128) k<128><<<148, NT, smem>>>(cs.c, iters, d); else k<256><<<148, NT, smem>>>(cs.c, iters, d);
      cudaDeviceSynchronize();
    }
    long long h[148 * 2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mma = 0, tm = 0;
    for (int i = 0; i < 148; ++i) { mma += h[i * 2]; tm += h[i * 2 + 1]; }
    printf("%-84s %.1f cycles/MMA (ideal %d) | TMEM ld %.1f B/clk/SM | %s\n", cs.name, mma / 148 / iters, cs.n / 2,
           tm / mma, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
