// Does st.async (STAS) to the CTA's own shared memory work without / with a cluster launch?
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(int use_mapa, uint32_t* out) {
  __shared__ alignas(16) uint32_t buf[4 * 32];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  uint32_t dst = s32(&buf[threadIdx.x * 4]), mb = s32(&bar);
  if (use_mapa) {
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(dst), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(mb) : "r"(mb), "r"(rank));
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(32u * 16u) : "memory");
  __syncthreads();
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(dst), "r"(threadIdx.x), "r"(1u), "r"(2u), "r"(3u), "r"(mb) : "memory");
  uint32_t ok = 0;
  for (int i = 0; i < 100000 && !ok; ++i)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
  out[threadIdx.x] = ok ? buf[threadIdx.x * 4] + buf[threadIdx.x * 4 + 3] : 0xdeadu;
}
int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  uint32_t* d; cudaMalloc(&d, 32 * 4);
  for (int mode = 0; mode < 4; ++mode) {
    if (only >= 0 && mode != only) continue;
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(mode == 3 ? 2 : 1); cfg.blockDim = dim3(32);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = mode == 3 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = mode >= 2 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, mode >= 1 ? 1 : 0, d);
    cudaError_t e2 = cudaDeviceSynchronize();
    uint32_t h[32] = {0}; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("mode %d (%s): launch %s, sync %s, out[5]=%u (want 8)\n", mode, mode == 0 ? "plain launch, cta addr" : mode == 1 ? "plain launch, mapa" : mode == 2 ? "cluster 1, mapa" : "cluster 2, mapa",
           cudaGetErrorString(e), cudaGetErrorString(e2), h[5]);
    if (e2 != cudaSuccess) { cudaDeviceReset(); cudaMalloc(&d, 32 * 4); }
  }
  return 0;
}
