// Microbenchmark + known-answer test for the CTA-pair MMA (tcgen05.mma.cta_group::2, M = 256 over two SMs, N = 256,
// K = 16, bf16, both operands in shared memory) that a two-tile SS-mode trunk would need:
//   1. correctness: D = A . B^T over one K = 64 k-block with random data, A rows split 128 / 128 over the two CTAs,
//      B rows (N) split 128 / 128 -- checks the operand split and the TMEM row ownership against a host reference;
//   2. rate: cycles per MMA issued back to back, alone and under the shared-memory side traffic of the real kernel
//      (mode & 1: 16 KB cp.async.bulk weight refills per 4 MMAs in each CTA; mode & 2: 16 KB of st.shared per 4 MMAs;
//       mode & 4: tcgen05.ld of 64 accumulator columns per 4 MMAs), and the same for the single-CTA M = 128 MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_2cta mma_2cta.cu && ./mma_2cta
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  return ok != 0;
}
// non-blocking poll (try_wait may suspend the thread for a while before it returns false)
__device__ __forceinline__ bool test_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int NT = 12 * 32;       // warps 0-3: TMEM readers, 4: MMA, 5: bulk copies, 6-11: st.shared traffic
constexpr int A_BYTES = 128 * 128, BH_BYTES = 128 * 128;     // per CTA: 128 A rows, 128 B rows, one k-block each

struct Params {
  const uint8_t* a_img;    // [2][128 rows x 128 B] swizzled images, one per CTA rank
  const uint8_t* b_img;    // [2][128 rows x 128 B]
  float* d_out;            // [256][256]
  const uint8_t* src;      // L2-resident source of the refill copies
  long long* cycles;       // per cluster: cycles of the MMA sequence
  int iters, mode, two_cta, pace;
};

template <int NCTA>
__global__ void __launch_bounds__(NT, 1) k(Params P) {
  extern __shared__ uint8_t raw[];
  uint8_t* buf = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);     // [A 16 KB][B 16/32 KB][refill ring 4 x 16 KB][scratch 16 KB]
  __shared__ uint64_t bar, ring_bar[4];
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = NCTA == 2 ? cta_rank() : 0u;
  uint8_t* sA = buf;
  uint8_t* sB = buf + A_BYTES;
  uint8_t* ring = sB + 2 * BH_BYTES;
  uint8_t* scratch = ring + 4 * 16384;
  // operands: this CTA's A rows; B: its N-half (2-CTA) or both halves (1-CTA)
  for (int i = threadIdx.x; i < A_BYTES / 16; i += NT) ((uint4*)sA)[i] = ((const uint4*)(P.a_img + (size_t)rank * A_BYTES))[i];
  if (NCTA == 2) {
    for (int i = threadIdx.x; i < BH_BYTES / 16; i += NT) ((uint4*)sB)[i] = ((const uint4*)(P.b_img + (size_t)rank * BH_BYTES))[i];
  } else {
    for (int i = threadIdx.x; i < 2 * BH_BYTES / 16; i += NT) ((uint4*)sB)[i] = ((const uint4*)P.b_img)[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&ring_bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 4) {
    if (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tbase)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tbase)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (NCTA == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tbase;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)((128 * NCTA) >> 4) << 24);

  if (warp == 4 && rank == 0) {
    if (lane == 0) {
      // mode & 32: the issuer-side work of the real kernel between k-blocks, selected by the bits of `pace`
      const int ov = (P.mode & 32) ? P.pace : 0;
      const long long t0 = clock64();
      if (P.mode & 8) { while (clock64() - t0 < (long long)P.iters * 4 * 128) { } }
      if (ov) {   // complete phase 0 of ring_bar[2], ring_bar[3] so that parity-0 waits on them succeed at once
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&ring_bar[2])) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&ring_bar[3])) : "memory");
      }
      uint32_t sink = 0;
      for (int it = 0; it < ((P.mode & 8) ? 0 : P.iters); ++it) {
        if ((ov & 128) && (it & 1)) goto mmas;          // bit 7: the sync work only every other k-block (8 MMAs per round)
        if (ov & 1) { while (!try_wait(&ring_bar[2], 0)) { } while (!try_wait(&ring_bar[3], 0)) { } }
        if (ov & 32) { while (!try_wait(&ring_bar[2], 0)) { } }
        if (ov & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ov & 4) { sink += test_wait(&ring_bar[2], 0); sink += test_wait(&ring_bar[3], 0); }
      mmas:
        const uint32_t rot = (ov & 16) ? (uint32_t)(it & 1) * 0u + (uint32_t)(it % 3 == 7) : 0u;   // runtime value, always 0
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = sw128(s32(sA) + rot * 16384 + ks * 32), bd = sw128(s32(sB) + rot * 16384 + ks * 32);
          const uint32_t acc = (it | ks) ? 1u : 0u;
          if (NCTA == 2)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        if ((ov & 128) && !(it & 1)) continue;
        if (ov & 64) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&ring_bar[0])) : "memory");
        if (ov & 8) {
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&ring_bar[0])) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&ring_bar[1])) : "memory");
        }
        if (P.mode & 16) {                       // extra commits per 4 MMAs (to barriers nobody waits on)
          for (int c = 0; c < P.pace; ++c)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&ring_bar[c & 3])) : "memory");
        }
      }
      if (NCTA == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(s32(&bar)), "h"((uint16_t)3) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      while (!try_wait(&bar, 0)) { }
      P.cycles[blockIdx.x / NCTA] = clock64() - t0;
      if (sink == 0xdeadbeefu) P.cycles[1028] = 1;
    }
  } else if (warp == 5 && (P.mode & 1)) {
    // weight-refill traffic: 16 KB per 4 MMAs per CTA (2-CTA: each CTA stages its half) or 32 KB (1-CTA), 4 in flight
    if (lane == 0) {
      const uint32_t bytes = NCTA == 2 ? 16384u : 32768u;   // 1-CTA: two 16 KB copies per step
      uint32_t ph[4] = {0, 0, 0, 0};
      int n = 0;
      for (int it = 0; !test_wait(&bar, 0); ++it) {
        const int s = it & 3;
        if (it >= 4) { while (!try_wait(&ring_bar[s], ph[s])) { } ph[s] ^= 1; }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&ring_bar[s])), "r"(bytes) : "memory");
        for (uint32_t o = 0; o < bytes; o += 16384)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(s32(ring + s * 16384)), "l"(P.src + ((size_t)(it * 37 + blockIdx.x) % 64) * 32768 + o), "r"(16384u),
                         "r"(s32(&ring_bar[s])) : "memory");
        ++n;
      }
      if (blockIdx.x == 0) P.cycles[1024] = n;
      // drain
      for (int s = 0; s < 4 && s < n; ++s) { while (!try_wait(&ring_bar[s], ph[s])) { } }
    }
  } else if (warp >= 6 && (P.mode & 2)) {
    // epilogue-like st.shared traffic into a scratch k-block (16-byte stores, conflict-free swizzled pattern);
    // the barrier is polled every 8 rounds, `pace` dependent FMAs between rounds set the rate
    int n = 0;
    float dummy = (float)threadIdx.x;
    while (!test_wait(&bar, 0)) {
      for (int rep = 0; rep < 8; ++rep) {
        const int row = ((warp - 6) * 32 + lane + rep * 13) & 127;
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(scratch + row * 128 + ((j ^ (row & 7)) << 4)) = make_uint4(n, j, row, 0);
        for (int d = 0; d < P.pace; ++d) dummy = fmaf(dummy, 1.0001f, 0.5f);
        ++n;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 6 * 32) P.cycles[1025] = n;
    if (dummy == 123.f) P.cycles[1027] = 1;
  } else if (warp < 4 && (P.mode & 4)) {
    int n = 0;
    uint32_t acc = 0;
    while (!test_wait(&bar, 0)) {
      for (int rep = 0; rep < 8; ++rep) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tm + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)((n & 15) * 16)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= r[0] ^ r[7];
        ++n;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) P.cycles[1026] = n;
    if (acc == 0x12345678u) P.cycles[1027] = n;
  }
  // everyone: wait for the MMA sequence, then read the accumulator rows of this CTA back
  while (!try_wait(&bar, 0)) { }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4 && P.d_out && blockIdx.x < NCTA) {
    const int row = warp * 32 + lane;
    for (int c = 0; c < 256; c += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) P.d_out[(size_t)(rank * 128 + row) * 256 + c + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (NCTA == 2) cluster_sync();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

static uint16_t f2bf(float f) { __nv_bfloat16 h = __float2bfloat16(f); return *reinterpret_cast<uint16_t*>(&h); }
static float bf2f(uint16_t u) { uint32_t w = (uint32_t)u << 16; float f; memcpy(&f, &w, 4); return f; }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <int NCTA>
int launch(Params P, int grid, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = NCTA; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaFuncSetAttribute(k<NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaLaunchKernelEx(&cfg, k<NCTA>, P));
  CK(cudaDeviceSynchronize());
  return 0;
}

int main() {
  const int M = 256, N = 256, K = 64;
  std::vector<float> A(M * K), B(N * K);
  srand(1);
  for (auto& v : A) v = bf2f(f2bf((float)rand() / RAND_MAX * 2.f - 1.f));
  for (auto& v : B) v = bf2f(f2bf((float)rand() / RAND_MAX * 2.f - 1.f));
  // swizzled images: row r (within a 128-row half), 16-byte chunk j at r*128 + ((j ^ (r & 7)) << 4)
  std::vector<uint16_t> a_img(M * K), b_img(N * K);
  for (int r = 0; r < M; ++r) for (int kk = 0; kk < K; ++kk) {
    const int half = r / 128, rr = r % 128, j = kk / 8, e = kk % 8;
    a_img[(size_t)half * 8192 + rr * 64 + ((j ^ (rr & 7)) * 8) + e] = f2bf(A[r * K + kk]);
    b_img[(size_t)half * 8192 + rr * 64 + ((j ^ (rr & 7)) * 8) + e] = f2bf(B[r * K + kk]);
  }
  uint8_t *da, *db, *src; float* dd; long long* dc;
  CK(cudaMalloc(&da, a_img.size() * 2)); CK(cudaMalloc(&db, b_img.size() * 2)); CK(cudaMalloc(&dd, M * N * 4));
  CK(cudaMalloc(&src, 64 * 32768)); CK(cudaMemset(src, 0, 64 * 32768)); CK(cudaMalloc(&dc, 2048 * 8));
  CK(cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = 1024 + A_BYTES + 2 * BH_BYTES + 4 * 16384 + 16384;
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

  // 1. known-answer test, one cluster (2-CTA) / two CTAs... the 1-CTA form computes rows 0..127 only
  for (int ncta = 2; ncta >= 1; --ncta) {
    CK(cudaMemset(dd, 0, M * N * 4));
    Params P{da, db, dd, src, dc, 1, 0, ncta == 2, 0};
    if ((ncta == 2 ? launch<2>(P, 2, smem) : launch<1>(P, 1, smem))) return 1;
    std::vector<float> D(M * N);
    CK(cudaMemcpy(D.data(), dd, M * N * 4, cudaMemcpyDeviceToHost));
    double worst = 0; int bad = 0;
    for (int r = 0; r < 128 * ncta; ++r) for (int c = 0; c < N; ++c) {
      double ref = 0; for (int kk = 0; kk < K; ++kk) ref += (double)A[r * K + kk] * B[c * K + kk];
      const double e = fabs(ref - D[r * N + c]);
      if (e > worst) worst = e;
      if (e > 1e-3) ++bad;
    }
    printf("known-answer %d-CTA M=%d N=256 K=64: max abs err %.3e, %d bad of %d\n", ncta, 128 * ncta, worst, bad, 128 * ncta * N);
  }
  // 2. rate
  const int iters = 4096;
  for (int commits = 0; commits <= 4; ++commits) {
    CK(cudaMemset(dc, 0, 2048 * 8));
    Params P{da, db, nullptr, src, dc, iters, 16, 0, commits};
    if (launch<1>(P, sms, smem)) return 1;
    std::vector<long long> c(2048);
    CK(cudaMemcpy(c.data(), dc, 2048 * 8, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < sms; ++i) cyc += (double)c[i];
    printf("1-CTA M=128, %d tcgen05.commit per 4 MMAs -> %.1f cycles/MMA\n", commits, cyc / sms / (iters * 4));
  }
  for (int ov : {0, 1 + 2 + 8, 32 + 2 + 64, 128 + 1 + 2 + 8, 128 + 32 + 2 + 64, 32 + 2, 64, 1 + 2 + 64, 32 + 2 + 8}) {
    CK(cudaMemset(dc, 0, 2048 * 8));
    Params P{da, db, nullptr, src, dc, iters, 32, 0, ov};
    if (launch<1>(P, sms, smem)) return 1;
    std::vector<long long> c(2048);
    CK(cudaMemcpy(c.data(), dc, 2048 * 8, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < sms; ++i) cyc += (double)c[i];
    printf("1-CTA M=128, issuer overhead per %d MMAs [%s%s%s%s%s%s%s] -> %.1f cycles/MMA\n", (ov & 128) ? 8 : 4, (ov & 32) ? "1 try_wait " : "", (ov & 64) ? "1 commit " : "", (ov & 1) ? "2 try_wait " : "", (ov & 2) ? "fence " : "",
           (ov & 4) ? "2 test_wait " : "", (ov & 8) ? "2 commit " : "", (ov & 16) ? "runtime-desc " : "", cyc / sms / (iters * 4));
  }
  for (int ncta = 1; ncta <= 1; ++ncta)
    for (int mode : {0})
      for (int pace = 0; pace <= ((mode & 2) ? 90 : 0); pace += 30) {
      CK(cudaMemset(dc, 0, 2048 * 8));
      Params P{da, db, nullptr, src, dc, iters, mode, ncta == 2, pace};
      const int grid = ncta == 2 ? (sms / 2) * 2 : sms;
      if ((ncta == 2 ? launch<2>(P, grid, smem) : launch<1>(P, grid, smem))) return 1;
      std::vector<long long> c(2048);
      CK(cudaMemcpy(c.data(), dc, 2048 * 8, cudaMemcpyDeviceToHost));
      double cyc = 0; int ncl = grid / ncta;
      for (int i = 0; i < ncl; ++i) cyc += (double)c[i];
      cyc /= ncl;
      printf("%d-CTA M=%d%s: %s%s%s(pace %d) -> %.1f cycles/MMA (ideal 128) | refill %.1f B/clk, st.shared %.1f B/clk, tcgen05.ld %.1f B/clk\n",
             ncta, 128 * ncta, (mode & 8) ? " NO MMA" : "", (mode & 1) ? "+refill " : "", (mode & 2) ? "+st.shared " : "", (mode & 4) ? "+tcgen05.ld " : "", pace,
             cyc / (iters * 4), (double)c[1024] * (ncta == 2 ? 16384 : 32768) / cyc, (double)c[1025] * 6 * 32 * 128 / cyc,
             (double)c[1026] * 4 * 32 * 64 / cyc);
    }
  return 0;
}
