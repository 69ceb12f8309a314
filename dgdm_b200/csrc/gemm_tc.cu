// Tensor-core GEMM with on-the-fly operand splitting:  C = act(A . W^T + bias), fp32 in / fp32 out,
// computed with tcgen05 kind::f16 MMAs on bf16 hi(+lo) splits (3 MMAs per product in fp32-grade mode).
//
// Used for every large contraction of the denoiser (K3): the Conv1d / strided Conv1d / ConvTranspose1d
// implicit GEMMs of ConditionalUnet1D (generator/diffusion_utils.py:42,51,66,96) read their A rows straight
// out of the zero-padded channels-last activation buffers through the same (sample, row, tap) mapping as the
// CUDA-core GEMM (GemmArgs, common.cuh) -- no im2col.
//
// One persistent CTA per SM, 14 warps, 128 output rows per tile, all N (128 or 256) columns per tile:
//   warps 4-11 A producers: per 64-wide k-block the 256 threads read the [128 x 64] fp32 block with coalesced
//              128-bit loads (lanes along K), split to bf16 hi/lo and write it into shared memory in the canonical
//              K-major SWIZZLE_128B layout (16-byte chunk j of row r at chunk j^(r&7)), then fence.proxy.async +
//              mbarrier arrive; loads run one k-block ahead of the conversion.
//   warp 12    W producer: cp.async.bulk (TMA unit) of the pre-packed, pre-swizzled weight tiles (gemm_tc_pack).
//   warp 13    MMA issuer: tcgen05.mma.cta_group::1.kind::f16, A and B from shared memory, M=128 x N x K=16,
//              fp32 accumulators in TMEM, double-buffered (2 x 256 columns) so the epilogue of tile i overlaps
//              the main loop of tile i+1.
//   warps 0-3  epilogue: tcgen05.ld, + bias, activation, optional mask (zero where mask <= 0), 128-bit stores through
//              the C row mapping.
#include <cuda_bf16.h>

#include "common.cuh"

namespace dgdm {
namespace {

constexpr int TM = 128, KB = 64, NTHR = 448;   // 4 epilogue + 8 A-producer + W-producer + MMA warps
constexpr int A_TILE = TM * 128;               // 16 KB: 128 rows x 128 B
constexpr int SMEM_BUDGET = 200 * 1024;
constexpr int MAX_STAGE = 4;

struct TcGemmParams {
  GemmArgs g;
  const uint8_t* wimg;
  int* err;
  int x3, n_tiles, n_kb, stage_bytes, n_stage;
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c));
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity, int* err, int code) {
  uint32_t a = s32(b);
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (it > (1u << 24)) { if (err) atomicExch(err, code); __trap(); }
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, SBO = 1024 B, version 1
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Bars {
  uint64_t full_a[MAX_STAGE], full_w[MAX_STAGE], empty[MAX_STAGE], d_full[2], d_empty[2];
  uint32_t tmem_base, pad_;
};

__global__ void __launch_bounds__(NTHR, 1) gemm_tc_kernel(const __grid_constant__ TcGemmParams P) {
  extern __shared__ uint8_t raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  Bars& S = *reinterpret_cast<Bars*>(ring + (size_t)P.n_stage * P.stage_bytes);
  const GemmArgs& g = P.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = g.N;
  const uint32_t w_tile = (uint32_t)N * 128u;
  // stage layout: [A_hi][A_lo (x3)][W_hi][W_lo (x3)]
  const uint32_t off_alo = A_TILE, off_whi = P.x3 ? 2 * A_TILE : A_TILE, off_wlo = off_whi + w_tile;

  if (tid == 0) {
    for (int s = 0; s < P.n_stage; ++s) { bar_init(&S.full_a[s], 8); bar_init(&S.full_w[s], 1); bar_init(&S.empty[s], 1); }
    for (int i = 0; i < 2; ++i) { bar_init(&S.d_full[i], 1); bar_init(&S.d_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;
  // all 512 columns allocated => base 0; the MMA issuer uses literal addresses so its operands stay warp-uniform
  if (tmem != 0u) { if (tid == 0 && P.err) atomicExch(P.err, 19); __trap(); }
  const int my_tiles = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp >= 4 && warp < 12) {
    // ------------------------------- A producers -------------------------------
    // 256 producer threads tile a [128 rows x 64 fp32] k-block as 16 float4 columns x 16 row groups: lane -> float4
    // column (+16 lanes -> next row), so a warp reads two contiguous 256-byte row segments per load instead of 32
    // scattered lines with a thread-per-row mapping (the L1TEX pipe was the limiter: ncu "Mem Busy" 75%).
    // Each thread owns column c4 of rows rg, rg+16, ..., rg+112; row offsets are computed once per tile.
    const int pt = (warp - 4) * 32 + lane;         // 0..255
    const int c4 = pt & 15, rg = pt >> 4;
    uint32_t stage = 0, phase = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TM;
      int64_t roff[8];
      uint32_t live = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + rg + 16 * i;
        const bool ok = m < g.M;
        live |= (ok ? 1u : 0u) << i;
        roff[i] = ok ? (m / g.a_lr) * g.a_ss + (m % g.a_lr) * g.a_rs + c4 * 4 : 0;
      }
      auto load8 = [&](int kb, float4 (&v)[8]) {
        const int k0 = kb * KB;
        const float* base = g.A + (int64_t)(k0 / g.a_ct) * g.a_ts + (k0 % g.a_ct);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[i] = (live >> i) & 1u ? __ldg(reinterpret_cast<const float4*>(base + roff[i])) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      auto convert_store = [&](const float4 (&v)[8]) {
        bar_wait(&S.empty[stage], phase ^ 1, P.err, 11);
        uint8_t* st = ring + (size_t)stage * P.stage_bytes;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rg + 16 * i;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i].x, v[i].y), h1 = __floats2bfloat162_rn(v[i].z, v[i].w);
          const uint32_t o = (uint32_t)r * 128u + (uint32_t)((((c4 >> 1) ^ (r & 7)) * 16) + (c4 & 1) * 8);
          *reinterpret_cast<uint2*>(st + o) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
          if (P.x3) {
            float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            __nv_bfloat162 l0 = __floats2bfloat162_rn(v[i].x - f0.x, v[i].y - f0.y), l1 = __floats2bfloat162_rn(v[i].z - f1.x, v[i].w - f1.y);
            *reinterpret_cast<uint2*>(st + off_alo + o) = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) bar_arrive(&S.full_a[stage]);
        if (++stage == (uint32_t)P.n_stage) { stage = 0; phase ^= 1; }
      };
      // loads run one k-block ahead of the convert/store (register double buffer)
      float4 va[8], vb[8];
      load8(0, va);
      for (int kb = 0; kb < P.n_kb; kb += 2) {
        if (kb + 1 < P.n_kb) load8(kb + 1, vb);
        convert_store(va);
        if (kb + 1 < P.n_kb) {
          if (kb + 2 < P.n_kb) load8(kb + 2, va);
          convert_store(vb);
        }
      }
    }
  } else if (warp == 12) {
    // ------------------------------- W producer -------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int kb = 0; kb < P.n_kb; ++kb) {
          bar_wait(&S.empty[stage], phase ^ 1, P.err, 12);
          uint8_t* st = ring + (size_t)stage * P.stage_bytes;
          const uint8_t* src = P.wimg + (size_t)kb * 2 * w_tile;          // image: kb-major, hi then lo
          bar_expect(&S.full_w[stage], P.x3 ? 2 * w_tile : w_tile);
          bulk_load(st + off_whi, src, w_tile, &S.full_w[stage]);
          if (P.x3) bulk_load(st + off_wlo, src + w_tile, w_tile, &S.full_w[stage]);
          if (++stage == (uint32_t)P.n_stage) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 13) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int acc = t & 1;
        bar_wait(&S.d_empty[acc], ((t >> 1) & 1) ^ 1, P.err, 13);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = (uint32_t)acc * 256u;
        uint32_t accum = 0;
        for (int kb = 0; kb < P.n_kb; ++kb) {
          bar_wait(&S.full_a[stage], phase, P.err, 14);
          bar_wait(&S.full_w[stage], phase, P.err, 15);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = s32(ring + (size_t)stage * P.stage_bytes);
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks) {
            const uint64_t a_hi = kmajor_desc(base + ks * 32), w_hi = kmajor_desc(base + off_whi + ks * 32);
            mma_ss(d, a_hi, w_hi, idesc, accum);
            accum = 1;
            if (P.x3) {
              mma_ss(d, kmajor_desc(base + off_alo + ks * 32), w_hi, idesc, 1);
              mma_ss(d, a_hi, kmajor_desc(base + off_wlo + ks * 32), idesc, 1);
            }
          }
          commit(&S.empty[stage]);
          if (++stage == (uint32_t)P.n_stage) { stage = 0; phase ^= 1; }
        }
        commit(&S.d_full[acc]);
      }
    }
  } else {
    // ------------------------------- epilogue warps 0..3 -------------------------------
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int t = 0; t < my_tiles; ++t) {
      const int acc = t & 1;
      const int64_t m = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TM + r;
      const bool live = m < g.M;
      float* crow = g.C + (live ? (m / g.c_lr) * g.c_ss + (m % g.c_lr) * g.c_rs : 0);
      const float* mrow = (g.mask && live) ? g.mask + (m / g.m_lr) * g.m_ss + (m % g.m_lr) * g.m_rs : nullptr;
      bar_wait(&S.d_full[acc], (t >> 1) & 1, P.err, 16);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < N / 32; ++c) {
        uint32_t rr[32];
        ld32(lane_addr + (uint32_t)acc * 256u + (uint32_t)c * 32u, rr);
        if (live) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o;
            float* po = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v = __uint_as_float(rr[i + e]);
              if (g.bias) v += __ldg(g.bias + c * 32 + i + e);
              if (g.act == ACT_RELU) v = fmaxf(v, 0.f);
              else if (g.act == ACT_SILU) v = v / (1.f + expf(-v));
              if (mrow && !(__ldg(mrow + c * 32 + i + e) > 0.f)) v = 0.f;      // ReLU backward: zero where mask <= 0
              po[e] = v;
            }
            *reinterpret_cast<float4*>(crow + c * 32 + i) = o;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(&S.d_empty[acc]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 13) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// W [N,K] fp32 -> image [K/64][hi,lo][N rows][128 B], 16-byte chunk j of row n at chunk (j ^ (n & 7))
__global__ void gemm_tc_pack_kernel(uint8_t* __restrict__ img, const float* __restrict__ W, int N, int K) {
  const int n_kb = K / KB;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_tile = (int64_t)N * 8;
  if (idx >= (int64_t)n_kb * 2 * per_tile) return;
  const int tl = (int)(idx / per_tile), rem = (int)(idx % per_tile);
  const int kb = tl >> 1, part = tl & 1, n = rem / 8, j = rem % 8;
  __nv_bfloat16 out[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float w = W[(int64_t)n * K + kb * KB + j * 8 + e];
    __nv_bfloat16 hi = __float2bfloat16_rn(w);
    out[e] = part == 0 ? hi : __float2bfloat16_rn(w - __bfloat162float(hi));
  }
  uint8_t* dst = img + (size_t)tl * N * 128 + (size_t)n * 128 + ((j ^ (n & 7)) * 16);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out);
}

}  // namespace

size_t gemm_tc_image_bytes(int N, int K) { return (size_t)(K / KB) * 2 * N * 128; }

bool gemm_tc_eligible(const GemmArgs& g) {
  return (g.N == 128 || g.N == 256) && g.K % KB == 0 && g.a_ct % KB == 0 && g.a_ss % 4 == 0 && g.a_rs % 4 == 0 &&
         g.a_ts % 4 == 0 && g.c_ss % 4 == 0 && g.c_rs % 4 == 0 && ((uintptr_t)g.A) % 16 == 0 && ((uintptr_t)g.C) % 16 == 0 &&
         g.add == nullptr && g.M > 0;
}

int gemm_tc_pack(const float* W, int N, int K, void* image, cudaStream_t s) {
  DGDM_CHECK_ARG(W && image && K % KB == 0 && N % 8 == 0, "gemm_tc_pack: bad arguments N=%d K=%d", N, K);
  const int64_t chunks = (int64_t)(K / KB) * 2 * N * 8;
  gemm_tc_pack_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>((uint8_t*)image, W, N, K);
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}

// err: device int for the bounded-wait error code (may be nullptr)
int gemm_tc(const GemmArgs& g, const void* wimg, int x3, int* err, cudaStream_t s) {
  DGDM_CHECK_ARG(gemm_tc_eligible(g), "gemm_tc: shape not eligible (N=%d K=%d a_ct=%d)", g.N, g.K, g.a_ct);
  static int sm_counts[64] = {0};
  int dev = 0;
  DGDM_CUDA(cudaGetDevice(&dev));
  DGDM_CHECK_ARG(dev >= 0 && dev < 64, "gemm_tc: device ordinal %d out of range", dev);
  if (sm_counts[dev] == 0) {
    int n = 0;
    DGDM_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    DGDM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET + 2048 + (int)sizeof(Bars)));
    sm_counts[dev] = n;
  }
  const int sm_count = sm_counts[dev];
  TcGemmParams P{};
  P.g = g; P.wimg = (const uint8_t*)wimg; P.err = err; P.x3 = x3;
  P.n_tiles = (int)((g.M + TM - 1) / TM);
  P.n_kb = g.K / KB;
  P.stage_bytes = (x3 ? 2 : 1) * (A_TILE + g.N * 128);
  P.n_stage = SMEM_BUDGET / P.stage_bytes;
  if (P.n_stage > MAX_STAGE) P.n_stage = MAX_STAGE;
  const size_t smem = 1024 + (size_t)P.n_stage * P.stage_bytes + sizeof(Bars);
  const int grid = P.n_tiles < sm_count ? P.n_tiles : sm_count;
  gemm_tc_kernel<<<grid, NTHR, smem, s>>>(P);
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}

}  // namespace dgdm

extern "C" size_t dgdm_linear_tc_workspace_bytes(int32_t N, int32_t K) {
  if (N <= 0 || K <= 0 || K % 64 != 0) return 0;
  return dgdm::gemm_tc_image_bytes(N, K) + 256;
}

extern "C" int dgdm_linear_tc(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc,
                              int64_t M, int32_t N, int32_t K, int32_t relu, int32_t precision, void* workspace,
                              size_t workspace_bytes, void* stream) {
  using namespace dgdm;
  DGDM_CHECK_ARG(A && W && C && workspace, "dgdm_linear_tc: null pointer");
  DGDM_CHECK_ARG(precision == DGDM_PREC_BF16X3 || precision == DGDM_PREC_BF16, "dgdm_linear_tc: precision must be a tensor-core mode");
  DGDM_CHECK_ARG(K % 64 == 0 && (N == 128 || N == 256), "dgdm_linear_tc: needs K %% 64 == 0 and N in {128,256}");
  DGDM_CHECK_ARG(workspace_bytes >= gemm_tc_image_bytes(N, K) + 256, "dgdm_linear_tc: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int* err = (int*)workspace;
  uint8_t* img = (uint8_t*)workspace + 256;
  DGDM_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s));
  DGDM_TRY(gemm_tc_pack(W, N, K, img, s));
  GemmArgs g = gemm_plain(A, lda, W, bias, C, ldc, M, N, K, relu ? ACT_RELU : ACT_NONE);
  return gemm_tc(g, img, precision == DGDM_PREC_BF16X3, err, s);
}
