// K3 convolutions on tcgen05 with bulk-copied operands on both sides (no SIMT producer).
//
// Every Conv1d / strided Conv1d / ConvTranspose1d of ConditionalUnet1D (generator/diffusion_utils.py:42,51,66,96)
// is an implicit GEMM  out[p, co] = bias[co] + sum_{tap, ci} W[co][tap][ci] * act[p + roff(tap), ci]  over the
// physical rows p of a chunk-major bf16 activation buffer (conv_tc.cuh).  One persistent CTA per SM, 128 rows per
// tile, all N (128 or 256) output channels per tile:
//   warp 8   producer: per 64-channel block, eight (x3: sixteen) 2112-byte cp.async.bulk copies stage the
//            132-row activation slab ONCE for all taps; per (block, tap) one (x3: two) copies stage the
//            pre-swizzled weight tile.  mbarrier transaction counts signal completion.
//   warp 9   MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 x N x K=16.  A descriptor = K-major, no swizzle,
//            start address advanced by 16 bytes per tap row; B descriptor = K-major SWIZZLE_128B.  fp32 accumulators
//            in TMEM, double-buffered (2 x 256 columns).  fp32-grade mode issues hi.hi, lo.hi, hi.lo.
//   warps 0-7 epilogue: tcgen05.ld, + bias, then either fp32 "quad-major" rows for the GroupNorm kernel or bf16
//            hi/lo chunk-major rows for the next conv (Downsample1d / Upsample1d outputs, which have no norm).
#include <cuda_bf16.h>

#include "conv_tc.cuh"

namespace dgdm {
namespace {

constexpr int TM = 128, SLAB_ROWS = TM + 4, PLANE_B = SLAB_ROWS * 16, SLAB_B = 8 * PLANE_B, A_STAGE_B = 2 * SLAB_B;
constexpr int NEPI = 8, NTHR = (NEPI + 2) * 32;
constexpr int MAX_W = 6, MAX_A = 3;

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
// commit / mma_ss are called by every lane of the issuing warp (uniform control flow, so every operand stays in
// uniform registers: no ELECT / R2UR / BRA.U.ANY waterfall in front of each UTCHMMA); lane 0 executes the instruction
__device__ __forceinline__ void commit(uint64_t* b) {
  if ((threadIdx.x & 31) == 0)
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  if ((threadIdx.x & 31) == 0)
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accum) : "memory");
}
// B: K-major, SWIZZLE_128B, SBO = 1024 B, descriptor version 1
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// A: K-major, no swizzle: 8-row x 16-byte core matrices; LBO = stride between the two 16-byte K chunks of one
// K=16 step (one slab plane), SBO = stride between 8-row groups (128 B: rows are contiguous in a plane)
__device__ __forceinline__ uint64_t nosw_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(PLANE_B >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
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
  uint64_t a_full[MAX_A], a_empty[MAX_A], w_full[MAX_W], w_empty[MAX_W], d_full[2], d_empty[2];
  uint32_t tmem_base, pad_;
};

__global__ void __launch_bounds__(NTHR, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams P) {
  extern __shared__ uint8_t raw[];
  const uint32_t w_tile = (uint32_t)P.N * 128u;
  uint8_t* wring = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);       // 1024-byte aligned, shared address space kept
  uint8_t* aring = wring + (size_t)P.n_w * w_tile;
  Bars& S = *reinterpret_cast<Bars*>(aring + (size_t)P.n_a * A_STAGE_B);
  // warp index through a shuffle: provably warp-uniform for the compiler (see the MMA issuer)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < P.n_a; ++s) { bar_init(&S.a_full[s], 1); bar_init(&S.a_empty[s], 1); }
    for (int s = 0; s < P.n_w; ++s) { bar_init(&S.w_full[s], 1); bar_init(&S.w_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { bar_init(&S.d_full[i], 1); bar_init(&S.d_empty[i], NEPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NEPI + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // all 512 columns are allocated, so the base is 0 by construction; a literal keeps the MMA operands uniform
  if (S.tmem_base != 0) { if (tid == 0 && P.err) atomicExch(P.err, 20); __trap(); }
  constexpr uint32_t tmem = 0;
  const int my_tiles = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == NEPI) {
    // ------------------------------- producer -------------------------------
    if (lane == 0) {
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int64_t f0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TM;
        for (int c = 0; c < P.n_cb; ++c) {
          const ConvTcBlock& cb = P.cb[c];
          bar_wait(&S.a_empty[sa], pa ^ 1, P.err, 21);
          uint8_t* slab = aring + (size_t)sa * A_STAGE_B;
          bar_expect(&S.a_full[sa], P.x3 ? 2 * SLAB_B : SLAB_B);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            bulk_load(slab + j * PLANE_B, P.a_hi + (int64_t)(cb.chunk0 + j) * P.a_plane + f0 * 16, PLANE_B, &S.a_full[sa]);
          if (P.x3) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              bulk_load(slab + SLAB_B + j * PLANE_B, P.a_lo + (int64_t)(cb.chunk0 + j) * P.a_plane + f0 * 16, PLANE_B, &S.a_full[sa]);
          }
          if (++sa == (uint32_t)P.n_a) { sa = 0; pa ^= 1; }
          for (int tp = 0; tp < cb.ntap; ++tp) {
            const uint8_t* src = P.wimg + (size_t)cb.tap[tp].wkb * 2 * w_tile;
            for (int part = 0; part < (P.x3 ? 2 : 1); ++part) {
              bar_wait(&S.w_empty[sw], pw ^ 1, P.err, 22);
              bar_expect(&S.w_full[sw], w_tile);
              bulk_load(wring + (size_t)sw * w_tile, src + (size_t)part * w_tile, w_tile, &S.w_full[sw]);
              if (++sw == (uint32_t)P.n_w) { sw = 0; pw ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == NEPI + 1) {
    // ------------------------------- MMA issuer -------------------------------
    {   // whole warp, uniform control flow; mma_ss / commit issue from lane 0
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int acc = t & 1;
        bar_wait(&S.d_empty[acc], ((t >> 1) & 1) ^ 1, P.err, 23);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + (uint32_t)acc * 256u;
        uint32_t accum = 0;
        for (int c = 0; c < P.n_cb; ++c) {
          const ConvTcBlock& cb = P.cb[c];
          bar_wait(&S.a_full[sa], pa, P.err, 24);
          const uint32_t abase = s32(aring + (size_t)sa * A_STAGE_B);
          for (int tp = 0; tp < cb.ntap; ++tp) {
            const uint32_t a0 = abase + (uint32_t)cb.tap[tp].roff * 16u;
            const uint64_t ad_hi = nosw_desc(a0), ad_lo = nosw_desc(a0 + SLAB_B);   // + ks * 2 planes: low word only, no carry
            // W hi: hi.hi and lo.hi
            bar_wait(&S.w_full[sw], pw, P.err, 25);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t wb = s32(wring + (size_t)sw * w_tile);
            // (one add per descriptor and K step: + 32 bytes = + 2 in the 14-bit address field; the issuing thread sets the
            // pace of N = 128 MMAs, scripts/microbench/mma_2cta.cu)
            const uint64_t wd0 = sw128_desc(wb);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t wd = wd0 + (uint64_t)(2 * ks);
              mma_ss(d, ad_hi + (uint64_t)(ks * 2 * PLANE_B >> 4), wd, idesc, accum);
              accum = 1;
              if (P.x3) mma_ss(d, ad_lo + (uint64_t)(ks * 2 * PLANE_B >> 4), wd, idesc, 1);
            }
            commit(&S.w_empty[sw]);
            if (++sw == (uint32_t)P.n_w) { sw = 0; pw ^= 1; }
            if (P.x3) {   // W lo: hi.lo
              bar_wait(&S.w_full[sw], pw, P.err, 26);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              wb = s32(wring + (size_t)sw * w_tile);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) mma_ss(d, ad_hi + (uint64_t)(ks * 2 * PLANE_B >> 4), sw128_desc(wb) + (uint64_t)(2 * ks), idesc, 1);
              commit(&S.w_empty[sw]);
              if (++sw == (uint32_t)P.n_w) { sw = 0; pw ^= 1; }
            }
          }
          commit(&S.a_empty[sa]);
          if (++sa == (uint32_t)P.n_a) { sa = 0; pa ^= 1; }
        }
        commit(&S.d_full[acc]);
      }
    }
  } else {
    // ------------------------------- epilogue warps 0..7 -------------------------------
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int Lp = P.Ld + 2;
    const int ncol = P.N / 2, c_lo = half * ncol;
    for (int t = 0; t < my_tiles; ++t) {
      const int acc = t & 1;
      const int64_t f = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TM + r;    // output physical row = f + 2
      const int64_t qf = f - 2;
      const int64_t b = qf >= 0 ? qf / Lp : 0;
      const int l = (int)(qf - b * Lp);
      const bool live = qf >= 0 && l < P.Ld && b < P.n;
      const int64_t orow = b * P.Lo + (int64_t)l * P.o_step + P.o_off;               // compact row (mode 0)
      const int64_t prow = 4 + b * (P.Lo + 2) + (int64_t)l * P.o_step + P.o_off;       // physical row (mode 1)
      bar_wait(&S.d_full[acc], (t >> 1) & 1, P.err, 27);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = c_lo; c < c_lo + ncol; c += 32) {
        uint32_t rr[32];
        ld32(lane_addr + (uint32_t)acc * 256u + (uint32_t)c, rr);
        if (!live) continue;
        if (P.out_mode == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = P.bias ? __ldg(reinterpret_cast<const float4*>(P.bias + c + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            o.x = __uint_as_float(rr[i]) + bv.x; o.y = __uint_as_float(rr[i + 1]) + bv.y;
            o.z = __uint_as_float(rr[i + 2]) + bv.z; o.w = __uint_as_float(rr[i + 3]) + bv.w;
            *reinterpret_cast<float4*>(P.o_f32 + ((int64_t)((c + i) >> 2) * P.o_rows + orow) * 4) = o;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 bv = P.bias ? __ldg(reinterpret_cast<const float2*>(P.bias + c + i + 2 * e)) : make_float2(0.f, 0.f);
              const float v0 = __uint_as_float(rr[i + 2 * e]) + bv.x, v1 = __uint_as_float(rr[i + 2 * e + 1]) + bv.y;
              __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
              hi[e] = *reinterpret_cast<uint32_t*>(&h);
              if (P.x3) {                         // (single-pass mode: the next conv reads only the hi plane)
                const float h0 = __uint_as_float(hi[e] << 16), h1 = __uint_as_float(hi[e] & 0xFFFF0000u);
                __nv_bfloat162 lw = __floats2bfloat162_rn(v0 - h0, v1 - h1);
                lo[e] = *reinterpret_cast<uint32_t*>(&lw);
              }
            }
            const int64_t off = (int64_t)(P.o_chunk0 + ((c + i) >> 3)) * P.o_plane + prow * 16;
            *reinterpret_cast<uint4*>(P.o_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (P.x3) *reinterpret_cast<uint4*>(P.o_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
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
  if (warp == NEPI + 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace

void conv_tc_blocks(ConvTcParams& P, int in_chunk0, int cin, int taps, int roff0) {
  P.n_cb = cin / 64;
  for (int c = 0; c < P.n_cb; ++c) {
    P.cb[c].chunk0 = (uint16_t)(in_chunk0 + c * 8);
    P.cb[c].ntap = (uint16_t)taps;
    for (int t = 0; t < taps; ++t) P.cb[c].tap[t] = ConvTcTap{(uint16_t)(roff0 + t), (uint16_t)(t * (cin / 64) + c)};
  }
}

int conv_tc_launch(ConvTcParams& P, cudaStream_t s) {
  DGDM_CHECK_ARG((P.N == 128 || P.N == 256) && P.n_cb >= 1 && P.n_cb <= 8, "conv_tc: N=%d n_cb=%d unsupported", P.N, P.n_cb);
  static int sm_counts[64] = {0};
  int dev = 0;
  DGDM_CUDA(cudaGetDevice(&dev));
  DGDM_CHECK_ARG(dev >= 0 && dev < 64, "conv_tc: device ordinal %d out of range", dev);
  const int smem_256 = 1024 + 4 * 256 * 128 + 2 * A_STAGE_B + (int)sizeof(Bars);
  const int smem_128 = 1024 + 6 * 128 * 128 + 3 * A_STAGE_B + (int)sizeof(Bars);
  const int max_smem = smem_256 > smem_128 ? smem_256 : smem_128;
  if (sm_counts[dev] == 0) {
    int n = 0;
    DGDM_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    DGDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    sm_counts[dev] = n;
  }
  const int sm_count = sm_counts[dev];
  P.n_w = P.N == 256 ? 4 : 6;
  P.n_a = P.N == 256 ? 2 : 3;
  P.n_tiles = (int)((2 + P.n * (P.Ld + 2) + TM - 1) / TM);
  const size_t smem = 1024 + (size_t)P.n_w * P.N * 128 + (size_t)P.n_a * A_STAGE_B + sizeof(Bars);
  DGDM_CHECK_ARG(smem <= (size_t)max_smem, "conv_tc: shared memory plan %zu exceeds %d", smem, max_smem);
  const int grid = P.n_tiles < sm_count ? P.n_tiles : sm_count;
  conv_tc_kernel<<<grid, NTHR, smem, s>>>(P);
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}

}  // namespace dgdm
