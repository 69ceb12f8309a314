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
//   warps 0-15 epilogue (warp w: TMEM lanes 32 (w % 4).., column quarter w / 4): tcgen05.ld, + bias, then
//            mode 0  fp32 "quad-major" rows (the 1x1 residual convs, and the unfused GroupNorm path),
//            mode 1  bf16 hi/lo chunk-major rows for the next conv (Downsample1d / Upsample1d outputs: no norm),
//            mode 2/3  GroupNorm(8) + Mish (+ FiLM) (+ residual) on the accumulator itself: tiles are aligned to whole
//                    samples, pass 1 reads the accumulator for per-(row, group) mean / M2, one thread per (sample, group)
//                    combines them in a fixed order, pass 2 reads the accumulator again, normalises and writes bf16
//                    chunk-major rows for the next conv (or the fused 1x1 output conv).  The conv output never goes to
//                    HBM in fp32 and there is no separate GroupNorm launch.
#include <cuda_bf16.h>

#include "conv_tc.cuh"

namespace dgdm {
namespace {

constexpr int TM = 128, SLAB_ROWS = TM + 4, PLANE_B = SLAB_ROWS * 16, SLAB_B = 8 * PLANE_B, A_STAGE_B = 2 * SLAB_B;
constexpr int NEPI = 16, NTHR = (NEPI + 2) * 32;
constexpr int NG = 1, GW_WARPS = NEPI / NG, NPART = GW_WARPS / 4;   // epilogue groups (see the epilogue), warps per group, column parts per group
constexpr int GN_MAX_S = 44;
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

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

struct Bars {
  uint64_t a_full[MAX_A], a_empty[MAX_A], w_full[MAX_W], w_empty[MAX_W], d_full[2], d_empty[2];
  uint32_t tmem_base, pad_;
};
// per-channel parameters and the statistics exchange of the fused GroupNorm epilogue ([NG]: one per epilogue group)
struct GnShared {
  float bias[256], gamma[256], beta[256];
  float x0[256], x1[256];           // FiLM scale | shift (conv0), or the Cin=1 residual conv's weight | bias (res_mode 3), or the 1x1 output conv's weight (mode 3) | -
  float2 part[NG][8][TM + 1];       // (mean, M2) of one row's channels of one group
  float2 stat[NG][GN_MAX_S][8];     // (mean, rstd) of one (sample, group)
  float proj[NG][NPART][TM];
};
constexpr int NGRP = GW_WARPS * 32;  // threads of one epilogue group
__device__ __forceinline__ void epi_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(NGRP) : "memory"); }

// GroupNorm(8) + Mish (+ FiLM) (+ residual) epilogue of one tile; GW = channels per group (16: N = 128, 32: N = 256)
template <int GW>
__device__ __forceinline__ void epilogue_gn(const ConvTcParams& P, GnShared& G, uint32_t taddr, int c_lo, int ncol, int r,
                                            int part, int etid, int grp, int64_t tile) {
  const int Lp = P.Ld + 2;
  const int bs = r / Lp, l = r - bs * Lp;
  const int64_t b = tile * P.gn_S + bs;
  const bool live = bs < P.gn_S && l < P.Ld && b < P.n;
  // ---- pass 1: per-row, per-group mean and centred sum of squares
#pragma unroll 1
  for (int c = c_lo; c < c_lo + ncol; c += 32) {
    uint32_t rr[32];
    ld32(taddr + (uint32_t)c, rr);
    if (live) {
#pragma unroll
    for (int h = 0; h < 32 / GW; ++h) {
      float v[GW], s = 0.f;
#pragma unroll
      for (int i = 0; i < GW; i += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(&G.bias[c + h * GW + i]);
        v[i] = __uint_as_float(rr[h * GW + i]) + bv.x; v[i + 1] = __uint_as_float(rr[h * GW + i + 1]) + bv.y;
        v[i + 2] = __uint_as_float(rr[h * GW + i + 2]) + bv.z; v[i + 3] = __uint_as_float(rr[h * GW + i + 3]) + bv.w;
        s += (v[i] + v[i + 1]) + (v[i + 2] + v[i + 3]);
      }
#pragma unroll
      for (int i = 0; i < GW; ++i) rr[h * GW + i] = __float_as_uint(v[i]);
      const float m = s * (1.f / GW);
      float m2 = 0.f;
#pragma unroll
      for (int i = 0; i < GW; ++i) { const float d = v[i] - m; m2 = fmaf(d, d, m2); }
      G.part[grp][(c + h * GW) / GW][r] = make_float2(m, m2);
    }
    }
    st32(taddr + (uint32_t)c, rr);      // accumulator + bias goes back to TMEM: pass 2 re-reads it without the add
  }
  epi_sync(grp);
  // ---- one thread per (sample, group): equal-count blocks, so the mean is the mean of the row means and
  //      M2 = sum_i M2_i + GW (m_i - mean)^2, summed in a fixed order (deterministic)
  for (int e = etid; e < P.gn_S * 8; e += NGRP) {
    const int sb = e >> 3, g = e & 7;
    if (tile * P.gn_S + sb < P.n) {
      const float2* pr = &G.part[grp][g][sb * Lp];
      float s0 = 0.f, s1 = 0.f;
      int l2 = 0;
#pragma unroll 4
      for (; l2 + 1 < P.Ld; l2 += 2) { s0 += pr[l2].x; s1 += pr[l2 + 1].x; }
      if (l2 < P.Ld) s0 += pr[l2].x;
      const float mean = (s0 + s1) / (float)P.Ld;
      float a0 = 0.f, a1 = 0.f;
      l2 = 0;
#pragma unroll 4
      for (; l2 + 1 < P.Ld; l2 += 2) {
        const float2 p0 = pr[l2], p1 = pr[l2 + 1];
        const float d0 = p0.x - mean, d1 = p1.x - mean;
        a0 += fmaf((float)GW * d0, d0, p0.y); a1 += fmaf((float)GW * d1, d1, p1.y);
      }
      if (l2 < P.Ld) { const float2 p0 = pr[l2]; const float d0 = p0.x - mean; a0 += fmaf((float)GW * d0, d0, p0.y); }
      G.stat[grp][sb][g] = make_float2(mean, rsqrtf((a0 + a1) / (float)(P.Ld * GW) + 1e-5f));
    }
  }
  epi_sync(grp);
  // ---- pass 2: normalise, Mish, FiLM, residual, write
  const int64_t crow = b * P.Ld + l;                                                    // compact row
  const int64_t prow = 4 + b * Lp + l;                                                  // physical row (input / residual)
  const int64_t orow = P.o_split ? 4 + b * (P.Ld / 2 + 2) + (l >> 1) : prow;            // physical output row
  const int o_chunk = P.o_chunk0 + (P.o_split ? (l & 1) * (P.N / 8) : 0);
  const float xin = (live && P.res_mode == 3) ? P.r_x[crow] : 0.f;
  const bool film = P.film != nullptr;
  float proj = 0.f;
#pragma unroll 1
  for (int c = c_lo; c < c_lo + ncol; c += 32) {
    uint32_t rr[32];
    ld32(taddr + (uint32_t)c, rr);
    if (!live) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = c + 8 * j, chunk = ch >> 3;
      const float2 st = G.stat[grp][bs][ch / GW];
      float y[8];
#pragma unroll
      for (int e4 = 0; e4 < 8; e4 += 4) {
        const float4 gm = *reinterpret_cast<const float4*>(&G.gamma[ch + e4]);
        const float4 bt = *reinterpret_cast<const float4*>(&G.beta[ch + e4]);
        const float4 fs = *reinterpret_cast<const float4*>(&G.x0[ch + e4]);
        const float4 ft = *reinterpret_cast<const float4*>(&G.x1[ch + e4]);
        const float gg[4] = {gm.x, gm.y, gm.z, gm.w}, tt[4] = {bt.x, bt.y, bt.z, bt.w};
        const float f0[4] = {fs.x, fs.y, fs.z, fs.w}, f1[4] = {ft.x, ft.y, ft.z, ft.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = __uint_as_float(rr[8 * j + e4 + e]);
          const float sc = st.y * gg[e];
          const float mv = mish_fast(fmaf(v, sc, fmaf(-st.x, sc, tt[e])));
          y[e4 + e] = film ? fmaf(f0[e], mv, f1[e]) : mv;
        }
      }
      if (P.res_mode == 1) {
        const int64_t off = (int64_t)(P.r_chunk0 + chunk) * P.r_plane + prow * 16;
        const uint4 h = *reinterpret_cast<const uint4*>(P.r_hi + off);
        const uint4 lw = P.x3 ? *reinterpret_cast<const uint4*>(P.r_lo + off) : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lv[4] = {lw.x, lw.y, lw.z, lw.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          y[2 * e] += __uint_as_float(hw[e] << 16) + __uint_as_float(lv[e] << 16);
          y[2 * e + 1] += __uint_as_float(hw[e] & 0xFFFF0000u) + __uint_as_float(lv[e] & 0xFFFF0000u);
        }
      } else if (P.res_mode == 2) {
        const float4 r0 = *reinterpret_cast<const float4*>(P.r_f32 + ((int64_t)(2 * chunk) * P.r_rows + crow) * 4);
        const float4 r1 = *reinterpret_cast<const float4*>(P.r_f32 + ((int64_t)(2 * chunk + 1) * P.r_rows + crow) * 4);
        y[0] += r0.x; y[1] += r0.y; y[2] += r0.z; y[3] += r0.w; y[4] += r1.x; y[5] += r1.y; y[6] += r1.z; y[7] += r1.w;
      } else if (P.res_mode == 3) {
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] += fmaf(G.x0[ch + e], xin, G.x1[ch + e]);
      }
      if (P.out_mode == 3) {
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) d = fmaf(y[e], G.x0[ch + e], d);
        proj += d;
      } else {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * e], y[2 * e + 1]);
          hi[e] = *reinterpret_cast<uint32_t*>(&h);
          if (P.x3) {
            const float h0 = __uint_as_float(hi[e] << 16), h1 = __uint_as_float(hi[e] & 0xFFFF0000u);
            __nv_bfloat162 lw = __floats2bfloat162_rn(y[2 * e] - h0, y[2 * e + 1] - h1);
            lo[e] = *reinterpret_cast<uint32_t*>(&lw);
          }
        }
        const int64_t off = (int64_t)(o_chunk + chunk) * P.o_plane + orow * 16;
        *reinterpret_cast<uint4*>(P.o_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (P.x3) *reinterpret_cast<uint4*>(P.o_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  if (P.out_mode == 3) {      // Conv1d(N,1,1): the column parts of a row are summed in a fixed order
    G.proj[grp][part][r] = proj;
    epi_sync(grp);
    if (part == 0 && live) {
      float d = G.proj[grp][0][r];
#pragma unroll
      for (int k = 1; k < NPART; ++k) d += G.proj[grp][k][r];
      P.eps[crow] = P.p_b[0] + d;
    }
  } else if (bs < P.gn_S && l >= P.Ld && b < P.n) {
    // the two rows after a sample: this kernel keeps them zero for the next conv (no separate memset of the buffer)
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int chunk = c_lo >> 3; chunk < (c_lo + ncol) >> 3; ++chunk) {
      if (!P.o_split) {
        const int64_t off = (int64_t)(P.o_chunk0 + chunk) * P.o_plane + prow * 16;
        *reinterpret_cast<uint4*>(P.o_hi + off) = z;
        if (P.x3) *reinterpret_cast<uint4*>(P.o_lo + off) = z;
      } else {   // row Ld -> the even half's two pad rows, row Ld + 1 -> the odd half's
        const int64_t off = (int64_t)(P.o_chunk0 + (l - P.Ld) * (P.N / 8) + chunk) * P.o_plane + (4 + b * (P.Ld / 2 + 2) + P.Ld / 2) * 16;
        *reinterpret_cast<uint4*>(P.o_hi + off) = z; *reinterpret_cast<uint4*>(P.o_hi + off + 16) = z;
        if (P.x3) { *reinterpret_cast<uint4*>(P.o_lo + off) = z; *reinterpret_cast<uint4*>(P.o_lo + off + 16) = z; }
      }
    }
  }
}

__global__ void __launch_bounds__(NTHR, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams P) {
  extern __shared__ uint8_t raw[];
  const uint32_t w_tile = (uint32_t)P.N * 128u;
  uint8_t* wring = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);       // 1024-byte aligned, shared address space kept
  uint8_t* aring = wring + (size_t)P.n_w * w_tile;
  Bars& S = *reinterpret_cast<Bars*>(aring + (size_t)P.n_a * A_STAGE_B);
  GnShared& G = *reinterpret_cast<GnShared*>(aring + (size_t)P.n_a * A_STAGE_B + ((sizeof(Bars) + 15) & ~(size_t)15));
  // warp index through a shuffle: provably warp-uniform for the compiler (see the MMA issuer)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < P.n_a; ++s) { bar_init(&S.a_full[s], 1); bar_init(&S.a_empty[s], 1); }
    for (int s = 0; s < P.n_w; ++s) { bar_init(&S.w_full[s], 1); bar_init(&S.w_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { bar_init(&S.d_full[i], 1); bar_init(&S.d_empty[i], GW_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (P.out_mode >= 2) {
    for (int i = tid; i < P.N; i += NTHR) {
      G.bias[i] = P.bias ? P.bias[i] : 0.f; G.gamma[i] = P.gamma[i]; G.beta[i] = P.beta[i];
      float x0 = 1.f, x1 = 0.f;
      if (P.film) { x0 = P.film[i]; x1 = P.film[P.N + i]; }
      else if (P.res_mode == 3) { x0 = P.r_w[i]; x1 = P.r_b[i]; }
      else if (P.out_mode == 3) x0 = P.p_w[i];
      G.x0[i] = x0; G.x1[i] = x1;
    }
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
        const int64_t f0 = P.f_base + ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * P.tile_rows;
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
    // ------------------------------- epilogue warps 0..15 -------------------------------
    // NG groups of NEPI / NG warps; group g takes this CTA's tiles g, g + NG, ...  (Measured: the epilogue is bound by
    // issue slots, not latency -- two groups of eight warps were 20 % slower than one of sixteen, because a group then
    // holds its accumulator twice as long and the MMAs of tile t + 2 wait for it.  NG = 1.)
    const int grp = warp / GW_WARPS, w8 = warp % GW_WARPS;
    const int q = w8 & 3, part = w8 >> 2;          // TMEM lane quarter = warp % 4
    const int r = q * 32 + lane, etid = w8 * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int Lp = P.Ld + 2;
    const int ncol = P.N / NPART, c_lo = part * ncol;
    for (int t = grp; t < my_tiles; t += NG) {
      const int acc = t & 1;
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
      bar_wait(&S.d_full[acc], (t >> 1) & 1, P.err, 27);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (P.out_mode >= 2) {
        if (P.N == 128) epilogue_gn<16>(P, G, lane_addr + (uint32_t)acc * 256u, c_lo, ncol, r, part, etid, grp, tile);
        else epilogue_gn<32>(P, G, lane_addr + (uint32_t)acc * 256u, c_lo, ncol, r, part, etid, grp, tile);
      } else {
      const int64_t f = tile * TM + r;                                                 // output physical row = f + 2
      const int64_t qf = f - 2;
      const int64_t b = qf >= 0 ? qf / Lp : 0;
      const int l = (int)(qf - b * Lp);
      const bool live = qf >= 0 && l < P.Ld && b < P.n;
      const bool pad = P.out_mode == 1 && P.o_step == 1 && qf >= 0 && l >= P.Ld && b < P.n;   // rows after a sample stay zero
      const int64_t orow = b * P.Lo + (int64_t)l * P.o_step + P.o_off;               // compact row (mode 0)
      const int64_t prow = 4 + b * (P.Lo + 2) + (int64_t)l * P.o_step + P.o_off;       // physical row (mode 1)
#pragma unroll 1
      for (int c = c_lo; c < c_lo + ncol; c += 32) {
        uint32_t rr[32];
        ld32(lane_addr + (uint32_t)acc * 256u + (uint32_t)c, rr);
        if (pad) {
          for (int i = 0; i < 32; i += 8) {
            const int64_t off = (int64_t)(P.o_chunk0 + ((c + i) >> 3)) * P.o_plane + prow * 16;
            *reinterpret_cast<uint4*>(P.o_hi + off) = make_uint4(0u, 0u, 0u, 0u);
            if (P.x3) *reinterpret_cast<uint4*>(P.o_lo + off) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
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
  const int tail = (int)(((sizeof(Bars) + 15) & ~(size_t)15) + sizeof(GnShared));
  const int smem_256 = 1024 + 4 * 256 * 128 + 2 * A_STAGE_B + tail;
  const int smem_128 = 1024 + 6 * 128 * 128 + 3 * A_STAGE_B + tail;
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
  if (P.out_mode >= 2) {
    DGDM_CHECK_ARG(P.Ld >= 1 && P.Ld + 2 <= TM && P.gamma && P.beta && P.o_step == 1 && P.o_off == 0 && P.Lo == P.Ld,
                   "conv_tc: fused GroupNorm needs whole samples per tile (Ld=%d) and a unit-stride output", P.Ld);
    P.gn_S = TM / (P.Ld + 2);
    DGDM_CHECK_ARG(P.gn_S <= GN_MAX_S, "conv_tc: %d samples per tile", P.gn_S);
    P.tile_rows = P.gn_S * (P.Ld + 2); P.f_base = 2;
    P.n_tiles = (int)((P.n + P.gn_S - 1) / P.gn_S);
  } else {
    P.gn_S = 0; P.tile_rows = TM; P.f_base = 0;
    P.n_tiles = (int)((2 + P.n * (P.Ld + 2) + TM - 1) / TM);
  }
  const size_t smem = 1024 + (size_t)P.n_w * P.N * 128 + (size_t)P.n_a * A_STAGE_B + (size_t)tail;
  DGDM_CHECK_ARG(smem <= (size_t)max_smem, "conv_tc: shared memory plan %zu exceeds %d", smem, max_smem);
  const int grid = P.n_tiles < sm_count ? P.n_tiles : sm_count;
  conv_tc_kernel<<<grid, NTHR, smem, s>>>(P);
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}

}  // namespace dgdm
