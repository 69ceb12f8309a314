// Interface of the tcgen05 conv kernel over "chunk-major" bf16 activations (conv_tc.cu), used by K3 (unet1d.cu).
#pragma once
#include "common.cuh"

namespace dgdm {

// A denoiser activation tensor in HBM, split into bf16 hi and lo parts (hi + lo ~ fp32 to 2^-17):
//   element (sample b, position l, channel c) lives at   part + (chunk0 + c/8) * plane + p * 16 + (c % 8) * 2
//   with physical row p = 4 + b * (L + 2) + l.
// Rows 0..3, the two rows between consecutive samples and everything after the last sample are zero, so a k=5
// "same" convolution of the whole batch is ONE sequence convolution over physical rows: output row p reads rows
// p-2..p+2, for every p, with no per-sample addressing (the outputs computed for the in-between rows are dropped).
// A 64-channel block of 132 consecutive rows is eight contiguous 2112-byte runs -- eight cp.async.bulk copies --
// and lands in shared memory exactly in the tcgen05 K-major no-swizzle canonical layout (8-row x 16-byte core
// matrices, SBO 128 B, LBO 2112 B), where a tap shift is +16 bytes on the descriptor start address: one staged
// slab feeds all five taps.
struct ActBuf {
  uint8_t* hi; uint8_t* lo;
  int64_t plane;      // bytes per chunk plane = rows * 16
};
// (+136: the sample-aligned tiles of the fused GroupNorm mode start at row 2 + b0 * (L + 2) and stage 132 rows)
inline int64_t act_rows(int64_t n, int L) { return ((2 + n * (L + 2)) + 127) / 128 * 128 + 136; }

struct ConvTcTap { uint16_t roff, wkb; };                 // A row offset (tap shift), weight k-block index in the image
struct ConvTcBlock { uint16_t chunk0, ntap; ConvTcTap tap[5]; };   // one 64-channel block of the A operand

struct ConvTcParams {
  const uint8_t* a_hi; const uint8_t* a_lo; int64_t a_plane;
  const uint8_t* wimg;          // gemm_tc_pack image: per k-block [hi N x 128 B][lo N x 128 B], SWIZZLE_128B K-major
  const float* bias;
  int N, x3, n_cb;
  ConvTcBlock cb[8];
  int64_t n;                    // samples
  int Ld;                       // positions per sample in the compute domain (period Ld + 2)
  int n_tiles;
  // output: row l of sample b goes to output position l * o_step + o_off of a sample with Lo positions
  int out_mode;                 // 0: fp32 quad-major compact [C/4][n*Lo][4]; 1: bf16 hi/lo chunk-major;
                                // 2: GroupNorm(8) + Mish (+ FiLM) (+ residual) in the epilogue -> bf16 hi/lo chunk-major;
                                // 3: the same, then the 1x1 output conv (Conv1d(N,1,1)) -> eps
  float* o_f32; int64_t o_rows;
  uint8_t* o_hi; uint8_t* o_lo; int64_t o_plane; int o_chunk0;
  int Lo, o_step, o_off;
  int n_w, n_a;                 // ring depths
  int* err;
  // ---- out_mode 2 / 3 (Conv1dBlock, diffusion_utils.py:80-97, and the tail of ConditionalResidualBlock1D, :100-120).
  // A tile holds gn_S = 128 / (Ld + 2) WHOLE samples (tile_rows = gn_S * (Ld + 2) of the 128 MMA rows are used), so the
  // statistics of a (sample, group) never leave the CTA.  o_step = 1, o_off = 0, Lo = Ld.
  int gn_S, tile_rows, f_base;  // set by conv_tc_launch (modes 0 / 1: tile_rows = 128, f_base = 0)
  const float* gamma; const float* beta; const float* film;   // film: [2N] (scale | shift) or null
  int res_mode;                 // 0 none, 1 chunk-major bf16, 2 fp32 quad-major [N/4][r_rows][4], 3 Cin=1 1x1 conv of x
  const uint8_t* r_hi; const uint8_t* r_lo; int64_t r_plane; int r_chunk0;
  const float* r_f32; int64_t r_rows;
  const float* r_x; const float* r_w; const float* r_b;
  int o_split;                  // 1: even / odd positions go to two channel halves of a half-length buffer (for the strided conv)
  const float* p_w; const float* p_b; float* eps;             // mode 3
};

#ifdef __CUDACC__
// mish for the tensor-core path: x * n / (n + 2), n = e^x (e^x + 2), on the raw MUFU approximations (ex2 / rcp, 2^-21-grade:
// far inside that path's tolerance; n + 2 >= 2 is never denormal, so none of the range fix-ups of __expf / __fdividef are
// needed -- they were 6 of the 15 instructions).  The clamp makes x >= 20 return x exactly (n / (n + 2) rounds to 1) and keeps e finite.
__device__ __forceinline__ float mish_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 20.f) * 1.4426950408889634f));
  const float n = fmaf(e, e, e + e);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n + 2.f));
  return x * n * r;
}
#endif

// standard conv: `taps` taps over `cin` channels starting at chunk `in_chunk0`, first tap at row offset `roff0`
void conv_tc_blocks(ConvTcParams& P, int in_chunk0, int cin, int taps, int roff0);
int conv_tc_launch(ConvTcParams& P, cudaStream_t s);

}  // namespace dgdm
