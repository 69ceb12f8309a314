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
inline int64_t act_rows(int64_t n, int L) { return ((2 + n * (L + 2)) + 127) / 128 * 128 + 8; }

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
  int out_mode;                 // 0: fp32 quad-major compact [C/4][n*Lo][4]; 1: bf16 hi/lo chunk-major
  float* o_f32; int64_t o_rows;
  uint8_t* o_hi; uint8_t* o_lo; int64_t o_plane; int o_chunk0;
  int Lo, o_step, o_off;
  int n_w, n_a;                 // ring depths
  int* err;
};

// standard conv: `taps` taps over `cin` channels starting at chunk `in_chunk0`, first tap at row offset `roff0`
void conv_tc_blocks(ConvTcParams& P, int in_chunk0, int cin, int taps, int roff0);
int conv_tc_launch(ConvTcParams& P, cudaStream_t s);

}  // namespace dgdm
