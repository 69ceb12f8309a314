// K3: ConditionalUnet1D forward (generator/diffusion_utils.py:123-285), configuration of generator/train.py:80.
//
// Activations live channels-last in zero-padded buffers [n][L+4][C] (2 zero rows either side), so a
// Conv1d(k=5, pad=2) is an implicit GEMM whose A row for (sample b, position l) is the 5*C contiguous
// floats starting at row l of the padded sample: no im2col is ever materialised.  The strided
// Downsample1d (k3 s2 p1) and the ConvTranspose1d Upsample1d (k4 s2 p1, split into even/odd output
// phases) are the same GEMM with different row strides.  GroupNorm(8)+Mish(+FiLM)(+residual) is one
// warp-per-(sample, group) kernel.  The FiLM vectors depend only on t, which is the same for every
// sample of a step (generator/diffusion.py:572), so they are computed once per call by one small CTA.
// The skip of down level 0 is pushed but never popped in the reference (diffusion_utils.py:264-275):
// it is simply not kept.
#include "common.cuh"

namespace dgdm {
// gemm_tc.cu
size_t gemm_tc_image_bytes(int N, int K);
bool gemm_tc_eligible(const GemmArgs& g);
int gemm_tc_pack(const float* W, int N, int K, void* image, cudaStream_t s);
int gemm_tc(const GemmArgs& g, const void* wimg, int x3, int* err, cudaStream_t s);

namespace {

constexpr int PADL = 2;          // zero rows before/after each sample
constexpr int CHUNK = 16384;     // samples per pass (bounds workspace: 1.7 GB at P=14, 3.9 GB at P=42)
constexpr int DSED = 32;

// mish(x) = x tanh(softplus(x)); with e = exp(x): tanh(log(1+e)) = n / (n + 2), n = e (e + 2).
// One exp and one division instead of exp + log1p + tanh; x > 20 returns x (torch's softplus threshold).
__device__ __forceinline__ float mish(float x) {
  if (x > 20.f) return x;
  const float e = expf(x);
  const float n = e * (e + 2.f);
  return x * (n / (n + 2.f));
}

// cond = Linear(Mish(Linear(SinusoidalPosEmb(t))));  film[b] = Linear_b(Mish(cond)) for the 8 res blocks.
// One CTA of 256 threads; everything is a few thousand MACs.
struct FilmArgs {
  const float* se_w0; const float* se_b0; const float* se_w1; const float* se_b1;
  const float* film_w[8]; const float* film_b[8]; int cout[8]; float* film[8];
};
__global__ void __launch_bounds__(256) film_kernel(FilmArgs a, float t) {
  __shared__ float emb[DSED], hid[DSED * 4], cond[DSED];
  const int tid = threadIdx.x;
  if (tid < DSED / 2) {
    // diffusion_utils.py:31-37: w_i = exp(-ln(1e4)/(half-1) * i); [sin | cos]
    float wv = expf((float)tid * -(9.210340371976184f / (float)(DSED / 2 - 1)));
    emb[tid] = sinf(t * wv);
    emb[DSED / 2 + tid] = cosf(t * wv);
  }
  __syncthreads();
  if (tid < DSED * 4) {
    float s = a.se_b0[tid];
    for (int k = 0; k < DSED; ++k) s = fmaf(a.se_w0[tid * DSED + k], emb[k], s);
    hid[tid] = mish(s);
  }
  __syncthreads();
  if (tid < DSED) {
    float s = a.se_b1[tid];
    for (int k = 0; k < DSED * 4; ++k) s = fmaf(a.se_w1[tid * DSED * 4 + k], hid[k], s);
    cond[tid] = mish(s);      // every cond_encoder starts with Mish (diffusion_utils.py:89-91)
  }
  __syncthreads();
  {
    const int b = blockIdx.x;                 // one CTA per residual block (cond is recomputed per CTA: ~5k MACs)
    for (int j = tid; j < 2 * a.cout[b]; j += blockDim.x) {
      float s = a.film_b[b][j];
      for (int k = 0; k < DSED; ++k) s = fmaf(a.film_w[b][j * DSED + k], cond[k], s);
      a.film[b][j] = s;
    }
  }
}

// x [n,P] -> padded channels-last buffer with C = 1: out[b][PADL + l][0] (pads zeroed separately)
__global__ void load_input_kernel(float* __restrict__ out, const float* __restrict__ x, int64_t n, int L) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * L) return;
  int64_t b = i / L; int l = (int)(i % L);
  out[b * (L + 2 * PADL) + PADL + l] = x[i];
}

// GroupNorm(8, C) over (C/8 channels x L positions) per sample, then Mish, optional FiLM (scale,shift per
// channel), optional residual add.  One warp per (sample, group).  in/out/res: padded channels-last,
// leading dims ldi/ldo/ldr (floats per row) and channel offsets folded into the pointers.
struct GnArgs {
  const float* in; float* out; const float* res; const float* gamma; const float* beta; const float* film;
  int64_t n; int L, C, ldi, ldo, ldr;
};
// One warp per (sample, group); the group's L x C/8 block is read ONCE with 128-bit loads and kept in
// registers (MAXV float4 per lane) for the mean, the centred variance and the normalise/Mish/FiLM/residual pass.
template <int MAXV>
__global__ void __launch_bounds__(256) gn_mish_kernel(GnArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= a.n * 8) return;
  const int64_t b = wid >> 3;
  const int grp = (int)(wid & 7);
  const int cg = a.C / 8, cg4 = cg / 4, cnt4 = cg4 * a.L, cnt = cg * a.L;
  const int rows = a.L + 2 * PADL;
  const float* ip = a.in + (b * rows + PADL) * a.ldi + grp * cg;
  float4 x[MAXV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int j = lane + 32 * k;
    x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < cnt4) {
      x[k] = *reinterpret_cast<const float4*>(ip + (int64_t)(j / cg4) * a.ldi + (j % cg4) * 4);
      s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)cnt;
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    if (lane + 32 * k < cnt4) {
      float d0 = x[k].x - mean, d1 = x[k].y - mean, d2 = x[k].z - mean, d3 = x[k].w - mean;
      v = fmaf(d0, d0, v); v = fmaf(d1, d1, v); v = fmaf(d2, d2, v); v = fmaf(d3, d3, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / (float)cnt + 1e-5f);
  float* op = a.out + (b * rows + PADL) * a.ldo + grp * cg;
  const float* rp = a.res ? a.res + (b * rows + PADL) * a.ldr + grp * cg : nullptr;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int j = lane + 32 * k;
    if (j >= cnt4) continue;
    const int l = j / cg4, c = (j % cg4) * 4, ch = grp * cg + c;
    const float4 gm = *reinterpret_cast<const float4*>(a.gamma + ch), bt = *reinterpret_cast<const float4*>(a.beta + ch);
    float4 y;
    y.x = mish((x[k].x - mean) * rstd * gm.x + bt.x);
    y.y = mish((x[k].y - mean) * rstd * gm.y + bt.y);
    y.z = mish((x[k].z - mean) * rstd * gm.z + bt.z);
    y.w = mish((x[k].w - mean) * rstd * gm.w + bt.w);
    if (a.film) {
      const float4 sc = *reinterpret_cast<const float4*>(a.film + ch), sh = *reinterpret_cast<const float4*>(a.film + a.C + ch);
      y.x = sc.x * y.x + sh.x; y.y = sc.y * y.y + sh.y; y.z = sc.z * y.z + sh.z; y.w = sc.w * y.w + sh.w;
    }
    if (rp) {
      const float4 r = *reinterpret_cast<const float4*>(rp + (int64_t)l * a.ldr + c);
      y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    *reinterpret_cast<float4*>(op + (int64_t)l * a.ldo + c) = y;
  }
}

int launch_gn(const GnArgs& a, cudaStream_t s) {
  const int cnt4 = (a.C / 8 / 4) * a.L;
  const unsigned blocks = (unsigned)((a.n * 8 * 32 + 255) / 256);
  if (cnt4 <= 64) gn_mish_kernel<2><<<blocks, 256, 0, s>>>(a);
  else if (cnt4 <= 192) gn_mish_kernel<6><<<blocks, 256, 0, s>>>(a);
  else { set_error("GroupNorm block of %d float4 is larger than supported (P <= 48)", cnt4); return DGDM_EUNSUPPORTED; }
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}

// eps[b,l] = sum_c h[b][l][c] * w[c] + bias      (final_conv.1, Conv1d(128,1,1))
__global__ void __launch_bounds__(256) out_proj_kernel(float* __restrict__ eps, const float* __restrict__ h,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       int64_t n, int L, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= n * L) return;
  const int64_t b = wid / L; const int l = (int)(wid % L);
  const float* hp = h + ((b * (L + 2 * PADL)) + PADL + l) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(hp[c], w[c], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) eps[wid] = s + bias[0];
}

inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// Implicit-GEMM conv over padded channels-last buffers.
//   out[b][PADL + l*o_step + o_off][co] = bias[co] + sum_{tap,ci} w[co][tap][ci] * in[b][l*i_step + i_off + tap][ci]
// l in [0, Lout); in has Lin+4 rows of ldi floats, out has Lo+4 rows of ldo floats.
struct TcCtx { const uint8_t* img; int x3; int* err; };   // img == nullptr: CUDA-core path

int conv_gemm(const float* in, int ldi, int Lin_rows, int cin, int taps, int i_step, int i_off, const float* wgt,
              const float* bias, float* out, int ldo, int Lo_rows, int cout, int o_step, int o_off, int64_t n,
              int Lout, cudaStream_t s, const TcCtx* tc = nullptr, size_t tc_off = 0) {
  GemmArgs g{};
  g.A = in + (int64_t)i_off * ldi; g.W = wgt; g.bias = bias; g.C = out + (int64_t)(PADL + o_off) * ldo;
  g.mask = nullptr; g.add = nullptr;
  g.M = n * Lout; g.N = cout; g.K = taps * cin;
  g.a_lr = Lout; g.a_ss = (int64_t)Lin_rows * ldi; g.a_rs = (int64_t)i_step * ldi; g.a_ts = ldi; g.a_ct = cin;
  g.c_lr = Lout; g.c_ss = (int64_t)Lo_rows * ldo; g.c_rs = (int64_t)o_step * ldo;
  g.m_lr = g.c_lr; g.m_ss = g.c_ss; g.m_rs = g.c_rs;
  g.act = ACT_NONE;
  if (tc && tc->img && gemm_tc_eligible(g)) return gemm_tc(g, tc->img + tc_off, tc->x3, tc->err, s);
  return gemm_f32(g, s);
}

// Tensor-core weight image: one entry per conv whose contraction is tcgen05-eligible (Cin % 64 == 0).
struct TcPlan {
  size_t conv0[8], conv1[8], res[8], down, up[2], fin, bytes;
};
TcPlan make_tc_plan(const dgdm_unet_weights* w) {
  TcPlan p{};
  size_t off = 0;
  auto take = [&](int N, int K) { size_t o = off; off += align_up(gemm_tc_image_bytes(N, K), 1024); return o; };
  for (int b = 0; b < 8; ++b) {
    const int ci = w->blocks[b].cin, co = w->blocks[b].cout;
    p.conv0[b] = (ci % 64 == 0) ? take(co, 5 * ci) : (size_t)-1;
    p.conv1[b] = take(co, 5 * co);
    p.res[b] = (w->blocks[b].res_w && ci % 64 == 0) ? take(co, ci) : (size_t)-1;
  }
  p.down = take(128, 3 * 128);
  p.up[0] = take(128, 2 * 128);
  p.up[1] = take(128, 2 * 128);
  p.fin = take(128, 5 * 128);
  p.bytes = off;
  return p;
}

// Every buffer keeps ONE (rows, leading-dim) view for its whole life so its pad rows stay zero.
struct Bufs {
  float *x0;                  // [n][L+4][1]
  float *a, *b, *c;           // [n][L+4][128]    full resolution
  float *p, *ta, *tb;         // [n][L/2+4][128]  half resolution, 128 wide
  float *q, *r, *u;           // [n][L/2+4][256]  half resolution, 256 wide
  float *cat;                 // [n][L/2+4][512]  concat(x, skip)
  float *film[8];
};

size_t bufs_floats(int64_t n, int L) {
  const int L2 = L / 2;
  const size_t x0 = ((size_t)n * (L + 4) + 63) / 64 * 64;     // keeps every later buffer 256-byte aligned
  return x0 + (size_t)n * (L + 4) * (3 * 128) + (size_t)n * (L2 + 4) * (3 * 128 + 3 * 256 + 512);
}

// One ConditionalResidualBlock1D (diffusion_utils.py:100-120).
//   in  : padded buffer (ld ldi) holding cin channels, L positions
//   out : padded buffer (ld ldo) receiving cout channels
//   t0,t1: scratch padded buffers with ld = cout
int res_block(const dgdm_unet_resblock& w, const float* film, const float* in, int ldi, float* out, int ldo, float* t0,
              float* t1, int64_t n, int L, cudaStream_t s, const TcCtx* tc, const TcPlan& pl, int bi) {
  const int rows = L + 2 * PADL;
  const int ci = w.cin, co = w.cout;
  // block 0: conv5 -> GN -> Mish -> FiLM
  // (Cin = 1 for the very first conv: K = 5 scalars per row, CUDA-core path)
  DGDM_TRY(conv_gemm(in, ldi, rows, ci, 5, 1, 0, w.conv0_w, w.conv0_b, t0, co, rows, co, 1, 0, n, L, s,
                     pl.conv0[bi] == (size_t)-1 ? nullptr : tc, pl.conv0[bi]));
  GnArgs g0{t0, t0, nullptr, w.gn0_w, w.gn0_b, film, n, L, co, co, co, 0};
  DGDM_TRY(launch_gn(g0, s));
  // block 1: conv5 -> GN -> Mish, + residual
  DGDM_TRY(conv_gemm(t0, co, rows, co, 5, 1, 0, w.conv1_w, w.conv1_b, t1, co, rows, co, 1, 0, n, L, s, tc, pl.conv1[bi]));
  const float* res = in;
  int ldr = ldi;
  if (w.res_w) {   // 1x1 residual conv when cin != cout; reuse t0 (block-0 output is dead after conv1)
    DGDM_TRY(conv_gemm(in, ldi, rows, ci, 1, 1, PADL, w.res_w, w.res_b, t0, co, rows, co, 1, 0, n, L, s,
                       pl.res[bi] == (size_t)-1 ? nullptr : tc, pl.res[bi]));
    res = t0;
    ldr = co;
  }
  GnArgs g1{t1, out, res, w.gn1_w, w.gn1_b, nullptr, n, L, co, co, ldo, ldr};
  DGDM_TRY(launch_gn(g1, s));
  return DGDM_OK;
}

}  // namespace
}  // namespace dgdm

extern "C" size_t dgdm_unet1d_workspace_bytes(int32_t n, int32_t P) {
  using namespace dgdm;
  if (n < 1 || P < 2) return 0;
  int64_t nc = n < CHUNK ? n : CHUNK;
  return align_up(bufs_floats(nc, P) * sizeof(float), 256) + 8 * align_up(2 * 512 * sizeof(float), 256) + 256 + 4096;
}

extern "C" size_t dgdm_unet_tc_image_bytes(const dgdm_unet_weights* w) {
  if (!w) return 0;
  return dgdm::make_tc_plan(w).bytes;
}

extern "C" int dgdm_unet_pack_tc(const dgdm_unet_weights* w, void* image, void* stream) {
  using namespace dgdm;
  DGDM_CHECK_ARG(w && image, "dgdm_unet_pack_tc: null pointer");
  DGDM_CHECK_ARG(((uintptr_t)image) % 128 == 0, "dgdm_unet_pack_tc: image must be 128-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  TcPlan pl = make_tc_plan(w);
  uint8_t* img = (uint8_t*)image;
  for (int b = 0; b < 8; ++b) {
    const int ci = w->blocks[b].cin, co = w->blocks[b].cout;
    if (pl.conv0[b] != (size_t)-1) DGDM_TRY(gemm_tc_pack(w->blocks[b].conv0_w, co, 5 * ci, img + pl.conv0[b], s));
    DGDM_TRY(gemm_tc_pack(w->blocks[b].conv1_w, co, 5 * co, img + pl.conv1[b], s));
    if (pl.res[b] != (size_t)-1) DGDM_TRY(gemm_tc_pack(w->blocks[b].res_w, co, ci, img + pl.res[b], s));
  }
  DGDM_TRY(gemm_tc_pack(w->down_w, 128, 3 * 128, img + pl.down, s));
  DGDM_TRY(gemm_tc_pack(w->up_w, 128, 2 * 128, img + pl.up[0], s));
  DGDM_TRY(gemm_tc_pack(w->up_w + 128 * 2 * 128, 128, 2 * 128, img + pl.up[1], s));
  DGDM_TRY(gemm_tc_pack(w->fin_w, 128, 5 * 128, img + pl.fin, s));
  return DGDM_OK;
}

extern "C" int dgdm_unet1d_forward(const dgdm_unet_weights* w, const float* x, int32_t n, int32_t P, int32_t t,
                                   float* eps, void* workspace, size_t workspace_bytes, int32_t precision,
                                   void* stream) {
  using namespace dgdm;
  DGDM_CHECK_ARG(w && x && eps && workspace, "dgdm_unet1d_forward: null pointer");
  DGDM_CHECK_ARG(n >= 1, "dgdm_unet1d_forward: n=%d", n);
  DGDM_CHECK_ARG(P >= 2 && P % 2 == 0 && P <= 48, "dgdm_unet1d_forward: P=%d must be even (one stride-2 level) and <= 48", P);
  static const int kCin[8] = {1, 128, 128, 256, 256, 256, 512, 128};
  static const int kCout[8] = {128, 128, 256, 256, 256, 256, 128, 128};
  for (int b = 0; b < 8; ++b)
    DGDM_CHECK_ARG(w->blocks[b].cin == kCin[b] && w->blocks[b].cout == kCout[b],
                   "dgdm_unet1d_forward: block %d is %d->%d, expected %d->%d (train.py:80 configuration)", b,
                   w->blocks[b].cin, w->blocks[b].cout, kCin[b], kCout[b]);
  cudaStream_t s = (cudaStream_t)stream;
  const int L = P, L2 = P / 2, R = L + 4, R2 = L2 + 4;
  const int64_t nc = n < CHUNK ? n : CHUNK;
  Arena ar(workspace, workspace_bytes);
  Bufs B{};
  float* big = ar.take<float>(bufs_floats(nc, L));
  for (int b = 0; b < 8; ++b) B.film[b] = ar.take<float>(2 * 512);
  int* tc_err = ar.take<int>(1);
  if (!ar.ok) { set_error("dgdm_unet1d_forward: workspace too small (need >= %zu bytes)", ar.off); return DGDM_EWORKSPACE; }
  {
    float* p = big;
    B.x0 = p; p += ((size_t)nc * R + 63) / 64 * 64;
    B.a = p; p += (size_t)nc * R * 128; B.b = p; p += (size_t)nc * R * 128; B.c = p; p += (size_t)nc * R * 128;
    B.p = p; p += (size_t)nc * R2 * 128; B.ta = p; p += (size_t)nc * R2 * 128; B.tb = p; p += (size_t)nc * R2 * 128;
    B.q = p; p += (size_t)nc * R2 * 256; B.r = p; p += (size_t)nc * R2 * 256; B.u = p; p += (size_t)nc * R2 * 256;
    B.cat = p;
  }
  DGDM_CHECK_ARG(precision == DGDM_PREC_FP32_SIMT || precision == DGDM_PREC_BF16X3 || precision == DGDM_PREC_BF16,
                 "dgdm_unet1d_forward: unknown precision %d", precision);
  DGDM_CHECK_ARG(precision == DGDM_PREC_FP32_SIMT || w->tc_image, "dgdm_unet1d_forward: tensor-core precision needs "
                 "weights->tc_image (dgdm_unet_pack_tc)");
  const TcPlan pl = make_tc_plan(w);
  TcCtx tcc{precision == DGDM_PREC_FP32_SIMT ? nullptr : (const uint8_t*)w->tc_image, precision == DGDM_PREC_BF16X3, tc_err};
  const TcCtx* tc = &tcc;
  if (tcc.img) DGDM_CUDA(cudaMemsetAsync(tc_err, 0, sizeof(int), s));
  // FiLM vectors for all 8 blocks, once per call
  FilmArgs fa{w->se_w0, w->se_b0, w->se_w1, w->se_b1, {}, {}, {}, {}};
  for (int b = 0; b < 8; ++b) {
    fa.film_w[b] = w->blocks[b].film_w; fa.film_b[b] = w->blocks[b].film_b; fa.cout[b] = w->blocks[b].cout;
    fa.film[b] = B.film[b];
  }
  film_kernel<<<8, 256, 0, s>>>(fa, (float)t);
  DGDM_LAUNCH_CHECK();
  // pads must be zero; interiors are always fully overwritten before they are read
  DGDM_CUDA(cudaMemsetAsync(big, 0, bufs_floats(nc, L) * sizeof(float), s));

  for (int64_t n0 = 0; n0 < n; n0 += nc) {
    const int64_t m = n - n0 < nc ? n - n0 : nc;
    load_input_kernel<<<nblk(m * L, 256), 256, 0, s>>>(B.x0, x + n0 * L, m, L);
    DGDM_LAUNCH_CHECK();
    // down level 0 @L: Res(1->128), Res(128->128), Downsample
    DGDM_TRY(res_block(w->blocks[0], B.film[0], B.x0, 1, B.a, 128, B.b, B.c, m, L, s, tc, pl, 0));
    DGDM_TRY(res_block(w->blocks[1], B.film[1], B.a, 128, B.a, 128, B.b, B.c, m, L, s, tc, pl, 1));
    //   Conv1d(128,128,3,stride 2,pad 1): out[j] = sum_tap in[2j - 1 + tap]  -> padded row 2j + 1 + tap
    DGDM_TRY(conv_gemm(B.a, 128, R, 128, 3, 2, PADL - 1, w->down_w, w->down_b, B.p, 128, R2, 128, 1, 0, m, L2, s, tc, pl.down));
    // down level 1 @L/2: Res(128->256), Res(256->256); its output is the skip => second half of cat
    DGDM_TRY(res_block(w->blocks[2], B.film[2], B.p, 128, B.q, 256, B.r, B.u, m, L2, s, tc, pl, 2));
    DGDM_TRY(res_block(w->blocks[3], B.film[3], B.q, 256, B.cat + 256, 512, B.r, B.u, m, L2, s, tc, pl, 3));
    // mid @L/2
    DGDM_TRY(res_block(w->blocks[4], B.film[4], B.cat + 256, 512, B.q, 256, B.r, B.u, m, L2, s, tc, pl, 4));
    DGDM_TRY(res_block(w->blocks[5], B.film[5], B.q, 256, B.cat, 512, B.r, B.u, m, L2, s, tc, pl, 5));
    // up level 0 @L/2: cat(x, skip) -> Res(512->128), Res(128->128), Upsample
    DGDM_TRY(res_block(w->blocks[6], B.film[6], B.cat, 512, B.p, 128, B.ta, B.tb, m, L2, s, tc, pl, 6));
    DGDM_TRY(res_block(w->blocks[7], B.film[7], B.p, 128, B.p, 128, B.ta, B.tb, m, L2, s, tc, pl, 7));
    //   ConvTranspose1d(128,128,4,stride 2,pad 1): out[t] = sum_{j,k: t = 2j - 1 + k} in[j] w[k]
    //   even t=2m: taps (k=3, j=m-1), (k=1, j=m); odd t=2m+1: taps (k=2, j=m), (k=0, j=m+1).
    //   up_w is packed [phase][cout][2][cin] (even: k=3,1; odd: k=2,0) so each phase is a 2-tap conv.
    DGDM_TRY(conv_gemm(B.p, 128, R2, 128, 2, 1, PADL - 1, w->up_w, w->up_b, B.a, 128, R, 128, 2, 0, m, L2, s, tc, pl.up[0]));
    DGDM_TRY(conv_gemm(B.p, 128, R2, 128, 2, 1, PADL, w->up_w + 128 * 2 * 128, w->up_b, B.a, 128, R, 128, 2, 1, m, L2, s, tc, pl.up[1]));
    // final: Conv1dBlock(128,128,5) then Conv1d(128,1,1)
    DGDM_TRY(conv_gemm(B.a, 128, R, 128, 5, 1, 0, w->fin_w, w->fin_b, B.b, 128, R, 128, 1, 0, m, L, s, tc, pl.fin));
    GnArgs gf{B.b, B.b, nullptr, w->fin_gn_w, w->fin_gn_b, nullptr, m, L, 128, 128, 128, 0};
    DGDM_TRY(launch_gn(gf, s));
    out_proj_kernel<<<nblk(m * L * 32, 256), 256, 0, s>>>(eps + n0 * L, B.b, w->out_w, w->out_b, m, L, 128);
    DGDM_LAUNCH_CHECK();
  }
  return DGDM_OK;
}
