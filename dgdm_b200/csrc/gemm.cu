// Generic fp32 CUDA-core GEMM with row/tap mapping (see GemmArgs in common.cuh).
//
// One kernel serves every small or irregular contraction on the path: the gripper/object/time
// encoders and their transposed backward (profile_forward_2d.py:92-107), the layer-1 hoists, the
// exact-fp32 trunk, every Conv1d / ConvTranspose1d of the denoiser as an implicit GEMM over a
// zero-padded channels-last buffer (diffusion_utils.py:42,51,66,96), and the PointNet++ 1x1 convs.
// 128x64x16 tiles, 256 threads, 8x4 register tile, next-tile register prefetch.
#include "common.cuh"

namespace dgdm {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

template <bool VEC>
__global__ void __launch_bounds__(NT) gemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = t % 16, ty = t / 16;

  // A: 2 rows-of-4 per thread; W: 1 row-of-4 per thread
  const int a_kq = (t % 4) * 4;
  int64_t a_base[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int64_t m = m0 + t / 4 + i * 64;
    a_ok[i] = m < g.M;
    int64_t mm = a_ok[i] ? m : 0;
    a_base[i] = (mm / g.a_lr) * g.a_ss + (mm % g.a_lr) * g.a_rs;
  }
  const int w_row = n0 + t / 4;
  const bool w_ok = w_row < g.N;
  const float* w_ptr = g.W + (int64_t)(w_ok ? w_row : 0) * g.K;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra[2], rw;
  auto load_tile = [&](int k0) {
    const int k = k0 + a_kq;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[i]) {
        if (VEC) {
          if (k < g.K) v = *reinterpret_cast<const float4*>(g.A + a_base[i] + (int64_t)(k / g.a_ct) * g.a_ts + (k % g.a_ct));
        } else {
          float e[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            int kk = k + q;
            e[q] = kk < g.K ? g.A[a_base[i] + (int64_t)(kk / g.a_ct) * g.a_ts + (kk % g.a_ct)] : 0.f;
          }
          v = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      ra[i] = v;
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w_ok) {
      if (VEC) {
        if (k < g.K) v = *reinterpret_cast<const float4*>(w_ptr + k);
      } else {
        float e[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) e[q] = (k + q) < g.K ? w_ptr[k + q] : 0.f;
        v = make_float4(e[0], e[1], e[2], e[3]);
      }
    }
    rw = v;
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = t / 4 + i * 64;
      As[a_kq + 0][r] = ra[i].x; As[a_kq + 1][r] = ra[i].y; As[a_kq + 2][r] = ra[i].z; As[a_kq + 3][r] = ra[i].w;
    }
    int r = t / 4;
    Bs[a_kq + 0][r] = rw.x; Bs[a_kq + 1][r] = rw.y; Bs[a_kq + 2][r] = rw.z; Bs[a_kq + 3][r] = rw.w;
  };

  load_tile(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    store_tile();
    __syncthreads();
    if (k0 + BK < g.K) load_tile(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
    int64_t c_off = (m / g.c_lr) * g.c_ss + (m % g.c_lr) * g.c_rs;
    int64_t x_off = (g.mask || g.add) ? (m / g.m_lr) * g.m_ss + (m % g.m_lr) * g.m_rs : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (g.act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (g.act == ACT_SILU) v = v / (1.f + expf(-v));
      if (g.mask) v = g.mask[x_off + n] > 0.f ? v : 0.f;
      if (g.add) v += g.add[x_off + n];
      g.C[c_off + n] = v;
    }
  }
}

}  // namespace

GemmArgs gemm_plain(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc,
                    int64_t M, int N, int K, int act) {
  GemmArgs g{};
  g.A = A; g.W = W; g.bias = bias; g.C = C; g.mask = nullptr; g.add = nullptr;
  g.M = M; g.N = N; g.K = K;
  g.a_lr = (int64_t)1 << 60; g.a_ss = 0; g.a_rs = lda; g.a_ts = 0; g.a_ct = K > 0 ? K : 1;
  g.c_lr = (int64_t)1 << 60; g.c_ss = 0; g.c_rs = ldc;
  g.m_lr = (int64_t)1 << 60; g.m_ss = 0; g.m_rs = ldc;
  g.act = act;
  return g;
}

int gemm_f32(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return DGDM_OK;
  DGDM_CHECK_ARG(g.K > 0 && g.a_ct > 0, "gemm: K=%d a_ct=%d", g.K, g.a_ct);
  dim3 grid((unsigned)((g.M + BM - 1) / BM), (unsigned)((g.N + BN - 1) / BN));
  bool vec = (g.K % 4 == 0) && (g.a_ct % 4 == 0) && (g.a_ss % 4 == 0) && (g.a_rs % 4 == 0) && (g.a_ts % 4 == 0) &&
             (((uintptr_t)g.A) % 16 == 0) && (((uintptr_t)g.W) % 16 == 0);
  if (vec) gemm_kernel<true><<<grid, NT, 0, s>>>(g);
  else gemm_kernel<false><<<grid, NT, 0, s>>>(g);
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}

}  // namespace dgdm

extern "C" int dgdm_linear_f32(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc,
                               int64_t M, int32_t N, int32_t K, int32_t relu, void* stream) {
  DGDM_CHECK_ARG(A && W && C, "dgdm_linear_f32: null pointer");
  DGDM_CHECK_ARG(lda >= K && ldc >= N, "dgdm_linear_f32: lda/ldc too small");
  return dgdm::gemm_f32(dgdm::gemm_plain(A, lda, W, bias, C, ldc, M, N, K, relu ? dgdm::ACT_RELU : dgdm::ACT_NONE), (cudaStream_t)stream);
}
