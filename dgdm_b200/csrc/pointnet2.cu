// K5: PointNet++ SSG object encoder, evaluated ONCE per object (it depends on neither x nor t, so the
// reference's per-guidance-row evaluation -- 99.8% of its 3D FLOPs -- is hoisted).
// Replaces PointNet2.forward (dynamics/models/pointnet2.py:21-31) and PointNetSetAbstraction.forward
// (pointnet2_utils.py:184-210):
//   sa1: FPS 512 of N, ball r=0.2 k=32, MLP 3->64->128, max over k
//   sa2: FPS 128 of 512, ball r=0.4 k=64, MLP 131->128->256, max over k
//   sa3: group-all over 128 points, MLP 259->256, max
// with BatchNorm2d(eval) folded into each 1x1 conv.  Index work (FPS, ball query) follows the
// reference's arithmetic exactly: squared distances as sum((p-c)^2) for FPS (pointnet2_utils.py:87) and
// in the expanded form -2 a.b + |a|^2 + |b|^2 for the ball query (:27-48, :107-108), no FMA contraction,
// first-maximum / ascending-index tie rules.
#include "common.cuh"

namespace dgdm {
namespace {

constexpr int NPTS = 512;

// One CTA (512 threads) per cloud; thread i owns point i.  npoint sequential rounds of
// "update min-distance, block arg-max (first maximum wins)".
__global__ void __launch_bounds__(NPTS) fps_kernel(const float* __restrict__ xyz, int n_in, int npoint,
                                                   const int64_t* __restrict__ start, int start_stride,
                                                   int32_t* __restrict__ out_idx, float* __restrict__ out_xyz) {
  __shared__ float sx[NPTS], sy[NPTS], sz[NPTS];
  __shared__ float wv[16];
  __shared__ int wi[16];
  __shared__ int s_far;
  const int b = blockIdx.x, i = threadIdx.x;
  const float* p = xyz + (int64_t)b * n_in * 3;
  const bool live = i < n_in;
  float x = 0.f, y = 0.f, z = 0.f;
  if (live) { x = p[i * 3]; y = p[i * 3 + 1]; z = p[i * 3 + 2]; sx[i] = x; sy[i] = y; sz[i] = z; }
  float dist = 1e10f;
  if (i == 0) s_far = (int)start[(int64_t)b * start_stride];
  __syncthreads();
  for (int it = 0; it < npoint; ++it) {
    const int far = s_far;
    if (i == 0) {
      out_idx[(int64_t)b * npoint + it] = far;
      float* o = out_xyz + ((int64_t)b * npoint + it) * 3;
      o[0] = sx[far]; o[1] = sy[far]; o[2] = sz[far];
    }
    float dx = __fsub_rn(x, sx[far]), dy = __fsub_rn(y, sy[far]), dz = __fsub_rn(z, sz[far]);
    float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (live && d < dist) dist = d;
    float bv = live ? dist : -1.f;
    int bi = i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();                 // everyone has read s_far
    if ((i & 31) == 0) { wv[i >> 5] = bv; wi[i >> 5] = bi; }
    __syncthreads();
    if (i < 32) {
      float v = i < (NPTS / 32) ? wv[i] : -2.f;
      int k = i < (NPTS / 32) ? wi[i] : 0x7fffffff;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, k, o);
        if (ov > v || (ov == v && oi < k)) { v = ov; k = oi; }
      }
      if (i == 0) s_far = k;
    }
    __syncthreads();
  }
}

// One thread per (cloud, query): first `nsample` source indices (ascending) with d2 <= r^2, padded with
// the first hit.  d2 in the reference's expanded form.
__global__ void ball_query_kernel(const float* __restrict__ xyz, int n_in, const float* __restrict__ qxyz, int S,
                                  float r2, int nsample, int32_t* __restrict__ gidx, int64_t total) {
  int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  const int64_t b = id / S;
  const float* p = xyz + b * n_in * 3;
  const float qx = qxyz[id * 3], qy = qxyz[id * 3 + 1], qz = qxyz[id * 3 + 2];
  const float qn = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
  int32_t* out = gidx + id * nsample;
  int cnt = 0, first = 0;
  for (int j = 0; j < n_in && cnt < nsample; ++j) {
    float x = p[j * 3], y = p[j * 3 + 1], z = p[j * 3 + 2];
    float dot = __fadd_rn(__fadd_rn(__fmul_rn(qx, x), __fmul_rn(qy, y)), __fmul_rn(qz, z));
    float pn = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    float d2 = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), qn), pn);
    if (!(d2 > r2)) {
      if (cnt == 0) first = j;
      out[cnt++] = j;
    }
  }
  // a query point is always within its own ball, so cnt >= 1 for FPS-sampled queries
  for (; cnt < nsample; ++cnt) out[cnt] = first;
}

// grouped[(b,s,k), :] = [xyz[idx] - q(b,s) (3) | feat[idx] (D)]   (centroid-relative xyz first, :136-140)
__global__ void group_kernel(float* __restrict__ grouped, const float* __restrict__ xyz, const float* __restrict__ feat,
                             int n_in, int D, const float* __restrict__ qxyz, const int32_t* __restrict__ gidx, int S,
                             int nsample, int64_t total_rows) {
  const int C = 3 + D;
  int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total_rows * C) return;
  int64_t row = id / C;
  int c = (int)(id % C);
  int64_t bs = row / nsample;       // (b*S + s)
  int64_t b = bs / S;
  int j = gidx[row];
  float v;
  if (c < 3) v = __fsub_rn(xyz[(b * n_in + j) * 3 + c], qxyz[bs * 3 + c]);
  else v = feat[(b * n_in + j) * D + (c - 3)];
  grouped[id] = v;
}

// out[g, c] = max_k in[(g*k_per + k), c]
__global__ void maxpool_kernel(float* __restrict__ out, const float* __restrict__ in, int k_per, int C, int64_t groups) {
  int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= groups * C) return;
  int64_t g = id / C;
  int c = (int)(id % C);
  const float* p = in + g * k_per * C + c;
  float m = p[0];
  for (int k = 1; k < k_per; ++k) m = fmaxf(m, p[(int64_t)k * C]);
  out[id] = m;
}

// rows[(b,j), :] = [xyz(b,j) (3) | feat(b,j) (D)]      (sample_and_group_all, :149-166)
__global__ void concat_kernel(float* __restrict__ out, const float* __restrict__ xyz, const float* __restrict__ feat,
                              int D, int64_t rows) {
  const int C = 3 + D;
  int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= rows * C) return;
  int64_t r = id / C;
  int c = (int)(id % C);
  out[id] = c < 3 ? xyz[r * 3 + c] : feat[r * D + (c - 3)];
}

inline unsigned nb(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

constexpr int OBJ_CHUNK = 64;   // clouds per pass (bounds the grouped-feature scratch)

struct Ws {
  int32_t *fidx1, *fidx2, *gidx;
  float *xyz1, *xyz2, *g0, *g1, *g2, *f1, *f2, *cat3, *h3;
};

size_t ws_bytes(int nc) {
  auto a = [](size_t n, size_t e) { return align_up(n * e, 256); };
  size_t b = 0;
  b += a((size_t)nc * 512, 4) + a((size_t)nc * 128, 4) + a((size_t)nc * 512 * 32, 4);  // fidx1, fidx2, gidx (512*32 == 128*64*2)
  b += a((size_t)nc * 512 * 3, 4) + a((size_t)nc * 128 * 3, 4);                        // xyz1, xyz2
  b += a((size_t)nc * 8192 * 131, 4);                                                  // g0: grouped input (max of 8192*131, 16384*3)
  b += a((size_t)nc * 16384 * 128, 4) * 2;                                             // g1, g2 MLP activations
  b += a((size_t)nc * 512 * 128, 4) + a((size_t)nc * 128 * 256, 4);                    // f1, f2
  b += a((size_t)nc * 128 * 259, 4) + a((size_t)nc * 128 * 256, 4);                    // cat3, h3
  return b;
}

}  // namespace
}  // namespace dgdm

extern "C" size_t dgdm_pointnet2_workspace_bytes(int32_t n_clouds, int32_t n_points) {
  using namespace dgdm;
  if (n_clouds < 1 || n_points != NPTS) return 0;
  return ws_bytes(n_clouds < OBJ_CHUNK ? n_clouds : OBJ_CHUNK) + 4096;
}

extern "C" int dgdm_pointnet2_encode(const dgdm_pointnet2_weights* w, const float* clouds, int32_t n_clouds,
                                     int32_t n_points, const int64_t* fps_start, float* codes, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  using namespace dgdm;
  DGDM_CHECK_ARG(w && clouds && fps_start && codes && workspace, "dgdm_pointnet2_encode: null pointer");
  DGDM_CHECK_ARG(n_clouds >= 1, "dgdm_pointnet2_encode: n_clouds=%d", n_clouds);
  DGDM_CHECK_ARG(n_points == NPTS, "dgdm_pointnet2_encode: n_points=%d, only %d supported "
                                   "(--object_max_num_vertices=512, generator/guided_sample_3d.sh)", n_points, NPTS);
  cudaStream_t s = (cudaStream_t)stream;
  const int ncmax = n_clouds < OBJ_CHUNK ? n_clouds : OBJ_CHUNK;
  Arena ar(workspace, workspace_bytes);
  Ws W{};
  W.fidx1 = ar.take<int32_t>((size_t)ncmax * 512);
  W.fidx2 = ar.take<int32_t>((size_t)ncmax * 128);
  W.gidx = ar.take<int32_t>((size_t)ncmax * 512 * 32);
  W.xyz1 = ar.take<float>((size_t)ncmax * 512 * 3);
  W.xyz2 = ar.take<float>((size_t)ncmax * 128 * 3);
  W.g0 = ar.take<float>((size_t)ncmax * 8192 * 131);
  W.g1 = ar.take<float>((size_t)ncmax * 16384 * 128);
  W.g2 = ar.take<float>((size_t)ncmax * 16384 * 128);
  W.f1 = ar.take<float>((size_t)ncmax * 512 * 128);
  W.f2 = ar.take<float>((size_t)ncmax * 128 * 256);
  W.cat3 = ar.take<float>((size_t)ncmax * 128 * 259);
  W.h3 = ar.take<float>((size_t)ncmax * 128 * 256);
  if (!ar.ok) { set_error("dgdm_pointnet2_encode: workspace too small (need >= %zu bytes)", ar.off); return DGDM_EWORKSPACE; }

  for (int c0 = 0; c0 < n_clouds; c0 += ncmax) {
    const int nc = n_clouds - c0 < ncmax ? n_clouds - c0 : ncmax;
    const float* xyz0 = clouds + (int64_t)c0 * NPTS * 3;
    const int64_t* st = fps_start + (int64_t)c0 * 2;
    // ---- sa1 ----  (radius thresholds: `radius ** 2` in double, rounded once to fp32, as torch compares an fp32
    // tensor with a Python scalar at pointnet2_utils.py:110)
    fps_kernel<<<nc, NPTS, 0, s>>>(xyz0, NPTS, 512, st, 2, W.fidx1, W.xyz1);
    DGDM_LAUNCH_CHECK();
    ball_query_kernel<<<nb((int64_t)nc * 512, 128), 128, 0, s>>>(xyz0, NPTS, W.xyz1, 512, (float)(0.2 * 0.2), 32, W.gidx, (int64_t)nc * 512);
    DGDM_LAUNCH_CHECK();
    int64_t rows = (int64_t)nc * 512 * 32;
    group_kernel<<<nb(rows * 3, 256), 256, 0, s>>>(W.g0, xyz0, nullptr, NPTS, 0, W.xyz1, W.gidx, 512, 32, rows);
    DGDM_LAUNCH_CHECK();
    DGDM_TRY(gemm_f32(gemm_plain(W.g0, 3, w->w[0], w->b[0], W.g1, 64, rows, 64, 3, ACT_RELU), s));
    DGDM_TRY(gemm_f32(gemm_plain(W.g1, 64, w->w[1], w->b[1], W.g2, 128, rows, 128, 64, ACT_RELU), s));
    maxpool_kernel<<<nb((int64_t)nc * 512 * 128, 256), 256, 0, s>>>(W.f1, W.g2, 32, 128, (int64_t)nc * 512);
    DGDM_LAUNCH_CHECK();
    // ---- sa2 ----
    fps_kernel<<<nc, NPTS, 0, s>>>(W.xyz1, 512, 128, st + 1, 2, W.fidx2, W.xyz2);
    DGDM_LAUNCH_CHECK();
    ball_query_kernel<<<nb((int64_t)nc * 128, 128), 128, 0, s>>>(W.xyz1, 512, W.xyz2, 128, (float)(0.4 * 0.4), 64, W.gidx, (int64_t)nc * 128);
    DGDM_LAUNCH_CHECK();
    rows = (int64_t)nc * 128 * 64;
    group_kernel<<<nb(rows * 131, 256), 256, 0, s>>>(W.g0, W.xyz1, W.f1, 512, 128, W.xyz2, W.gidx, 128, 64, rows);
    DGDM_LAUNCH_CHECK();
    DGDM_TRY(gemm_f32(gemm_plain(W.g0, 131, w->w[2], w->b[2], W.g1, 128, rows, 128, 131, ACT_RELU), s));
    DGDM_TRY(gemm_f32(gemm_plain(W.g1, 128, w->w[3], w->b[3], W.g2, 256, rows, 256, 128, ACT_RELU), s));
    maxpool_kernel<<<nb((int64_t)nc * 128 * 256, 256), 256, 0, s>>>(W.f2, W.g2, 64, 256, (int64_t)nc * 128);
    DGDM_LAUNCH_CHECK();
    // ---- sa3 (group all) ----
    rows = (int64_t)nc * 128;
    concat_kernel<<<nb(rows * 259, 256), 256, 0, s>>>(W.cat3, W.xyz2, W.f2, 256, rows);
    DGDM_LAUNCH_CHECK();
    DGDM_TRY(gemm_f32(gemm_plain(W.cat3, 259, w->w[4], w->b[4], W.h3, 256, rows, 256, 259, ACT_RELU), s));
    maxpool_kernel<<<nb((int64_t)nc * 256, 256), 256, 0, s>>>(codes + (int64_t)c0 * 256, W.h3, 128, 256, nc);
    DGDM_LAUNCH_CHECK();
  }
  return DGDM_OK;
}
