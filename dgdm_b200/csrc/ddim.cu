// K4: fused guided DDIM update (generator/diffusion.py:575-576 + diffusers DDIMScheduler.step, eta=0).
//
// HBM-bound elementwise kernel: 16 algorithmic bytes per control point (read x, eps, grad; write
// x_prev).  128-bit loads/stores, grid-stride, grid = 148 SMs x 8 resident CTAs.  Arithmetic uses
// explicit round-to-nearest mul/add/div intrinsics (no FMA contraction) in the reference's operation
// order, so the result is bit-identical to the torch-CPU expression.
#include "common.cuh"

namespace dgdm {
namespace {

struct DdimCoef { float s1mat, sat, saprev, s1maprev, scale; int clip, has_grad; };

__device__ __forceinline__ float ddim_one(float x, float e, float g, const DdimCoef& c) {
  // eps_hat = eps - (sqrt(1-a_t) * grad) * scale
  float eh = c.has_grad ? __fsub_rn(e, __fmul_rn(__fmul_rn(c.s1mat, g), c.scale)) : e;
  // x0 = (x - sqrt(1-a_t) * eps_hat) / sqrt(a_t), clamped
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c.s1mat, eh)), c.sat);
  if (c.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
  // x_prev = sqrt(a_prev) * x0 + sqrt(1-a_prev) * eps_hat
  return __fadd_rn(__fmul_rn(c.saprev, x0), __fmul_rn(c.s1maprev, eh));
}

__global__ void __launch_bounds__(256) ddim_kernel_v4(float4* __restrict__ out, const float4* __restrict__ x,
                                                      const float4* __restrict__ eps, const float4* __restrict__ grad,
                                                      int64_t n4, DdimCoef c) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 xv = x[i], ev = eps[i];
    float4 gv = c.has_grad ? grad[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = ddim_one(xv.x, ev.x, gv.x, c);
    o.y = ddim_one(xv.y, ev.y, gv.y, c);
    o.z = ddim_one(xv.z, ev.z, gv.z, c);
    o.w = ddim_one(xv.w, ev.w, gv.w, c);
    out[i] = o;
  }
}

__global__ void __launch_bounds__(256) ddim_kernel_v1(float* __restrict__ out, const float* __restrict__ x,
                                                      const float* __restrict__ eps, const float* __restrict__ grad,
                                                      int64_t n, DdimCoef c) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ddim_one(x[i], eps[i], c.has_grad ? grad[i] : 0.f, c);
}

}  // namespace
}  // namespace dgdm

extern "C" int dgdm_ddim_guided_update(float* x_prev, const float* x, const float* eps, const float* grad, int64_t n,
                                       float sqrt_1m_at, float sqrt_at, float sqrt_aprev, float sqrt_1m_aprev,
                                       float scale, int clip, void* stream) {
  using namespace dgdm;
  DGDM_CHECK_ARG(x_prev && x && eps, "dgdm_ddim_guided_update: null pointer");
  DGDM_CHECK_ARG(n >= 0, "dgdm_ddim_guided_update: n < 0");
  DGDM_CHECK_ARG(sqrt_at > 0.f, "dgdm_ddim_guided_update: sqrt_at must be > 0");
  if (n == 0) return DGDM_OK;
  DdimCoef c{sqrt_1m_at, sqrt_at, sqrt_aprev, sqrt_1m_aprev, scale, clip, grad != nullptr};
  cudaStream_t s = (cudaStream_t)stream;
  const int max_blocks = 148 * 8;
  bool vec = (n % 4 == 0) && (((uintptr_t)x_prev | (uintptr_t)x | (uintptr_t)eps | (uintptr_t)grad) % 16 == 0);
  if (vec) {
    int64_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256 < max_blocks ? (n4 + 255) / 256 : max_blocks);
    ddim_kernel_v4<<<blocks, 256, 0, s>>>((float4*)x_prev, (const float4*)x, (const float4*)eps, (const float4*)grad, n4, c);
  } else {
    int blocks = (int)((n + 255) / 256 < max_blocks ? (n + 255) / 256 : max_blocks);
    ddim_kernel_v1<<<blocks, 256, 0, s>>>(x_prev, x, eps, grad, n, c);
  }
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}
