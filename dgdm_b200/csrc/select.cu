// K6: best-of-N selection (np.argmax semantics of generator/diffusion.py:346-428): per object the k
// best candidates by score, descending, ties to the lowest candidate index.  One CTA per object;
// k rounds of a block-wide arg-max with warp shuffles (k <= 32, n_cand arbitrary).  NaN scores rank last.
#include <math_constants.h>

#include "common.cuh"

namespace dgdm {
namespace {

struct Best { float v; int i; };

__device__ __forceinline__ bool better(const Best& a, const Best& b) {
  return a.v > b.v || (a.v == b.v && a.i < b.i);
}

__global__ void __launch_bounds__(256) topk_kernel(const float* __restrict__ scores, int n_cand, int k,
                                                   int64_t* __restrict__ idx, float* __restrict__ best) {
  __shared__ Best warp_best[8];
  __shared__ Best prev;        // last selected (value, index); selection proceeds in strict descending order
  const float* s = scores + (int64_t)blockIdx.x * n_cand;
  if (threadIdx.x == 0) prev = Best{CUDART_INF_F, -1};
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    Best p = prev;
    Best b{-CUDART_INF_F, 0x7fffffff};
    for (int i = threadIdx.x; i < n_cand; i += blockDim.x) {
      float v = s[i];
      if (!(v == v)) v = -CUDART_INF_F;                 // NaN ranks last
      Best c{v, i};
      // eligible iff strictly after prev in (descending value, ascending index) order
      bool elig = (p.i < 0) || v < p.v || (v == p.v && i > p.i);
      if (elig && better(c, b)) b = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Best c{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
      if (better(c, b)) b = c;
    }
    if ((threadIdx.x & 31) == 0) warp_best[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
      Best w = warp_best[0];
      for (int j = 1; j < 8; ++j) if (better(warp_best[j], w)) w = warp_best[j];
      bool none = w.i == 0x7fffffff;
      idx[(int64_t)blockIdx.x * k + r] = none ? -1 : w.i;
      if (best) best[(int64_t)blockIdx.x * k + r] = none ? -CUDART_INF_F : s[w.i];
      prev = none ? Best{-CUDART_INF_F, 0x7ffffffe} : w;
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace dgdm

extern "C" int dgdm_best_of_n(const float* scores, int32_t n_obj, int32_t n_cand, int32_t k, int64_t* idx,
                              float* best, void* stream) {
  using namespace dgdm;
  DGDM_CHECK_ARG(scores && idx, "dgdm_best_of_n: null pointer");
  DGDM_CHECK_ARG(n_obj >= 0 && n_cand >= 1, "dgdm_best_of_n: bad sizes n_obj=%d n_cand=%d", n_obj, n_cand);
  DGDM_CHECK_ARG(k >= 1 && k <= 32 && k <= n_cand, "dgdm_best_of_n: k=%d must be in [1, min(32, n_cand)]", k);
  if (n_obj == 0) return DGDM_OK;
  topk_kernel<<<n_obj, 256, 0, (cudaStream_t)stream>>>(scores, n_cand, k, idx, best);
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}
