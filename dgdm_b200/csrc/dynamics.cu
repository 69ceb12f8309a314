// K1 + K2 host orchestration and the exact-fp32 (CUDA-core) trunk.
//
// Replaces Diffusion.cond_fn (generator/diffusion.py:473-504) and the profile pass of
// get_convergence_centers (:506-531) over the dynamics networks of dynamics/profile_forward_{2d,3d}.py.
//
// Exact restructurings, all row-invariant work hoisted out of the B*G guidance rows:
//   layer-1 pre-activation(row = pair p, pose g) = Cst[obj(p)] + U[design(p)] + V[g]
//     U   = w1_ctrl . gripper_encoder(x)             per design          (no GEMM at B*G scale)
//     V   = w1_pose . fourier(ori_g, pos_g)          per pose-grid row
//     Cst = w1_obj . object_code + w1_time . time_emb + b1   per object, per step
//   trunk layers 2..8 + output, objective, input-gradient back to layer 1, sum over g   -> K1 (this file: SIMT
//     path; dynamics_tc.cu: tcgen05 path), then K2 folds pairs into designs and the tiny encoder backward
//     turns dU into d/dx.
#include "common.cuh"

namespace dgdm {

// dynamics_tc.cu
int tc_trunk(const dgdm_dyn_weights* w, const float* U, const float* Cst, const float* V, int n_designs, int n_obj,
             int opd, const int32_t* pair_object, int G, const dgdm_objective* obj, bool backward, float* dUp,
             float* score_sum, float* logits, void* ws, size_t ws_bytes, int precision, cudaStream_t s);
size_t tc_trunk_workspace_bytes(int H1, int64_t n_pairs, int G);
// gemm_tc.cu
size_t gemm_tc_image_bytes(int N, int K);
bool gemm_tc_eligible(const GemmArgs& g);
int gemm_tc_pack(const float* W, int N, int K, void* image, cudaStream_t s);
int gemm_tc(const GemmArgs& g, const void* wimg, int x3, int* err, cudaStream_t s);

namespace {

constexpr int W256 = 256;

__device__ __forceinline__ int pair_obj(int64_t p, int opd, int n_designs, int n_obj, const int32_t* pair_object) {
  if (pair_object) return pair_object[p];
  if (opd == 1) return (int)(p / (n_designs / n_obj));
  return (int)(p % opd);
}

// pose[g, 0:27] = [ori, sin/cos(2^k ori)_{k<4}, pos(2), sin/cos(2^k pos)_{k<4}]  (profile_forward_2d.py:29-38,149-151)
__global__ void pose_embed_kernel(float* __restrict__ pose, dgdm_pose_grid grid, int G) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  int np2 = grid.pos_zero ? 1 : grid.num_pos * grid.num_pos;
  int o = g / np2, rem = g % np2;
  // torch.linspace: start + i*step for the first half, end - (n-1-i)*step for the second (step in fp32)
  auto lin = [](float lo, float hi, int n, int i) -> float {
    if (n == 1) return lo;
    float step = (hi - lo) / (float)(n - 1);
    return i < n / 2 ? lo + step * (float)i : hi - step * (float)(n - 1 - i);
  };
  float ori = lin(grid.ori_lo, grid.ori_hi, grid.grid_size, o);
  float px = 0.f, py = 0.f;
  if (!grid.pos_zero) {
    px = lin(-1.f, 1.f, grid.num_pos, rem / grid.num_pos);
    py = lin(-1.f, 1.f, grid.num_pos, rem % grid.num_pos);
  }
  float* out = pose + (int64_t)g * 27;
  out[0] = ori;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float f = (float)(1 << k);
    out[1 + 2 * k] = sinf(ori * f);
    out[2 + 2 * k] = cosf(ori * f);
  }
  out[9] = px; out[10] = py;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float f = (float)(1 << k);
    out[11 + 4 * k] = sinf(px * f); out[12 + 4 * k] = sinf(py * f);
    out[13 + 4 * k] = cosf(px * f); out[14 + 4 * k] = cosf(py * f);
  }
}

// timestep_embedding(t, dim): [cos(t f_i), sin(t f_i)], f_i = exp(-ln(1e4) i / half)  (profile_forward_2d.py:58-76)
// V [G][H1] -> Vq [H1/4][G][4]: the tensor-core trunk reads four layer-1 features of a pose row with one 128-bit load,
// consecutive lanes = consecutive pose rows (512 contiguous bytes per warp request).
__global__ void pose_table_quads_kernel(float4* __restrict__ vq, const float* __restrict__ V, int G, int H1) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= G * (H1 / 4)) return;
  const int c4 = idx / G, g = idx - c4 * G;
  vq[idx] = *reinterpret_cast<const float4*>(V + (size_t)g * H1 + 4 * c4);
}

__global__ void time_embed_kernel(float* __restrict__ out, float t, int dim) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int half = dim / 2;
  if (i >= half) return;
  float f = expf(-9.210340371976184f * (float)i / (float)half);
  float a = t * f;
  out[i] = cosf(a);
  out[half + i] = sinf(a);
}

// explicit rows: pose[r, 0:27] from (ori[r], pos[r,0:2]); same layout as pose_embed_kernel
__global__ void pose_embed_rows_kernel(float* __restrict__ pose, const float* __restrict__ ori, const float* __restrict__ pos, int64_t n) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float o = ori[r], px = pos[2 * r], py = pos[2 * r + 1];
  float* out = pose + r * 27;
  out[0] = o;
#pragma unroll
  for (int k = 0; k < 4; ++k) { float f = (float)(1 << k); out[1 + 2 * k] = sinf(o * f); out[2 + 2 * k] = cosf(o * f); }
  out[9] = px; out[10] = py;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float f = (float)(1 << k);
    out[11 + 4 * k] = sinf(px * f); out[12 + 4 * k] = sinf(py * f);
    out[13 + 4 * k] = cosf(px * f); out[14 + 4 * k] = cosf(py * f);
  }
}
// explicit rows: temb[r, 0:dim] = [cos(t_r f_i) | sin(t_r f_i)]
__global__ void time_embed_rows_kernel(float* __restrict__ out, const float* __restrict__ t, int64_t n, int dim) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= n * half) return;
  const int64_t r = idx / half; const int i = (int)(idx % half);
  const float a = t[r] * expf(-9.210340371976184f * (float)i / (float)half);
  out[r * dim + i] = cosf(a);
  out[r * dim + half + i] = sinf(a);
}
// base[r,:] = ((o[r,:] + u[r,:]) + v[r,:]) + tt[r,:]
__global__ void sum4_kernel(float* __restrict__ base, const float* __restrict__ o, const float* __restrict__ u,
                            const float* __restrict__ v, const float* __restrict__ tt, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<const float4*>(o)[i], b = reinterpret_cast<const float4*>(u)[i];
  float4 c = reinterpret_cast<const float4*>(v)[i], d = reinterpret_cast<const float4*>(tt)[i];
  reinterpret_cast<float4*>(base)[i] = make_float4(((a.x + b.x) + c.x) + d.x, ((a.y + b.y) + c.y) + d.y,
                                                   ((a.z + b.z) + c.z) + d.z, ((a.w + b.w) + c.w) + d.w);
}

// a1[r,:] = relu(Cst[obj(p)] + U[design(p)] + V[g]),  r = p*G + g, rows [r0, r0+rows)
__global__ void build_a1_kernel(float* __restrict__ a1, const float* __restrict__ U, const float* __restrict__ Cst,
                                const float* __restrict__ V, int64_t r0, int64_t rows, int G, int H1, int opd,
                                int n_designs, int n_obj, const int32_t* __restrict__ pair_object) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // float4 index
  int h4 = H1 / 4;
  if (idx >= rows * h4) return;
  int64_t r = r0 + idx / h4;
  int c = (int)(idx % h4) * 4;
  int64_t p = r / G;
  int g = (int)(r % G);
  int d = (int)(p / opd);
  int o = pair_obj(p, opd, n_designs, n_obj, pair_object);
  float4 u = *reinterpret_cast<const float4*>(U + (int64_t)d * H1 + c);
  float4 k = *reinterpret_cast<const float4*>(Cst + (int64_t)o * H1 + c);
  float4 v = *reinterpret_cast<const float4*>(V + (int64_t)g * H1 + c);
  float4 y;
  y.x = fmaxf((k.x + u.x) + v.x, 0.f);
  y.y = fmaxf((k.y + u.y) + v.y, 0.f);
  y.z = fmaxf((k.z + u.z) + v.z, 0.f);
  y.w = fmaxf((k.w + u.w) + v.w, 0.f);
  *reinterpret_cast<float4*>(a1 + idx * 4) = y;
}

// delta8[r,j] = (sum_c dl[r,c] * w_out[c,j]) * (a8[r,j] > 0);  dl = coef * (c + 2*sq0*d0*e0)
// (deltas_to_objective, generator/diffusion.py:430-471, differentiated)
__global__ void seed_kernel(float* __restrict__ d8, const float* __restrict__ a8, const float* __restrict__ lg,
                            const float* __restrict__ w_out, dgdm_objective obj, int64_t r0, int64_t rows) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * W256) return;
  int64_t r = idx / W256;
  int j = (int)(idx % W256);
  float coef = obj.row_coef ? obj.row_coef[r0 + r] : 1.f;
  float d0 = lg[r * 3];
  float s0 = coef * (obj.c[0] + 2.f * obj.sq0 * d0), s1 = coef * obj.c[1], s2 = coef * obj.c[2];
  float v = s0 * w_out[j] + s1 * w_out[W256 + j] + s2 * w_out[2 * W256 + j];
  d8[idx] = a8[idx] > 0.f ? v : 0.f;
}

// K2 (SIMT path): dUp[p,:] += sum over this chunk's rows of pair p of delta1[r,:].
// One CTA per (pair touched by the chunk), threads over columns, rows in ascending order => deterministic.
__global__ void colsum_kernel(float* __restrict__ dUp, const float* __restrict__ d1, int64_t r0, int64_t rows, int G,
                              int H1) {
  int64_t p = r0 / G + blockIdx.x;
  int64_t lo = p * G > r0 ? p * G : r0;
  int64_t hi = (p + 1) * G < r0 + rows ? (p + 1) * G : r0 + rows;
  for (int c = threadIdx.x; c < H1; c += blockDim.x) {
    float s = 0.f;
    for (int64_t r = lo; r < hi; ++r) s += d1[(r - r0) * H1 + c];
    dUp[p * H1 + c] += s;
  }
}

// score_sum[p] += sum over this chunk's rows of pair p of objective(logits[r])
__global__ void objective_sum_kernel(float* __restrict__ score_sum, const float* __restrict__ lg, dgdm_objective obj,
                                     int64_t r0, int64_t rows, int G) {
  int64_t p = r0 / G + blockIdx.x;
  int64_t lo = p * G > r0 ? p * G : r0;
  int64_t hi = (p + 1) * G < r0 + rows ? (p + 1) * G : r0 + rows;
  __shared__ float part[256];
  float s = 0.f;
  for (int64_t r = lo + threadIdx.x; r < hi; r += blockDim.x) {
    const float* l = lg + (r - r0) * 3;
    float coef = obj.row_coef ? obj.row_coef[r] : 1.f;
    s += coef * (obj.c[0] * l[0] + obj.c[1] * l[1] + obj.c[2] * l[2] + obj.sq0 * l[0] * l[0]);
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) score_sum[p] += part[0];
}

// dU[d,:] = mul * sum_j dUp[d*opd + j, :]   (fixed order)
__global__ void fold_pairs_kernel(float* __restrict__ dU, const float* __restrict__ dUp, int64_t n, int H1, int opd,
                                  float mul) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * H1) return;
  int64_t d = idx / H1;
  int c = (int)(idx % H1);
  float s = 0.f;
  for (int j = 0; j < opd; ++j) s += dUp[(d * opd + j) * H1 + c];
  dU[idx] = s * mul;
}

__global__ void scale_kernel(float* __restrict__ out, const float* __restrict__ in, int64_t n, float mul) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * mul;
}

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// A linear layer at row scale (explicit-row mode, where nothing can be hoisted).  In the tensor-core modes it runs on
// tcgen05 in fp32-grade arithmetic (bf16 hi/lo operand split, 3 MMAs per product) whenever the shape allows: the
// weight is packed into `scratch` on the fly (a 256x256 image is 256 KB; one stream => the buffer can be reused by
// the next layer).  N = 512 runs as two 256-column halves.  Everything else stays on the CUDA-core GEMM.
struct RowLinear {
  bool tc; uint8_t* scratch; int* err; cudaStream_t s;
  int operator()(GemmArgs g) const {
    if (tc && g.N % 256 == 0 && g.N > 256) {
      for (int n0 = 0; n0 < g.N; n0 += 256) {
        GemmArgs h = g;
        h.W = g.W + (int64_t)n0 * g.K; h.bias = g.bias ? g.bias + n0 : nullptr; h.C = g.C + n0; h.N = 256;
        if (g.mask) h.mask = g.mask + n0;
        DGDM_TRY((*this)(h));
      }
      return DGDM_OK;
    }
    if (tc && gemm_tc_eligible(g) && gemm_tc_image_bytes(g.N, g.K) <= ROW_SCRATCH_BYTES) {
      DGDM_TRY(gemm_tc_pack(g.W, g.N, g.K, scratch, s));
      return gemm_tc(g, scratch, 1, err, s);
    }
    return gemm_f32(g, s);
  }
  static constexpr size_t ROW_SCRATCH_BYTES = 512 * 1024;     // K = 512, N = 256
};

struct Hoist {
  float *h0, *e, *U, *Cst, *V, *Vt, *pose, *temb, *th, *te, *tc, *oh, *oc, *dUp, *dU, *de, *dh0, *score_sum;
  uint8_t* lin_scratch; int* lin_err;     // packed-weight scratch + error word of the tcgen05 linear kernel (RowLinear)
  int G;
  int64_t n_pairs;
};

constexpr int64_t SIMT_CHUNK = 32768;

size_t hoist_bytes(int P, int H1, int obj_dim, int64_t nd, int n_obj, int64_t n_pairs, int G) {
  auto a = [](size_t n) { return align_up(n * sizeof(float), 256); };
  size_t b = 0;
  b += a(nd * 256) * 2 + a(nd * H1);                     // h0, e, U
  b += a((size_t)n_obj * H1) + 2 * a((size_t)G * H1) + a((size_t)G * 27);
  b += a(256) * 3 + a(H1);                               // temb, th, te, tc
  b += a((size_t)n_obj * 256) * 2;                       // oh, oc
  b += a(n_pairs * H1) + a(nd * H1) + a(nd * 256) * 2;   // dUp, dU, de, dh0
  b += a(n_pairs);                                       // score_sum
  b += align_up(RowLinear::ROW_SCRATCH_BYTES, 256) + 256; // lin_scratch, lin_err
  (void)P; (void)obj_dim;
  return b;
}

size_t simt_bytes(int H1, int64_t n_rows) {
  int64_t rc = n_rows < SIMT_CHUNK ? n_rows : SIMT_CHUNK;
  auto a = [](size_t n) { return align_up(n * sizeof(float), 256); };
  return a(rc * H1) * 3 + a(rc * 256) * 7 + a(rc * 3);
}

int check_common(const dgdm_dyn_weights* w, const float* x, int n_designs, const float* objects, int n_obj, int opd,
                 const int32_t* pair_object, const dgdm_pose_grid* grid, const dgdm_objective* obj) {
  DGDM_CHECK_ARG(w && x && objects && grid && obj, "dynamics: null pointer");
  DGDM_CHECK_ARG(w->H1 == 256 || w->H1 == 512, "dynamics: H1=%d unsupported (256 or 512)", w->H1);
  DGDM_CHECK_ARG(w->P >= 1 && w->P <= 4096, "dynamics: P=%d", w->P);
  DGDM_CHECK_ARG(n_designs >= 1 && n_obj >= 1 && opd >= 1, "dynamics: n_designs=%d n_obj=%d objs_per_design=%d",
                 n_designs, n_obj, opd);
  DGDM_CHECK_ARG(grid->grid_size >= 1 && (grid->pos_zero || grid->num_pos >= 1), "dynamics: empty pose grid");
  if (!pair_object) {
    if (opd == 1) DGDM_CHECK_ARG(n_designs % n_obj == 0, "dynamics: n_designs %% n_obj != 0 with implicit pairing");
    else DGDM_CHECK_ARG(opd == n_obj, "dynamics: objs_per_design must equal n_obj with implicit pairing");
  }
  return DGDM_OK;
}

// Everything that does not scale with B*G.
int run_hoists(const dgdm_dyn_weights* w, const float* x, int nd, const float* objects, int n_obj, float t_frac,
               const dgdm_pose_grid* grid, Hoist& h, int precision, cudaStream_t s) {
  const int H1 = w->H1, P = w->P;
  // The design-sized GEMMs (K, N multiples of 64 / 256) run on the tcgen05 linear kernel in fp32-grade arithmetic in
  // every tensor-core mode (they were 1-2 % of a pass on CUDA cores); fp32_simt keeps exact fp32 throughout.
  const bool tc = precision != DGDM_PREC_FP32_SIMT;
  if (tc) DGDM_CUDA(cudaMemsetAsync(h.lin_err, 0, sizeof(int), s));
  const RowLinear lin{tc, h.lin_scratch, h.lin_err, s};
  // gripper encoder: 14|42 -> 256 -> 256 (ReLU between), then its layer-1 block
  DGDM_TRY(gemm_f32(gemm_plain(x, P, w->ge_w0, w->ge_b0, h.h0, 256, nd, 256, P, ACT_RELU), s));
  DGDM_TRY(lin(gemm_plain(h.h0, 256, w->ge_w1, w->ge_b1, h.e, 256, nd, 256, 256, ACT_NONE)));
  DGDM_TRY(lin(gemm_plain(h.e, 256, w->w1_ctrl, nullptr, h.U, H1, nd, H1, 256, ACT_NONE)));
  // pose table
  pose_embed_kernel<<<blocks_for(h.G, 128), 128, 0, s>>>(h.pose, *grid, h.G);
  DGDM_LAUNCH_CHECK();
  DGDM_TRY(gemm_f32(gemm_plain(h.pose, 27, w->w1_pose, nullptr, h.V, H1, h.G, H1, 27, ACT_NONE), s));
  // the same table as [H1/4][G][4] for the tensor-core trunk's coalesced 128-bit per-row reads
  pose_table_quads_kernel<<<blocks_for((int64_t)h.G * (H1 / 4), 256), 256, 0, s>>>(reinterpret_cast<float4*>(h.Vt), h.V, h.G, H1);
  DGDM_LAUNCH_CHECK();
  // time: 2D = MLP(SiLU) of a 128-d embedding; 3D = raw 256-d embedding (profile_forward_3d.py:83)
  const float* te;
  if (w->is_3d) {
    time_embed_kernel<<<1, 128, 0, s>>>(h.temb, t_frac, 256);
    DGDM_LAUNCH_CHECK();
    te = h.temb;
  } else {
    time_embed_kernel<<<1, 64, 0, s>>>(h.temb, t_frac, 128);
    DGDM_LAUNCH_CHECK();
    DGDM_TRY(gemm_f32(gemm_plain(h.temb, 128, w->te_w0, w->te_b0, h.th, 256, 1, 256, 128, ACT_SILU), s));
    DGDM_TRY(gemm_f32(gemm_plain(h.th, 256, w->te_w1, w->te_b1, h.te, 256, 1, 256, 256, ACT_NONE), s));
    te = h.te;
  }
  DGDM_TRY(gemm_f32(gemm_plain(te, 256, w->w1_time, w->b1, h.tc, H1, 1, H1, 256, ACT_NONE), s));
  // object code: 2D = MLP(ReLU) of the flattened contour; 3D = PointNet++ code computed once per object (K5)
  const float* oc;
  if (w->is_3d) {
    oc = objects;
  } else {
    DGDM_TRY(gemm_f32(gemm_plain(objects, w->obj_dim, w->oe_w0, w->oe_b0, h.oh, 256, n_obj, 256, w->obj_dim, ACT_RELU), s));
    DGDM_TRY(gemm_f32(gemm_plain(h.oh, 256, w->oe_w1, w->oe_b1, h.oc, 256, n_obj, 256, 256, ACT_NONE), s));
    oc = h.oc;
  }
  DGDM_TRY(gemm_f32(gemm_plain(oc, 256, w->w1_obj, h.tc, h.Cst, H1, n_obj, H1, 256, ACT_NONE), s));
  return DGDM_OK;
}

int carve(const dgdm_dyn_weights* w, int nd, int n_obj, int opd, const dgdm_pose_grid* grid, Arena& ar, Hoist& h) {
  const int H1 = w->H1;
  h.G = grid->pos_zero ? grid->grid_size : grid->grid_size * grid->num_pos * grid->num_pos;
  h.n_pairs = (int64_t)nd * opd;
  h.h0 = ar.take<float>((size_t)nd * 256);
  h.e = ar.take<float>((size_t)nd * 256);
  h.U = ar.take<float>((size_t)nd * H1);
  h.Cst = ar.take<float>((size_t)n_obj * H1);
  h.V = ar.take<float>((size_t)h.G * H1);
  h.Vt = ar.take<float>((size_t)h.G * H1);
  h.pose = ar.take<float>((size_t)h.G * 27);
  h.temb = ar.take<float>(256);
  h.th = ar.take<float>(256);
  h.te = ar.take<float>(256);
  h.tc = ar.take<float>(H1);
  h.oh = ar.take<float>((size_t)n_obj * 256);
  h.oc = ar.take<float>((size_t)n_obj * 256);
  h.dUp = ar.take<float>((size_t)h.n_pairs * H1);
  h.dU = ar.take<float>((size_t)nd * H1);
  h.de = ar.take<float>((size_t)nd * 256);
  h.dh0 = ar.take<float>((size_t)nd * 256);
  h.score_sum = ar.take<float>((size_t)h.n_pairs);
  h.lin_scratch = ar.take<uint8_t>(RowLinear::ROW_SCRATCH_BYTES);
  h.lin_err = ar.take<int>(1);
  return DGDM_OK;
}

// Exact-fp32 trunk over all rows, chunked so activations stay bounded.
int simt_trunk(const dgdm_dyn_weights* w, const Hoist& h, int nd, int n_obj, int opd, const int32_t* pair_object,
               const dgdm_objective* obj, bool backward, float* logits_out, Arena& ar, cudaStream_t s) {
  const int H1 = w->H1, G = h.G;
  const int64_t n_rows = h.n_pairs * G;
  const int64_t rc = n_rows < SIMT_CHUNK ? n_rows : SIMT_CHUNK;
  float* a1 = ar.take<float>((size_t)rc * H1);
  float* dA = ar.take<float>((size_t)rc * H1);
  float* dB = ar.take<float>((size_t)rc * H1);
  float* act[7];
  for (int l = 0; l < 7; ++l) act[l] = ar.take<float>((size_t)rc * 256);
  float* lg_scratch = ar.take<float>((size_t)rc * 3);
  if (!ar.ok) { set_error("dynamics: workspace too small (need >= %zu bytes)", ar.off); return DGDM_EWORKSPACE; }

  if (backward) DGDM_CUDA(cudaMemsetAsync(h.dUp, 0, (size_t)h.n_pairs * H1 * sizeof(float), s));
  else DGDM_CUDA(cudaMemsetAsync(h.score_sum, 0, (size_t)h.n_pairs * sizeof(float), s));

  for (int64_t r0 = 0; r0 < n_rows; r0 += rc) {
    const int64_t rows = n_rows - r0 < rc ? n_rows - r0 : rc;
    build_a1_kernel<<<blocks_for(rows * (H1 / 4), 256), 256, 0, s>>>(a1, h.U, h.Cst, h.V, r0, rows, G, H1, opd, nd,
                                                                    n_obj, pair_object);
    DGDM_LAUNCH_CHECK();
    const float* prev = a1;
    int prev_w = H1;
    for (int l = 0; l < 7; ++l) {      // trunk layers 2..8 (BN folded) + ReLU
      DGDM_TRY(gemm_f32(gemm_plain(prev, prev_w, w->wl[l], w->bl[l], act[l], 256, rows, 256, prev_w, ACT_RELU), s));
      prev = act[l];
      prev_w = 256;
    }
    float* lg = logits_out ? logits_out + r0 * 3 : lg_scratch;
    DGDM_TRY(gemm_f32(gemm_plain(act[6], 256, w->w_out, w->b_out, lg, 3, rows, 3, 256, ACT_NONE), s));
    const unsigned pairs_in_chunk = (unsigned)((r0 + rows - 1) / G - r0 / G + 1);
    if (!backward) {
      objective_sum_kernel<<<pairs_in_chunk, 256, 0, s>>>(h.score_sum, lg, *obj, r0, rows, G);
      DGDM_LAUNCH_CHECK();
      continue;
    }
    seed_kernel<<<blocks_for(rows * 256, 256), 256, 0, s>>>(dA, act[6], lg, w->w_out, *obj, r0, rows);
    DGDM_LAUNCH_CHECK();
    float* cur = dA;
    float* nxt = dB;
    for (int l = 6; l >= 0; --l) {     // delta_{l+1} = (delta_{l+2} . W_{l+2}) * 1[a_{l+1} > 0]
      const int kw = (l == 0) ? H1 : 256;
      GemmArgs g = gemm_plain(cur, 256, w->wl_t[l], nullptr, nxt, kw, rows, kw, 256, ACT_NONE);
      g.mask = (l == 0) ? a1 : act[l - 1];
      g.m_rs = kw;
      DGDM_TRY(gemm_f32(g, s));
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    colsum_kernel<<<pairs_in_chunk, 256, 0, s>>>(h.dUp, cur, r0, rows, G, H1);
    DGDM_LAUNCH_CHECK();
  }
  return DGDM_OK;
}

int trunk_dispatch(const dgdm_dyn_weights* w, const Hoist& h, int nd, int n_obj, int opd, const int32_t* pair_object,
                   const dgdm_objective* obj, bool backward, float* logits, Arena& ar, int precision, cudaStream_t s) {
  if (precision == DGDM_PREC_FP32_SIMT)
    return simt_trunk(w, h, nd, n_obj, opd, pair_object, obj, backward, logits, ar, s);
  DGDM_CHECK_ARG(precision == DGDM_PREC_BF16X3 || precision == DGDM_PREC_BF16 || precision == DGDM_PREC_FP16 ||
                     precision == DGDM_PREC_FP16X3,
                 "dynamics: unknown precision %d", precision);
  DGDM_CHECK_ARG(w->tc_image != nullptr, "dynamics: tensor-core precision requested but weights carry no tc_image "
                                         "(call dgdm_dyn_pack_tc)");
  size_t need = tc_trunk_workspace_bytes(w->H1, h.n_pairs, h.G);
  void* tws = ar.take<char>(need);
  if (!ar.ok) { set_error("dynamics: workspace too small (need >= %zu bytes)", ar.off); return DGDM_EWORKSPACE; }
  return tc_trunk(w, h.U, h.Cst, h.Vt, nd, n_obj, opd, pair_object, h.G, obj, backward, h.dUp, h.score_sum, logits, tws,
                  need, precision, s);
}

}  // namespace
}  // namespace dgdm

extern "C" size_t dgdm_dyn_rows_workspace_bytes(const dgdm_dyn_weights* w, int64_t n_rows, int32_t precision) {
  using namespace dgdm;
  if (!w || n_rows < 1) return 0;
  auto a = [](size_t n) { return align_up(n * sizeof(float), 256); };
  const int H1 = w->H1;
  size_t b = a(n_rows * 256) * 8 + a((size_t)n_rows * H1) * 7 + a(n_rows * 27) + a(2 * (size_t)H1) + a(n_rows);
  if (precision == DGDM_PREC_FP32_SIMT) b += simt_bytes(H1, n_rows);
  else b += align_up(tc_trunk_workspace_bytes(H1, n_rows, 1), 256) + RowLinear::ROW_SCRATCH_BYTES + 256;
  return b + 4096;
}

extern "C" int dgdm_dyn_forward_rows(const dgdm_dyn_weights* w, const float* x, const float* ori, const float* pos,
                                     const float* t_frac, const float* objects, int64_t n_rows,
                                     const dgdm_objective* objective, float* logits, float* grad_x, void* workspace,
                                     size_t workspace_bytes, int32_t precision, void* stream) {
  using namespace dgdm;
  DGDM_CHECK_ARG(w && x && ori && pos && t_frac && objects && workspace, "dgdm_dyn_forward_rows: null pointer");
  DGDM_CHECK_ARG(logits || grad_x, "dgdm_dyn_forward_rows: nothing to compute (logits and grad_x both NULL)");
  DGDM_CHECK_ARG(!grad_x || objective, "dgdm_dyn_forward_rows: grad_x needs an objective");
  DGDM_CHECK_ARG(n_rows >= 1 && n_rows < (1ll << 31), "dgdm_dyn_forward_rows: n_rows out of range");
  DGDM_CHECK_ARG(w->H1 == 256 || w->H1 == 512, "dgdm_dyn_forward_rows: H1=%d unsupported", w->H1);
  cudaStream_t s = (cudaStream_t)stream;
  const int H1 = w->H1, P = w->P;
  const int64_t n = n_rows;
  Arena ar(workspace, workspace_bytes);
  Hoist h{};
  h.G = 1; h.n_pairs = n;
  h.h0 = ar.take<float>(n * 256);
  h.e = ar.take<float>(n * 256);
  float* temb = ar.take<float>(n * 256);
  float* th = ar.take<float>(n * 256);
  float* te = ar.take<float>(n * 256);
  h.oh = ar.take<float>(n * 256);
  h.oc = ar.take<float>(n * 256);
  h.de = ar.take<float>(n * 256);
  h.dh0 = h.e;                                   // e is dead once U is formed
  float* Ur = ar.take<float>((size_t)n * H1);
  float* Vr = ar.take<float>((size_t)n * H1);
  float* Tr = ar.take<float>((size_t)n * H1);
  float* Or = ar.take<float>((size_t)n * H1);
  h.U = ar.take<float>((size_t)n * H1);          // base = Or + Ur + Vr + Tr
  h.dUp = ar.take<float>((size_t)n * H1);
  h.dU = ar.take<float>((size_t)n * H1);
  h.pose = ar.take<float>(n * 27);
  float* zeros = ar.take<float>(2 * (size_t)H1);
  h.score_sum = ar.take<float>(n);
  const bool tc = precision != DGDM_PREC_FP32_SIMT;
  uint8_t* scratch = tc ? ar.take<uint8_t>(RowLinear::ROW_SCRATCH_BYTES) : nullptr;
  int* lin_err = tc ? ar.take<int>(1) : nullptr;
  if (!ar.ok) { set_error("dgdm_dyn_forward_rows: workspace too small (need >= %zu bytes)", ar.off); return DGDM_EWORKSPACE; }
  h.Cst = zeros; h.V = zeros + H1; h.Vt = zeros + H1;
  DGDM_CUDA(cudaMemsetAsync(zeros, 0, 2 * (size_t)H1 * sizeof(float), s));
  if (tc) DGDM_CUDA(cudaMemsetAsync(lin_err, 0, sizeof(int), s));
  const RowLinear lin{tc, scratch, lin_err, s};
  // every encoder at row scale -- this is the un-hoistable "paired" mode of the forward signature
  DGDM_TRY(lin(gemm_plain(x, P, w->ge_w0, w->ge_b0, h.h0, 256, n, 256, P, ACT_RELU)));
  DGDM_TRY(lin(gemm_plain(h.h0, 256, w->ge_w1, w->ge_b1, h.e, 256, n, 256, 256, ACT_NONE)));
  DGDM_TRY(lin(gemm_plain(h.e, 256, w->w1_ctrl, nullptr, Ur, H1, n, H1, 256, ACT_NONE)));
  pose_embed_rows_kernel<<<blocks_for(n, 128), 128, 0, s>>>(h.pose, ori, pos, n);
  DGDM_LAUNCH_CHECK();
  DGDM_TRY(lin(gemm_plain(h.pose, 27, w->w1_pose, nullptr, Vr, H1, n, H1, 27, ACT_NONE)));
  const float* tep;
  if (w->is_3d) {
    time_embed_rows_kernel<<<blocks_for(n * 128, 256), 256, 0, s>>>(temb, t_frac, n, 256);
    DGDM_LAUNCH_CHECK();
    tep = temb;
  } else {
    time_embed_rows_kernel<<<blocks_for(n * 64, 256), 256, 0, s>>>(temb, t_frac, n, 128);
    DGDM_LAUNCH_CHECK();
    DGDM_TRY(lin(gemm_plain(temb, 128, w->te_w0, w->te_b0, th, 256, n, 256, 128, ACT_SILU)));
    DGDM_TRY(lin(gemm_plain(th, 256, w->te_w1, w->te_b1, te, 256, n, 256, 256, ACT_NONE)));
    tep = te;
  }
  DGDM_TRY(lin(gemm_plain(tep, 256, w->w1_time, w->b1, Tr, H1, n, H1, 256, ACT_NONE)));
  const float* oc = objects;
  if (!w->is_3d) {
    DGDM_TRY(lin(gemm_plain(objects, w->obj_dim, w->oe_w0, w->oe_b0, h.oh, 256, n, 256, w->obj_dim, ACT_RELU)));
    DGDM_TRY(lin(gemm_plain(h.oh, 256, w->oe_w1, w->oe_b1, h.oc, 256, n, 256, 256, ACT_NONE)));
    oc = h.oc;
  }
  DGDM_TRY(lin(gemm_plain(oc, 256, w->w1_obj, nullptr, Or, H1, n, H1, 256, ACT_NONE)));
  sum4_kernel<<<blocks_for(n * H1 / 4, 256), 256, 0, s>>>(h.U, Or, Ur, Vr, Tr, n * H1 / 4);
  DGDM_LAUNCH_CHECK();
  dgdm_objective dummy{{0.f, 0.f, 0.f}, 0.f, nullptr};
  const dgdm_objective* ob = objective ? objective : &dummy;
  if (!grad_x) {
    DGDM_TRY(trunk_dispatch(w, h, (int)n, 1, 1, nullptr, ob, false, logits, ar, precision, s));
    return DGDM_OK;
  }
  DGDM_TRY(trunk_dispatch(w, h, (int)n, 1, 1, nullptr, ob, true, logits, ar, precision, s));
  DGDM_TRY(lin(gemm_plain(h.dUp, H1, w->w1_ctrl_t, nullptr, h.de, 256, n, 256, H1, ACT_NONE)));
  GemmArgs g = gemm_plain(h.de, 256, w->ge_w1_t, nullptr, h.dh0, 256, n, 256, 256, ACT_NONE);
  g.mask = h.h0;
  DGDM_TRY(lin(g));
  DGDM_TRY(lin(gemm_plain(h.dh0, 256, w->ge_w0_t, nullptr, grad_x, P, n, P, 256, ACT_NONE)));
  return DGDM_OK;
}

extern "C" size_t dgdm_dyn_guidance_workspace_bytes(const dgdm_dyn_weights* w, int32_t n_designs, int32_t n_obj,
                                                    int32_t objs_per_design, const dgdm_pose_grid* grid,
                                                    int32_t precision) {
  using namespace dgdm;
  if (!w || !grid || n_designs < 1 || n_obj < 1 || objs_per_design < 1) return 0;
  int G = grid->pos_zero ? grid->grid_size : grid->grid_size * grid->num_pos * grid->num_pos;
  int64_t n_pairs = (int64_t)n_designs * objs_per_design;
  size_t b = hoist_bytes(w->P, w->H1, w->obj_dim, n_designs, n_obj, n_pairs, G);
  if (precision == DGDM_PREC_FP32_SIMT) b += simt_bytes(w->H1, n_pairs * G);
  else b += align_up(tc_trunk_workspace_bytes(w->H1, n_pairs, G), 256);
  return b + 4096;
}

extern "C" int dgdm_dyn_guidance(const dgdm_dyn_weights* w, const float* x, int32_t n_designs, const float* objects,
                                 int32_t n_obj, int32_t objs_per_design, const int32_t* pair_object, float t_frac,
                                 const dgdm_pose_grid* grid, const dgdm_objective* objective, float grad_mul,
                                 float* grad, float* logits, void* workspace, size_t workspace_bytes,
                                 int32_t precision, void* stream) {
  using namespace dgdm;
  DGDM_TRY(check_common(w, x, n_designs, objects, n_obj, objs_per_design, pair_object, grid, objective));
  DGDM_CHECK_ARG(grad && workspace, "dgdm_dyn_guidance: null output/workspace");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(workspace, workspace_bytes);
  Hoist h{};
  carve(w, n_designs, n_obj, objs_per_design, grid, ar, h);
  if (!ar.ok) { set_error("dgdm_dyn_guidance: workspace too small (need >= %zu bytes)", ar.off); return DGDM_EWORKSPACE; }
  DGDM_TRY(run_hoists(w, x, n_designs, objects, n_obj, t_frac, grid, h, precision, s));
  DGDM_TRY(trunk_dispatch(w, h, n_designs, n_obj, objs_per_design, pair_object, objective, true, logits, ar, precision, s));
  const int H1 = w->H1, P = w->P;
  // K2 tail: fold pairs into designs, then the gripper-encoder backward (B-sized GEMMs)
  fold_pairs_kernel<<<blocks_for((int64_t)n_designs * H1, 256), 256, 0, s>>>(h.dU, h.dUp, n_designs, H1, objs_per_design, grad_mul);
  DGDM_LAUNCH_CHECK();
  const RowLinear lin{precision != DGDM_PREC_FP32_SIMT, h.lin_scratch, h.lin_err, s};
  DGDM_TRY(lin(gemm_plain(h.dU, H1, w->w1_ctrl_t, nullptr, h.de, 256, n_designs, 256, H1, ACT_NONE)));
  GemmArgs g = gemm_plain(h.de, 256, w->ge_w1_t, nullptr, h.dh0, 256, n_designs, 256, 256, ACT_NONE);
  g.mask = h.h0;
  DGDM_TRY(lin(g));
  DGDM_TRY(gemm_f32(gemm_plain(h.dh0, 256, w->ge_w0_t, nullptr, grad, P, n_designs, P, 256, ACT_NONE), s));
  return DGDM_OK;
}

extern "C" int dgdm_dyn_score(const dgdm_dyn_weights* w, const float* x, int32_t n_designs, const float* objects,
                              int32_t n_obj, int32_t objs_per_design, const int32_t* pair_object, float t_frac,
                              const dgdm_pose_grid* grid, const dgdm_objective* objective, float* scores,
                              float* logits, void* workspace, size_t workspace_bytes, int32_t precision,
                              void* stream) {
  using namespace dgdm;
  DGDM_TRY(check_common(w, x, n_designs, objects, n_obj, objs_per_design, pair_object, grid, objective));
  DGDM_CHECK_ARG(scores && workspace, "dgdm_dyn_score: null output/workspace");
  cudaStream_t s = (cudaStream_t)stream;
  Arena ar(workspace, workspace_bytes);
  Hoist h{};
  carve(w, n_designs, n_obj, objs_per_design, grid, ar, h);
  if (!ar.ok) { set_error("dgdm_dyn_score: workspace too small (need >= %zu bytes)", ar.off); return DGDM_EWORKSPACE; }
  DGDM_TRY(run_hoists(w, x, n_designs, objects, n_obj, t_frac, grid, h, precision, s));
  DGDM_TRY(trunk_dispatch(w, h, n_designs, n_obj, objs_per_design, pair_object, objective, false, logits, ar, precision, s));
  scale_kernel<<<blocks_for(h.n_pairs, 256), 256, 0, s>>>(scores, h.score_sum, h.n_pairs, 1.f / (float)h.G);
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}
