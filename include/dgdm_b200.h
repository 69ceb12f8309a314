/*
 * dgdm_b200.h -- C ABI of the B200-native dynamics-guided diffusion sampling path.
 *
 * The reference (real-stanford/dgdm) is pure Python and exposes no FFI; its boundary for this path is
 * a set of Python methods.  Each entry point below states the reference code it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer
 * of the reference would add to call them from generator/diffusion.py.
 *
 * Conventions (every function):
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless named host_*;
 *     fp32, row-major, densely packed unless a stride is given;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - no device allocation: scratch comes from the caller, sized by the matching *_workspace_bytes();
 *   - returns 0 on success, a negative DGDM_E* code otherwise; dgdm_last_error() gives the message
 *     (thread-local); nothing throws;
 *   - one host thread per device; no global mutable state besides the error string.
 */
#ifndef DGDM_B200_H_
#define DGDM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGDM_ABI_VERSION 1

enum {
  DGDM_OK = 0,
  DGDM_EINVAL = -1,     /* bad argument (shape, null pointer, unsupported size) */
  DGDM_EWORKSPACE = -2, /* workspace too small */
  DGDM_ECUDA = -3,      /* CUDA runtime error, message carries cudaGetErrorString */
  DGDM_EUNSUPPORTED = -4
};

/* Arithmetic mode of the dynamics-network trunk (layers 2..8 forward + input-gradient backward). */
enum {
  DGDM_PREC_FP32_SIMT = 0, /* fp32 FFMA on CUDA cores: exact-order reference path on the GPU            */
  DGDM_PREC_BF16X3 = 1,    /* tcgen05 kind::f16, operands split hi+lo bf16, 3 MMAs: fp32-grade (<=1e-3)  */
  DGDM_PREC_BF16 = 2,      /* tcgen05 kind::f16, single bf16 pass (<=2e-2)                               */
  DGDM_PREC_FP16 = 3,      /* tcgen05 kind::f16, single fp16 pass in the trunk (3 more mantissa bits than bf16, same
                            * tensor rate); the denoiser runs in BF16X3                                          */
  DGDM_PREC_FP16X3 = 4     /* as BF16X3 with fp16 hi+lo operand parts in the trunk (22 operand bits instead of 16, same
                            * 3 MMAs per product); the denoiser runs in BF16X3.  Both fp16 modes keep operands scaled
                            * by powers of two inside the kernel (weights x256, activations x64, gradients x64) and
                            * saturate: |BN-folded weight| < 255 and |activation| < 1023 are assumed               */
};

const char* dgdm_last_error(void);
int dgdm_abi_version(void);
/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
uint64_t dgdm_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K4  fused guided DDIM update.
 * Replaces generator/diffusion.py:575-576 (and :645-647) + diffusers DDIMScheduler.step (eta = 0,
 * epsilon prediction, clip_sample):
 *     eps_hat = eps - (sqrt_1m_at * grad) * scale
 *     x0      = clamp((x - sqrt_1m_at * eps_hat) / sqrt_at, -1, 1)        (clamp iff clip != 0)
 *     x_prev  = sqrt_aprev * x0 + sqrt_1m_aprev * eps_hat
 * with the same operation order and no FMA contraction, so it is bit-exact against the CPU.
 * x_prev may alias x.  grad may be NULL (unguided sampling, generator/diffusion.py:249-256).
 * 16 algorithmic bytes per element (read x, eps, grad; write x_prev).
 */
int dgdm_ddim_guided_update(float* x_prev, const float* x, const float* eps, const float* grad, int64_t n,
                            float sqrt_1m_at, float sqrt_at, float sqrt_aprev, float sqrt_1m_aprev,
                            float scale, int clip, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dynamics network (ProfileForward2DModel / ProfileForward3DModel), pre-packed.
 * Built on the host by dgdm_b200.pack from a reference checkpoint (dynamics/trainer.py:105-106 format):
 * BatchNorm1d (eval) folded into the preceding Linear; layer 1 split by input block
 * [object | gripper | pose | time] (profile_forward_2d.py:154 concatenation order); transposed copies
 * for the input-gradient GEMMs.
 */
typedef struct dgdm_dyn_weights {
  int32_t P;       /* control points fed to the gripper encoder: 14 (2D) or 42 (3D)                 */
  int32_t H1;      /* width of trunk layer 1: 256 (2D) or 512 (3D)                                  */
  int32_t obj_dim; /* 2D: 2*V raw contour floats into the object MLP; 3D: 256 (PointNet++ code)     */
  int32_t is_3d;   /* 3D: object code used as is, raw 256-d time embedding (time_encoder unused)    */
  const float* ge_w0; const float* ge_b0;   /* gripper_encoder.0  [256,P],[256]                     */
  const float* ge_w1; const float* ge_b1;   /* gripper_encoder.2  [256,256],[256]                   */
  const float* oe_w0; const float* oe_b0;   /* object_encoder.0   [256,obj_dim] (2D only)           */
  const float* oe_w1; const float* oe_b1;   /* object_encoder.2   [256,256]     (2D only)           */
  const float* te_w0; const float* te_b0;   /* time_encoder.0     [256,128]     (2D only)           */
  const float* te_w1; const float* te_b1;   /* time_encoder.2     [256,256]     (2D only)           */
  const float* w1_obj;  /* [H1,256]  BN-folded layer-1 block acting on the object code              */
  const float* w1_ctrl; /* [H1,256]  ... on the gripper code                                        */
  const float* w1_ctrl_t; /* [256,H1] transpose of w1_ctrl (encoder backward)                       */
  const float* w1_pose; /* [H1,27]   ... on the Fourier pose embedding                              */
  const float* w1_time; /* [H1,256]  ... on the time embedding                                      */
  const float* b1;      /* [H1]                                                                     */
  const float* ge_w1_t; /* [256,256] transpose of ge_w1 (encoder backward)                          */
  const float* ge_w0_t; /* [P,256]   transpose of ge_w0 (encoder backward)                          */
  const float* wl[7];   /* trunk layers 2..8, BN folded, [256,K_l] with K_2 = H1, else 256          */
  const float* wl_t[7]; /* transposes [K_l,256]                                                     */
  const float* bl[7];   /* [256]                                                                    */
  const float* w_out;   /* [3,256]                                                                  */
  const float* b_out;   /* [3]                                                                      */
  const void* tc_image; /* tensor-core weight image written by dgdm_dyn_pack_tc, or NULL            */
} dgdm_dyn_weights;

/* Pose grid of cond_fn (generator/diffusion.py:478-482): g = (o*num_pos + ix)*num_pos + iy with
 * ori = linspace(ori_lo, ori_hi, grid_size)[o], pos = linspace(-1,1,num_pos)[ix|iy].
 * pos_zero != 0 selects the profile-pass convention of get_convergence_centers (:509-511):
 * G = grid_size rows, pos = (0,0). */
typedef struct dgdm_pose_grid {
  float ori_lo, ori_hi;
  int32_t grid_size, num_pos, pos_zero;
} dgdm_pose_grid;

/* Objective of deltas_to_objective (generator/diffusion.py:430-471) on logits d = (dθ,dx,dy)/std:
 *   value(row) = coef_row * ( c[0]*d0 + c[1]*d1 + c[2]*d2 + sq0 * d0^2 )
 * 'rotate' is {c=0, sq0=1}; the 14 linear objectives have sq0 = 0 and c in {-1,0,1}^3;
 * 'convergence' (:445-452) is c = (1,0,0) with a per-row sign table row_coef (device, [n_pairs*G],
 * pair-major) that the host derives from the convergence centres; row_coef == NULL means 1. */
typedef struct dgdm_objective {
  float c[3];
  float sq0;
  const float* row_coef;
} dgdm_objective;

/* Size in bytes of the tensor-core weight image for a network with layer-1 width H1. */
size_t dgdm_dyn_tc_image_bytes(int32_t H1);
/* Build the image (a bf16 hi/lo copy followed by an fp16 hi/lo copy; 64-wide K blocks, 128B-swizzled K-major
 * tiles ready for cp.async.bulk + tcgen05.mma) from the fp32 fields of *w into tc_image (device). */
int dgdm_dyn_pack_tc(const dgdm_dyn_weights* w, void* tc_image, void* stream);

size_t dgdm_dyn_guidance_workspace_bytes(const dgdm_dyn_weights* w, int32_t n_designs, int32_t n_obj,
                                         int32_t objs_per_design, const dgdm_pose_grid* grid, int32_t precision);

/* K1 + K2  guidance gradient.  Replaces Diffusion.cond_fn (generator/diffusion.py:473-504) for a whole
 * batch of (design, object) pairs at once:
 *     grad[d] = grad_mul * sum_{j < objs_per_design} d/dx_d  sum_{g < G} objective(f(x_d, pose_g, t, obj(d,j)))
 * pair p = d*objs_per_design + j uses object pair_object[p] (device int32; NULL => p / (n_designs/n_obj)
 * when objs_per_design == 1, i.e. designs are object-major blocks as in guided_sample (:561-570), or j when
 * objs_per_design == n_obj as in guided_sample_multi_object (:640-644) with grad_mul = 1/n_obj).
 *   x        [n_designs, P]
 *   objects  2D: [n_obj, obj_dim] flattened contours (x0,y0,x1,y1,..); 3D: [n_obj,256] PointNet++ codes
 *   t_frac   t / num_train_timesteps (generator/diffusion.py:487)
 *   grad     [n_designs, P]  (out)
 *   logits   optional out [n_pairs*G, 3], pair-major rows (p*G + g); NULL to skip
 */
int dgdm_dyn_guidance(const dgdm_dyn_weights* w, const float* x, int32_t n_designs, const float* objects,
                      int32_t n_obj, int32_t objs_per_design, const int32_t* pair_object, float t_frac,
                      const dgdm_pose_grid* grid, const dgdm_objective* objective, float grad_mul, float* grad,
                      float* logits, void* workspace, size_t workspace_bytes, int32_t precision, void* stream);

/* Forward-only pass: score[p] = mean_g objective(...) (SURVEY.md §8c "final predicted task score"; the
 * input convention of get_convergence_centers, generator/diffusion.py:509-516, when grid->pos_zero).
 * scores [n_pairs] (out); logits optional as above. */
int dgdm_dyn_score(const dgdm_dyn_weights* w, const float* x, int32_t n_designs, const float* objects, int32_t n_obj,
                   int32_t objs_per_design, const int32_t* pair_object, float t_frac, const dgdm_pose_grid* grid,
                   const dgdm_objective* objective, float* scores, float* logits, void* workspace,
                   size_t workspace_bytes, int32_t precision, void* stream);

/* Explicit rows -- the forward signature of the dynamics networks themselves,
 *   classifier_model(pts, ori, pos, t_float, object_vertices) -> (N,3)
 * (ProfileForward2DModel.forward, dynamics/profile_forward_2d.py:137-156; call sites generator/diffusion.py:487,496,
 * 500,516), where every row brings its own finger, pose, time and object, so nothing can be hoisted (BASELINE.json
 * configs[3], the "paired" throughput sweep).  x [N,P]; ori [N]; pos [N,2]; t_frac [N]; objects 2D [N,obj_dim] /
 * 3D [N,256] PointNet++ codes.  logits [N,3] (out, may be NULL).  If grad_x != NULL it receives
 * d/dx_r objective(logits_r), [N,P] (forward + input-gradient).  In the tensor-core modes the row-scale encoder
 * layers with K % 64 == 0 also run on tcgen05 (fp32-grade operand split, weights packed into the workspace per call). */
size_t dgdm_dyn_rows_workspace_bytes(const dgdm_dyn_weights* w, int64_t n_rows, int32_t precision);
int dgdm_dyn_forward_rows(const dgdm_dyn_weights* w, const float* x, const float* ori, const float* pos,
                          const float* t_frac, const float* objects, int64_t n_rows, const dgdm_objective* objective,
                          float* logits, float* grad_x, void* workspace, size_t workspace_bytes, int32_t precision,
                          void* stream);

/* Optional CUDA-event timing of the dominant kernel (the fused tensor-core trunk), used by bench.py for the
 * roofline line.  enable != 0 resets the counters and makes every subsequent launch of that kernel record an
 * event pair on its own stream; read() synchronises those events and returns the summed device time, the number
 * of launches and the number of guidance rows they processed. */
int dgdm_trunk_timing(int32_t enable);
int dgdm_trunk_timing_read(double* total_ms, int64_t* launches, int64_t* rows);
/* The same, split by launch kind: index 0 = guidance launches (forward + input-gradient), index 1 = scoring
 * launches (forward only).  ms, launches, rows each point at two elements.  Launches made while their stream is
 * being captured into a CUDA graph are not timed. */
int dgdm_trunk_timing_read_split(double* ms, int64_t* launches, int64_t* rows);

/* ------------------------------------------------------------------------------------------------
 * K3  denoiser.  Replaces ConditionalUnet1D.forward (generator/diffusion_utils.py:238-285) for the one
 * configuration the reference builds (generator/train.py:80): input_dim 1, down_dims [128,256],
 * step-embedding 32, kernel 5, 8 groups, no global_cond, the same timestep for every sample
 * (generator/diffusion.py:572).  Conv weights are pre-packed tap-major: [Cout][tap][Cin].
 */
typedef struct dgdm_unet_resblock {
  int32_t cin, cout;
  const float* conv0_w; const float* conv0_b; const float* gn0_w; const float* gn0_b; /* [cout][5][cin] */
  const float* conv1_w; const float* conv1_b; const float* gn1_w; const float* gn1_b; /* [cout][5][cout] */
  const float* film_w;  const float* film_b;  /* cond_encoder.1 [2*cout,32],[2*cout]              */
  const float* res_w;   const float* res_b;   /* residual_conv [cout][cin] or NULL (identity)     */
} dgdm_unet_resblock;

typedef struct dgdm_unet_weights {
  const float* se_w0; const float* se_b0;   /* diffusion_step_encoder.1 [128,32]                   */
  const float* se_w1; const float* se_b1;   /* diffusion_step_encoder.3 [32,128]                   */
  dgdm_unet_resblock blocks[8];             /* d00 d01 d10 d11 m0 m1 u00 u01                       */
  const float* down_w; const float* down_b; /* down_modules.0.2.conv [128][3][128]                 */
  const float* up_w;   const float* up_b;   /* up_modules.0.2.conv as [phase 2][cout 128][2][cin 128]:
                                             * even outputs use taps k=3,1, odd outputs k=2,0      */
  const float* fin_w;  const float* fin_b;  const float* fin_gn_w; const float* fin_gn_b; /* final_conv.0 */
  const float* out_w;  const float* out_b;  /* final_conv.1 [1][128],[1]                           */
  const void* tc_image;                     /* tensor-core weight image (dgdm_unet_pack_tc) or NULL */
} dgdm_unet_weights;

/* Tensor-core image of every conv whose contraction is tcgen05-eligible (Cin % 64 == 0): bf16 hi/lo split,
 * 64-wide k-blocks, 128B-swizzled K-major tiles. */
size_t dgdm_unet_tc_image_bytes(const dgdm_unet_weights* w);
int dgdm_unet_pack_tc(const dgdm_unet_weights* w, void* image, void* stream);

size_t dgdm_unet1d_workspace_bytes(int32_t n, int32_t P);
/* x [n,P] (the reference's (B,P,1)), t integer timestep, eps [n,P] (out).  precision: DGDM_PREC_* -- the
 * tensor-core modes run the implicit-GEMM convs on tcgen05 (3 bf16 MMAs per product in BF16X3). */
int dgdm_unet1d_forward(const dgdm_unet_weights* w, const float* x, int32_t n, int32_t P, int32_t t, float* eps,
                        void* workspace, size_t workspace_bytes, int32_t precision, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5  PointNet++ SSG object encoder, once per object.  Replaces PointNet2.forward
 * (dynamics/models/pointnet2.py:21-31) and pointnet2_utils.py:71-210.  BatchNorm2d folded.
 * Farthest-point-sampling starts are host-supplied (the reference draws them with torch.randint,
 * pointnet2_utils.py:83).
 */
typedef struct dgdm_pointnet2_weights {
  const float* w[5]; /* sa1: [64,3],[128,64]; sa2: [128,131],[256,128]; sa3: [256,259]  (BN folded) */
  const float* b[5];
} dgdm_pointnet2_weights;

size_t dgdm_pointnet2_workspace_bytes(int32_t n_clouds, int32_t n_points);
/* clouds [n,n_points,3] (n_points == 512), fps_start [n,2] int64, codes [n,256] (out). */
int dgdm_pointnet2_encode(const dgdm_pointnet2_weights* w, const float* clouds, int32_t n_clouds, int32_t n_points,
                          const int64_t* fps_start, float* codes, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6  best-of-N.  Replaces the np.argmax selection of get_best_ids* (generator/diffusion.py:346-428):
 * per object, the k best candidates by score, descending, ties to the lowest index.
 * scores [n_obj,n_cand]; idx [n_obj,k] int64 (out); best [n_obj,k] (out, may be NULL). 1 <= k <= 32.
 */
int dgdm_best_of_n(const float* scores, int32_t n_obj, int32_t n_cand, int32_t k, int64_t* idx, float* best,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Building block exported for unit tests: C = act(A . W^T + bias), fp32 CUDA-core GEMM.
 * A [M,K] (lda), W [N,K], C [M,N] (ldc).  relu != 0 applies max(.,0).
 */
int dgdm_linear_f32(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc,
                    int64_t M, int32_t N, int32_t K, int32_t relu, void* stream);
/* Same contraction on the tcgen05 path (K % 64 == 0, N in {128,256}); packs W into the workspace first. */
size_t dgdm_linear_tc_workspace_bytes(int32_t N, int32_t K);
int dgdm_linear_tc(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc, int64_t M,
                   int32_t N, int32_t K, int32_t relu, int32_t precision, void* workspace, size_t workspace_bytes,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGDM_B200_H_ */
