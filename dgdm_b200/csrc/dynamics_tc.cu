// K1 (tensor-core path) + K2: the dynamics-network trunk, forward + input-gradient backward, fused per
// 128-row tile on tcgen05 / TMEM, with the guidance reduction over pose rows fused on the tail.
//
// Work per guidance row r = (pair p, pose g)   [reference: dynamics/profile_forward_2d.py:109-135,154-155 forward,
// torch.autograd.grad at generator/diffusion.py:503-504 backward]:
//     a1 = relu(Cst[obj(p)] + U[design(p)] + V[g])                                  (hoisted layer 1)
//     a_l = relu(a_{l-1} W_l^T + b_l), l = 2..8 ;  logits = a_8 W_out^T + b_out      (7 + 1 GEMMs)
//     d8 = (dObj/dlogits . W_out) * 1[a_8>0] ;  d_{l-1} = (d_l W_l) * 1[a_{l-1}>0]    (7 GEMMs)
//     dUp[p] += d_1                                                                 (K2: sum over g)
// = 2*2*(7*256^2 + 256*3) = 1 838 080 FLOP per row (2D), none of it ever touching HBM: the row's activations
// live in TMEM, only ReLU sign bits (32 B per layer per row) are kept, in shared memory.
//
// Mapping to the SM (one persistent CTA per SM, 18 warps):
//   warp 16  producer: streams weight tiles global(L2) -> smem ring with cp.async.bulk (TMA unit, UBLKCP),
//            completion on mbarriers.  Tiles are pre-swizzled (128B swizzle, K-major) by dgdm_dyn_pack_tc so
//            the bytes land exactly in the canonical UMMA layout.
//   warp 17  MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16, M=128 (tile rows) x N=256 x K=16,
//            A operand from TMEM (the activations), B from smem (the weight tile), D (fp32) in TMEM;
//            tcgen05.commit releases smem stages and signals the epilogue.
//   warps 0-15 epilogue: tcgen05.ld the accumulator, bias + ReLU (or ReLU-mask in the backward), collect
//            sign bits, convert to bf16 (hi [+ lo]) and tcgen05.st it back as the next layer's A operand.
//            The last backward layer is reduced over the tile's rows per pair with warp shuffles and written
//            to a per-(pair,tile) slot; a small second kernel adds the slots in fixed order (deterministic).
// TMEM: two 256-column regions that swap roles every layer (A operand, per 16-feature group [hi 8 | lo 8] columns,
// and fp32 accumulator); the epilogue rewrites the accumulator in place into the next A operand.
//
// Why the accumulator is NOT split into N = 128 halves (round-2 experiment, scripts/dev/experiments/): issuing a
// layer as interleaved (half, k-block) units hides the accumulate -> epilogue -> next-layer chain (issuer waits drop
// from ~1000 to ~70 cycles per layer), but a TS-mode N = 128 MMA has to fetch its 4 KB A operand from TMEM every 64
// cycles, and that fetch loses against the epilogue warps' tcgen05.ld/st traffic: 64.0 cycles per MMA alone, 88-101
// with 16 epilogue warps active (scripts/microbench/mma_tmem_conflict.cu); an N = 256 MMA stays at 128.0.  Net: no
// gain (fp32-grade 0.917 vs 0.91, bf16 0.67 vs 0.70 of roofline), so a layer is issued as N = 256 MMAs.
//
// Precision modes: BF16X3 / FP16X3 split both operands into 16-bit hi + lo parts and issue 3 MMAs (hi*hi, lo*hi,
// hi*lo) into the same fp32 accumulator (bf16 parts: ~2^-16 relative product error; fp16 parts with the operand
// scaling below: ~2^-21), fp32-grade (<=1e-3 end to end).  BF16 and FP16 issue one MMA per product with bf16 / fp16
// operands (fp16: 3 more mantissa bits, ~8x fewer ReLU sign flips).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace dgdm {
namespace {

constexpr int TILE_M = 128;
constexpr int KBLK = 64;                       // bf16 elements per k-block row = 128 bytes = one swizzle span
constexpr int WTILE_BYTES = 256 * KBLK * 2;    // 32 KB ring slot: 256 rows x 128 B
// Operand scaling of the fp16 modes (all powers of two, so exact).  An fp16 residual part is ~2^-11 of its value and
// falls below the smallest normal fp16 (6.1e-5) for |value| < 0.125 -- every weight of this network, most activations
// and gradients -- where it keeps fewer and fewer bits (measured: unscaled fp16 hi+lo operands were 2x LESS accurate
// than bf16 hi+lo).  So the weight image holds W * SW, activations live in TMEM as a * SA and backward operands as
// d * GS; the epilogues undo SW with the multiply they already do (forward: fma(D, 1/SW, SA*b)) or one FMUL (backward).
constexpr float F16_SW = 256.f;                // weights of the fp16 hi+lo (x3) image; the single-pass fp16 image is unscaled:
                                               // without a lo part nothing goes subnormal, and the backward epilogue saves
                                               // its 16 FMULs per k-block (fp16 trunk 0.67 -> see DESIGN.md 4.1)
template <bool F16, bool X3> constexpr float sw_of() { return (F16 && X3) ? F16_SW : 1.f; }
constexpr float F16_SA = 64.f;                 // forward activations
constexpr float F16_GS = 64.f;                 // backward operands (seed); undone by reduce_slots together with SW
constexpr int NSTAGE = 5;
constexpr int NEPI = 16;                       // epilogue warps: 4 per TMEM lane quadrant, 16 features of every k-block each
constexpr int NTHREADS = (NEPI + 2) * 32;      // + producer warp + MMA warp
constexpr int MAX_SEG = 20;
// two-tile single-pass kernel (tc_trunk2_kernel)
constexpr int AKB_BYTES = 128 * 128;           // one A k-block: 128 rows x 128 B
// operand ring of the two-tile kernel: a stage = [weight part | A k-block 16 KB] of one k-block.  One CTA: the whole
// 32 KB weight tile, 4 stages (= one segment).  CTA pair (cta_group::2, M = 256 over two SMs): each CTA stages its
// 128 of the 256 weight rows (16 KB), so six 32 KB stages fit: one and a half segments of operands in flight.
template <int NCTA> struct Ring2 {
  static constexpr int NS = NCTA == 2 ? 6 : 4;
  static constexpr int WPART = 256 * KBLK * 2 / NCTA;
  static constexpr int STAGE = WPART + AKB_BYTES;          // a multiple of 1024: both parts stay swizzle-aligned
};
constexpr int NS2_MAX = 6;
constexpr int MAX_ITEMS = 44;                  // work items per tile pair and role: bit 7 = tile slot, bits 0..6 = segment
constexpr int IT_A1 = 0x7E;                    // item codes 0x7E / 0x7F: build the layer-1 operand, K-half 0 / 1
constexpr int MAX_CTAS2 = 160;                 // mask scratch is sized for this many CTAs
constexpr int MASK_WORDS = 16 + 7 * 8;         // layer 1 up to 512 wide + 7 layers of 256

// K_OUT: the 3-wide output layer as a 16-column MMA segment (bf16 mode).  K_FOUT (fp32-grade mode): the last hidden
// layer whose epilogue also evaluates the output layer on CUDA cores and seeds the backward pass -- with three MMAs per
// product the 16-column segment cost almost a full layer of tensor time plus two hand-off chains (+2.5 % without it);
// in bf16 mode the segment is cheap and the heavier fused epilogue measured 2 % slower, so both forms exist.
enum { K_FWD = 0, K_MID = 1, K_OUT = 2, K_BWD = 3, K_LAST = 4, K_FOUT = 5 };

struct Seg {
  uint32_t img_off;    // byte offset of the segment's tiles in the image
  uint16_t n_rows;     // B-operand rows = MMA N (256 or 16)
  uint8_t kind, layer, half, accum;
};

struct TcParams {
  const uint8_t* img;
  const float* U; const float* Cst; const float* Vt;   // Vt: pose table in quads, [H1/4][G][4]
  const float* bias[7]; const float* w_out; const float* b_out;
  const int32_t* pair_object;
  float* part;          // [n_slots][H1] per-(pair,tile) partial column sums  (backward)
  float* score_part;    // [n_slots] per-(pair,tile) partial objective sums   (forward only)
  float* logits;        // optional [n_rows,3]
  int* err;             // device error flag (mbarrier timeout)
  dgdm_objective obj;
  int64_t n_rows;       // n_pairs * G
  int n_tiles, G, H1, opd, n_designs, n_obj, n_seg, x3, backward;
  float gscale;         // scale of the backward seed (1, or F16_GS in the fp16 modes)
  int alt;              // 1: an odd number of region swaps per tile -> alternate the start region tile by tile
  Seg seg[MAX_SEG];
  // acc_mode (backward, G >= 128, one-tile kernel): a CTA owns a CONTIGUOUS range of tiles and adds a pair's per-tile
  // column sums up itself, in ascending tile order (the association of the old per-(pair, tile) slot buffer + reduce
  // kernel, bit for bit), writing dUp directly; only pairs cut by a CTA boundary go through `prefix` / `cont`.
  int acc_mode, max_cont;
  int merge_ab;         // single-pass modes: the A k-block's hand-over arrives on the weight tile's `full` barrier (one wait)
  float inv_gscale;
  float* dUp;           // [n_pairs][H1]
  float* prefix;        // [grid][H1]            sum over the tiles of the CTA's last pair when it continues in the next CTA
  float* cont;          // [grid][max_cont][H1]  per-tile sums of the CTA's leading tiles that continue an earlier CTA's pair
  // two-tile kernel (tc_trunk2_kernel): ReLU sign bits live in global scratch, work order lists for the two roles
  uint16_t* mask_scratch;
  int n_e, n_m;
  uint8_t e_items[MAX_ITEMS], m_items[MAX_ITEMS];
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps with an error code instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    if (it > (1u << 24)) {
      if (err) atomicExch(err, code);
      __trap();
    }
  }
}
// (A suspend-time hint on try_wait -- 2 us / 20 us, so that a warp waiting for a layer of MMAs sleeps in one try_wait instead of
// polling every ~200 cycles, 17 % of the bf16 trunk's executed instructions -- measured 1 % SLOWER on the same box: the wake-up is
// later, and the polling warps were not taking issue slots anybody needed.  The opposite -- a non-suspending test_wait poll for a
// faster wake -- is 1.5 % (bf16) / 2.3 % (fp32-grade) slower: sixteen spinning warps do take slots from the MMA issuer.)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tc_commit / tc_mma_ts are called by every lane of the issuing warp; lane 0 executes the instruction
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  if ((threadIdx.x & 31) == 0)
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T ; kind::f16 (bf16 in, fp32 accumulate)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  if ((threadIdx.x & 31) == 0)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  if ((threadIdx.x & 31) == 0)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
// CTA-pair forms (issued by the leader CTA only): M = 256 over the two SMs of the pair, each CTA supplies its 128 rows
// of A and its half of the B rows from the same shared-memory offsets; the commit arrives on the barrier at the same
// offset in BOTH CTAs
__device__ __forceinline__ void tc_mma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  if ((threadIdx.x & 31) == 0)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
  if ((threadIdx.x & 31) == 0)
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
// Issue only; the registers may be read after tmem_ld_wait16 (which names them so the compiler keeps the order).
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of one weight tile: K-major, 128-byte swizzle, rows of 128 B,
// 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: D fp32, A/B bf16 (format 1) or fp16 (format 0), both K-major, M = 128, N = n.
template <bool F16>
__device__ __forceinline__ constexpr uint32_t make_idesc(int n, int m = TILE_M) {
  return (1u << 4) | (F16 ? 0u : (1u << 7) | (1u << 10)) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ int pair_obj(int64_t p, int opd, int n_designs, int n_obj, const int32_t* pair_object) {
  if (pair_object) return pair_object[p];
  // 32-bit arithmetic whenever the pair index allows it: this runs on the tile-to-tile critical path and an emulated
  // 64-bit division costs ~150 cycles
  if (opd == 1) {
    const int dpo = n_designs / n_obj;
    return p <= 0x7fffffffll ? (int)((uint32_t)p / (uint32_t)dpo) : (int)(p / dpo);
  }
  return p <= 0x7fffffffll ? (int)((uint32_t)p % (uint32_t)opd) : (int)(p % opd);
}

// accumulator word -> value with the weight scale of the fp16 modes removed (see F16_SW)
template <bool F16, bool X3>
__device__ __forceinline__ float unscale(uint32_t acc) {
  return (F16 && X3) ? __uint_as_float(acc) * (1.f / F16_SW) : __uint_as_float(acc);
}
template <bool F16, bool X3>
__device__ __forceinline__ float unscale_add(uint32_t acc, float b) {
  return (F16 && X3) ? fmaf(__uint_as_float(acc), 1.f / F16_SW, b) : __uint_as_float(acc) + b;
}

__device__ __forceinline__ uint32_t cvt_rn_f16x2(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
__device__ __forceinline__ uint32_t cvt_rn_relu_f16x2(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
__device__ __forceinline__ uint32_t cvt_rz_relu_f16x2(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}

// Split 16 fp32 values into packed 16-bit hi (and, X3, bf16 lo = rn(v - hi)) pairs; element 2i in the low half.
template <bool X3, bool F16>
__device__ __forceinline__ void split_pack(const float (&v)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (F16) {
      hi[i] = cvt_rn_f16x2(v[2 * i], v[2 * i + 1]);
      if (X3) {
        const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi[i]));
        lo[i] = cvt_rn_f16x2(v[2 * i] - h.x, v[2 * i + 1] - h.y);
      }
      continue;
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    hi[i] = *reinterpret_cast<uint32_t*>(&h);
    if (X3) {
      // unpack with a shift and a mask (no PRMT pair): fewer ops on the half-rate ALU pipe
      const float h0 = __uint_as_float(hi[i] << 16), h1 = __uint_as_float(hi[i] & 0xFFFF0000u);
      __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * i] - h0, v[2 * i + 1] - h1);
      lo[i] = *reinterpret_cast<uint32_t*>(&l);
    }
  }
}

// ---- ReLU + split + sign bits with the conversion unit doing the ReLU ------------------------------------------
// The epilogue is bound by the half-rate ALU pipe (FSETP/SEL/FMNMX/PRMT...), not by TMEM or latency, so the
// forward path avoids it: cvt.{rz|rn}.relu.bf16x2 clamps negatives while converting, the lo residual
// (z - float(hi), >= 0 for z >= 0 because hi is truncated, negative for z < 0) goes through the same .relu
// conversion, and the 16 sign bits are read off the exponent bytes of the packed hi words, four elements at a time.
__device__ __forceinline__ uint32_t cvt_rz_relu_bf16x2(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
__device__ __forceinline__ uint32_t cvt_rn_relu_bf16x2(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
// bit position of element e (0..15) inside the 16-bit sign word produced by relu_split16
__host__ __device__ constexpr int mask_pos(int e) { return ((e & 1) * 8) + (((e >> 1) & 1) * 4) + 3 - (e >> 2); }

// z[16] (pre-activation) -> packed 16-bit hi[8] (+ bf16 lo[8]) of relu(z).
template <bool X3, bool F16>
__device__ __forceinline__ void relu_split16(const float (&z)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (F16 && X3) {
      // hi truncated (z - hi >= 0 for z >= 0), so the residual goes through the same .relu conversion
      hi[i] = cvt_rz_relu_f16x2(z[2 * i], z[2 * i + 1]);
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi[i]));
      lo[i] = cvt_rn_relu_f16x2(z[2 * i] - h.x, z[2 * i + 1] - h.y);
    } else if (F16) {
      hi[i] = cvt_rn_relu_f16x2(z[2 * i], z[2 * i + 1]);
    } else if (X3) {
      hi[i] = cvt_rz_relu_bf16x2(z[2 * i], z[2 * i + 1]);
      const float h0 = __uint_as_float(hi[i] << 16), h1 = __uint_as_float(hi[i] & 0xFFFF0000u);
      lo[i] = cvt_rn_relu_bf16x2(z[2 * i] - h0, z[2 * i + 1] - h1);
    } else {
      hi[i] = cvt_rn_relu_bf16x2(z[2 * i], z[2 * i + 1]);
    }
  }
}
// The 16 sign bits (mask_pos layout) of the packed hi words of relu(z).  Computed AFTER the k-block has been handed to
// the MMA issuer: the ~20 ALU instructions are then off the accumulate -> epilogue -> next-layer critical path.
// A value whose top byte is zero counts as zero: bf16 exponent field < 2 (|z| < 2^-125); fp16 z < 2^-16 (a
// subnormal below a quarter of the smallest normal -- 30x below the fp16 rounding noise of an O(1) pre-activation).
__device__ __forceinline__ uint32_t sign_bits16(const uint32_t (&hi)[8]) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // top bytes of elements 4k..4k+3 -> one word; byte != 0  <=>  (byte + 0x7F) has its MSB set
    const uint32_t t = __byte_perm(hi[2 * k], hi[2 * k + 1], 0x7531) + 0x7F7F7F7Fu;
    m |= (t & 0x80808080u) >> k;
  }
  m >>= 4;                                         // flags now at bits {0..3, 8..11, 16..19, 24..27}
  return (m | (m >> 12)) & 0xFFFFu;
}

// Column sums over the 32 lanes of a warp for 32 per-lane values: after the butterfly lane i holds
// sum over lanes of v[i].  31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      float keep = upper ? v[i + w] : v[i];
      float send = upper ? v[i] : v[i + w];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

// Optional in-kernel timeline (build with DGDM_NVCC_EXTRA=-DDGDM_TRUNK_TRACE python -m dgdm_b200.build -f, read with
// dgdm_trunk_trace_read; scratch/trace2.py prints it): clock64 stamps of CTA 0, third tile, per role.
// Development tool; compiled out by default.
#ifdef DGDM_TRUNK_TRACE
__device__ long long g_trace[8192];
#define TR(slot) do { if (blockIdx.x == 0 && t == 2) g_trace[slot] = clock64(); } while (0)
// epilogue warp 0, k-block 0 of segment tr_sg: fine-grained stamps of the hand-off chain (slots 2048 + sg*16 + 6..10)
#define TRE(kb, k) do { if ((kb) == 0 && warp == 0 && lane == 0) TR(2048 + tr_sg * 16 + (k)); } while (0)
#else
#define TR(slot) do { } while (0)
#define TRE(kb, k) do { } while (0)
#endif

struct Smem {
  uint64_t full[NSTAGE], empty[NSTAGE], a_ready[4], d_ready;
  uint32_t tmem_base, pad_;
  alignas(16) float bias[7][256];
  alignas(16) float w_out[3][256];
  float b_out[4];
  float red[4][256];            // per lane-quadrant partial column sums
  float red_s[4];
  float lg[4][TILE_M][3];      // partial logits of the four 16-feature-group warps of a row
  uint16_t mask[2 * MASK_WORDS][TILE_M];      // ReLU sign bits, one halfword per (16-feature group, row)
  float acc[512];               // acc_mode: running column sums of the pair that spans the current tiles
};

// acc_mode tile ownership: CTA b owns tiles [tile0, tile0 + count): the first n_tiles % grid CTAs one tile more
__host__ __device__ inline void cta_tiles(int n_tiles, int grid, int b, int& tile0, int& count) {
  const int q = n_tiles / grid, r = n_tiles % grid;
  tile0 = b * q + (b < r ? b : r);
  count = q + (b < r ? 1 : 0);
}
__host__ __device__ inline int tile_owner(int n_tiles, int grid, int tile) {
  const int q = n_tiles / grid, r = n_tiles % grid;
  return tile < r * (q + 1) ? tile / (q + 1) : r + (tile - r * (q + 1)) / q;
}

template <bool X3, bool F16>
__global__ void __launch_bounds__(NTHREADS, 1) tc_trunk_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  // weight ring first (1024-byte aligned for the 128B swizzle), bookkeeping after
  // (pointer + offset, not an integer round trip: the compiler then keeps the shared address space and emits LDS/STS
  // for the bias / sign-bit accesses instead of generic loads and stores)
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Smem& S = *reinterpret_cast<Smem*>(ring + NSTAGE * WTILE_BYTES);

  // the warp index through a shuffle: the compiler then knows it is warp-uniform, and the MMA issuer below -- executed by
  // all 32 lanes of its warp, only the instructions themselves predicated on lane 0 -- keeps every tcgen05.mma operand
  // in uniform registers (inside an `if (lane == 0)` region each MMA paid an R2UR of its operand address)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

  if (tid == 0) {
    // merge_ab (single-pass modes): ONE barrier per weight tile covers both operands -- the weight bytes (producer's
    // arrive.expect_tx) and the A k-block (one arrive per epilogue warp) -- so the issuing thread, which sets the pace
    // between weight tiles (scripts/microbench/mma_2cta.cu), does one wait per four MMAs instead of two.  The fp32-grade
    // modes keep the a_ready barriers (a k-block there is two weight tiles, and they run at the power cap anyway).
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&S.full[s], P.merge_ab ? 1 + NEPI : 1); mbar_init(&S.empty[s], 1); }
    for (int k = 0; k < 4; ++k) mbar_init(&S.a_ready[k], NEPI);
    mbar_init(&S.d_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 7 * 256; i += NTHREADS) S.bias[i / 256][i % 256] = P.bias[i / 256][i % 256] * (F16 ? F16_SA : 1.f);
  for (int i = tid; i < 3 * 256; i += NTHREADS) S.w_out[i / 256][i % 256] = P.w_out[i];
  if (tid < 3) S.b_out[tid] = P.b_out[tid];
  if (warp == NEPI + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  // All 512 columns are allocated, so the base is column 0 of lane 0.  The MMA issuer relies on that: with literal
  // TMEM addresses every tcgen05.mma operand is warp-uniform and lives in uniform registers; taking the base from
  // shared memory made nvcc wrap EVERY mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~100 cycles of issue each,
  // which starved the tensor pipe between weight tiles).
  if (tmem != 0u) { if (tid == 0 && P.err) atomicExch(P.err, 9); __trap(); }

  int tile0 = (int)blockIdx.x, tiles_mine = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  int tile_step = (int)gridDim.x;
  if (P.acc_mode) { cta_tiles(P.n_tiles, (int)gridDim.x, (int)blockIdx.x, tile0, tiles_mine); tile_step = 1; }
  const int tiles_per_seg = X3 ? 8 : 4;       // weight tiles per segment: 4 k-blocks x (hi [, lo])

  if (warp == NEPI) {
    // =============================== producer ===============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < tiles_mine; ++t) {
        for (int sg = 0; sg < P.n_seg; ++sg) {
          const Seg sgm = P.seg[sg];
          const uint32_t tile_bytes = (uint32_t)sgm.n_rows * 128u;
          for (int j = 0; j < tiles_per_seg; ++j) {
            // image order inside a segment: kb0.hi, kb0.lo, kb1.hi, kb1.lo, ...
            const int kb = X3 ? (j >> 1) : j, part = X3 ? (j & 1) : 0;
            mbar_wait(&S.empty[stage], phase ^ 1, P.err, 1);
            TR(6144 + sg * 8 + j);
            mbar_arrive_expect_tx(&S.full[stage], tile_bytes);
            bulk_g2s(ring + stage * WTILE_BYTES, P.img + sgm.img_off + (size_t)(kb * 2 + part) * tile_bytes, tile_bytes,
                     &S.full[stage]);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == NEPI + 1) {
    // =============================== MMA issuer ===============================
    // TMEM is two 256-column regions.  During a segment the A operand lives in region `cur` (per 64-wide
    // K=16 step ks of k-block kb: hi at [64kb+16ks, +8), lo at [64kb+16ks+8, +8)) and the accumulator in the other one.
    // The epilogue converts the accumulator IN PLACE into the next layer's A operand, k-block by k-block, and hands
    // each k-block over on its own mbarrier, so the next layer's MMAs start after a quarter of the epilogue instead
    // of all of it; the roles of the two regions then swap.
    {   // whole warp, uniform control flow; tc_mma_ts / tc_commit issue from lane 0
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int t = 0; t < tiles_mine; ++t) {
        // fused-output plan: 13 region swaps per backward tile: alternate the start region so that the next tile's layer-1 operand lands in
        // the region K_LAST's MMAs have finished with, never in the accumulator its epilogue is still reading
        uint32_t cur = P.alt ? (uint32_t)(t & 1) : 0u;
        for (int sg = 0; sg < P.n_seg; ++sg) {
          const Seg sgm = P.seg[sg];
          const uint32_t a_base = cur * 256u, d_base = (cur ^ 1u) * 256u;      // TMEM base is 0 (checked above)
          const uint32_t idesc = make_idesc<F16>(sgm.n_rows);
          uint32_t accum = sgm.accum;
          for (int j = 0; j < tiles_per_seg; ++j) {
            const int kb = X3 ? (j >> 1) : j, part = X3 ? (j & 1) : 0;
            TR(sg * 64 + j * 4 + 0);
            if (part == 0 && !(!X3 && P.merge_ab)) mbar_wait(&S.a_ready[kb], a_phase, P.err, 2);
            TR(sg * 64 + j * 4 + 1);
            mbar_wait(&S.full[stage], phase, P.err, 3);
            tc_fence_after();
            TR(sg * 64 + j * 4 + 2);
            // descriptors of the four K = 16 steps: + 32 bytes = + 2 in the 14-bit address field (no carry below 256 KB):
            // one add per MMA instead of a shift / mask / or chain on the uniform datapath (the issuing thread, not the
            // tensor pipe, sets the pace between weight tiles: scripts/microbench/mma_2cta.cu)
            const uint64_t bdesc0 = make_b_desc(smem_u32(ring + stage * WTILE_BYTES));
#pragma unroll
            for (int ks = 0; ks < KBLK / 16; ++ks) {
              const uint64_t bdesc = bdesc0 + (uint64_t)(2 * ks);
              const uint32_t a_col = (uint32_t)(kb * 64 + ks * 16);            // 16 operand elements = 8 TMEM columns: [hi 8 | lo 8]
              tc_mma_ts(d_base, a_base + a_col, bdesc, idesc, accum);
              accum = 1;
              if (X3 && part == 0) tc_mma_ts(d_base, a_base + a_col + 8, bdesc, idesc, 1);
            }
            tc_commit(&S.empty[stage]);           // stage reusable once these MMAs have read it
            TR(sg * 64 + j * 4 + 3);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          tc_commit(&S.d_ready);                  // accumulator complete
          a_phase ^= 1;
          if (sgm.kind == K_FWD || sgm.kind == K_BWD || sgm.kind == K_OUT || sgm.kind == K_FOUT) cur ^= 1;
        }
      }
    }
  } else {
    // =============================== epilogue warps 0..15 ===============================
    // warp (q, hq): TMEM lane quadrant q (rows 32q..32q+31); of every 64-feature k-block it owns the 16 features
    // [16hq, 16hq+16).  Four warps per scheduler keep the issue slots busy while a sibling waits on a TMEM load, a
    // tcgen05.wait::st (with two lock-stepped warps per scheduler the epilogue was latency-bound at ~1100 cycles per
    // k-block).  A warp rewrites exactly the 16 accumulator columns it read ([hi 8 | lo 8] of the same features), so
    // the four warps of a quadrant never wait for each other in the layer loop.
    const int q = warp & 3, hq = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const int l1_hw = P.H1 / 16;                  // halfwords of layer-1 sign bits per row
    uint32_t d_phase = 0;
    [[maybe_unused]] uint32_t pos_stage = 0;     // merge_ab: ring stage of weight tile 0 of the segment whose operand comes next
    auto quad_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory"); };

    for (int t = 0; t < tiles_mine; ++t) {
      const int tile = tile0 + t * tile_step;
      // Row -> (pair, pose row) with ONE 64-bit division per tile; everything per row is 32-bit (this code sits on the
      // tile-to-tile critical path: the timeline showed ~1900 cycles of index arithmetic here with 64-bit divisions).
      const int64_t base = (int64_t)tile * TILE_M;
      const int64_t p_first = base / P.G;
      const uint32_t rem0 = (uint32_t)(base - p_first * P.G), Gu = (uint32_t)P.G;
      const uint32_t rofs = rem0 + (uint32_t)row, dq = rofs / Gu;          // rofs < G + 128
      const int64_t r_glob = base + row;
      const bool live = r_glob < P.n_rows;
      const int64_t pr = live ? p_first + dq : -1;
      const int g = live ? (int)(rofs - dq * Gu) : 0;
      const float coef = (live && P.obj.row_coef) ? P.obj.row_coef[r_glob] : 1.f;
      const float* u_row = nullptr; const float* c_row = nullptr;
      if (live) {
        const int64_t design = P.opd == 1 ? pr : (pr <= 0x7fffffffll ? (int64_t)((uint32_t)pr / (uint32_t)P.opd) : pr / P.opd);
        u_row = P.U + design * P.H1;
        c_row = P.Cst + (int64_t)pair_obj(pr, P.opd, P.n_designs, P.n_obj, P.pair_object) * P.H1;
      }
      int64_t last_row = base + TILE_M - 1;
      if (last_row >= P.n_rows) last_row = P.n_rows - 1;
      const int64_t p_last = p_first + (rem0 + (uint32_t)(last_row - base)) / Gu;
      uint32_t cur = P.alt ? (uint32_t)(t & 1) : 0u;   // region holding the A operand of the current segment (see the issuer)

      // hand k-block kb of the A operand over to the MMA issuer
      [[maybe_unused]] int tr_sg = 0;
      // (merge_ab: the k-block goes with weight tile kb of the segment it feeds, i.e. ring stage pos_stage + kb; every
      // segment's operand is produced exactly once and takes four weight tiles, so pos_stage tracks the issuer's counter.
      // Phase safety: position p - NSTAGE of that stage lies in the layer whose accumulator this epilogue is reading, so
      // its phase has completed.)
      auto signal_kb = [&](int kb) {
        tmem_wait_st();
        TRE(kb, 9);
        tc_fence_before();
        __syncwarp();
        if (!X3 && P.merge_ab) {
          if (lane == 0) { uint32_t st = pos_stage + (uint32_t)kb; if (st >= NSTAGE) st -= NSTAGE; mbar_arrive(&S.full[st]); }
          if (kb == 3) { pos_stage += 4; if (pos_stage >= NSTAGE) pos_stage -= NSTAGE; }
        } else if (lane == 0) {
          mbar_arrive(&S.a_ready[kb]);
        }
        TRE(kb, 10);
      };
      // store 16 features (k-block kb, group hq) of this row as bf16 hi [+ lo] into region `reg`
      auto store_packed = [&](uint32_t reg, int kb, const uint32_t (&hi)[8], const uint32_t (&lo)[8]) {
        // the 16 accumulator columns this warp read become [hi 8 | lo 8] of the same 16 features: no other warp's
        // columns are touched, so the in-place rewrite needs no barrier between the warps of a lane quadrant
        const uint32_t col = reg * 256u + (uint32_t)(kb * 64 + hq * 16);
        tmem_st8(lane_addr + col, hi);
        if (X3) tmem_st8(lane_addr + col + 8, lo);
      };
      auto store_a = [&](uint32_t reg, int kb, const float (&v)[16]) {
        uint32_t hi[8], lo[8];
        split_pack<X3, F16>(v, hi, lo);
        store_packed(reg, kb, hi, lo);
      };
      // layer-1 activations for K-half `kh` -> A operand in region `reg` + layer-1 sign bits
      // z = Cst[obj] + U[design] + V[g] for the 16 features [col0, col0 + 16) of this row.
      // The pose table is read TRANSPOSED ([H1][G]): lanes are consecutive pose rows g, so each feature is one
      // coalesced 128-byte request; all loads are issued before use.  U/Cst rows are warp-broadcast loads.
      auto load_z = [&](int col0, float (&z)[16]) {
        float4 pv[4];
        {
          // 32-bit byte offsets from the (uniform) table base: one add per load instead of a 64-bit multiply-add chain
          // (the address arithmetic was half of this function's instructions, and the epilogue is issue-bound)
          const char* vb = reinterpret_cast<const char*>(P.Vt);
          const uint32_t step = (uint32_t)P.G * 16u;
          uint32_t off = (uint32_t)(col0 >> 2) * step + (uint32_t)g * 16u;
#pragma unroll
          for (int i = 0; i < 4; ++i, off += step) pv[i] = live ? __ldg(reinterpret_cast<const float4*>(vb + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) {
            float4 k4 = *reinterpret_cast<const float4*>(c_row + col0 + i);
            float4 u4 = *reinterpret_cast<const float4*>(u_row + col0 + i);
            const float4 p4 = pv[i >> 2];
            a.x = (k4.x + u4.x) + p4.x; a.y = (k4.y + u4.y) + p4.y; a.z = (k4.z + u4.z) + p4.z; a.w = (k4.w + u4.w) + p4.w;
          }
          if (F16) { a.x *= F16_SA; a.y *= F16_SA; a.z *= F16_SA; a.w *= F16_SA; }     // fp16 modes: operands are a * SA
          z[i] = a.x; z[i + 1] = a.y; z[i + 2] = a.z; z[i + 3] = a.w;
        }
      };
      // The loads of k-block kb+1 are in flight while k-block kb is converted and handed over: an L2 round trip is
      // ~1300 cycles here, far more than a bf16 k-block of MMAs (512), so without the overlap layer 1 is load-bound.
      // (Measured: +1.6 % in bf16.  In fp32-grade mode a k-block of MMAs (1536 cycles) already covers the round trip and
      // the second buffer only costs registers at the 96-register cap, so that mode keeps the plain loop.)
      auto build_a1 = [&](int kh, uint32_t reg) {
        if constexpr (X3) {
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            float z[16];
            load_z(kh * 256 + kb * 64 + hq * 16, z);
            uint32_t hi[8], lo[8];
            relu_split16<X3, F16>(z, hi, lo);
            store_packed(reg, kb, hi, lo);
            signal_kb(kb);
            S.mask[kh * 16 + kb * 4 + hq][row] = (uint16_t)sign_bits16(hi);
          }
        } else {
        float zb[2][16];
        load_z(kh * 256 + hq * 16, zb[0]);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (kb + 1 < 4) load_z(kh * 256 + (kb + 1) * 64 + hq * 16, zb[(kb + 1) & 1]);
          uint32_t hi[8], lo[8];
          relu_split16<X3, F16>(zb[kb & 1], hi, lo);
          store_packed(reg, kb, hi, lo);
          signal_kb(kb);
          S.mask[kh * 16 + kb * 4 + hq][row] = (uint16_t)sign_bits16(hi);
        }
        }
      };

      build_a1(0, cur);

      for (int sg = 0; sg < P.n_seg; ++sg) {
        const Seg sgm = P.seg[sg];
        const bool last_seg = sg == P.n_seg - 1;
        const uint32_t dreg = cur ^ 1u;             // accumulator region of this segment
        const uint32_t d_addr = lane_addr + dreg * 256u;
        tr_sg = sg;
        mbar_wait(&S.d_ready, d_phase, P.err, 4);
        d_phase ^= 1;
        tc_fence_after();
        if (lane == 0 && (warp == 0 || warp == 15)) TR(2048 + (warp == 15 ? 2048 : 0) + sg * 16);

        if (sgm.kind == K_MID) {
          build_a1(1, cur);                         // second K-half of a1 replaces the first; accumulator stays
        } else if (sgm.kind == K_FWD || sgm.kind == K_BWD) {
          // FWD: a = relu(D + b), record sign bits of this layer.   BWD: d = D * 1[a_{layer} > 0].
          const int mbase = sgm.layer == 0 ? 0 : l1_hw + (sgm.layer - 1) * 16;
          uint32_t rr[2][16];
          tmem_ld16_async(d_addr + (uint32_t)(hq * 16), rr[0]);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            tmem_ld_wait16(rr[kb & 1]);
            TRE(kb, 6);
            if (kb + 1 < 4) tmem_ld16_async(d_addr + (uint32_t)((kb + 1) * 64 + hq * 16), rr[(kb + 1) & 1]);
            if (sgm.kind == K_FWD) {
              const float4* b4 = reinterpret_cast<const float4*>(&S.bias[sgm.layer - 1][kb * 64 + hq * 16]);
              float z[16];
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 bb = b4[i4];                // (fp16 modes: SA * b; D = SA * SW * (a . w))
                z[i4 * 4 + 0] = unscale_add<F16, X3>(rr[kb & 1][i4 * 4 + 0], bb.x);
                z[i4 * 4 + 1] = unscale_add<F16, X3>(rr[kb & 1][i4 * 4 + 1], bb.y);
                z[i4 * 4 + 2] = unscale_add<F16, X3>(rr[kb & 1][i4 * 4 + 2], bb.z);
                z[i4 * 4 + 3] = unscale_add<F16, X3>(rr[kb & 1][i4 * 4 + 3], bb.w);
              }
              uint32_t hi[8], lo[8];
              relu_split16<X3, F16>(z, hi, lo);
              TRE(kb, 7);
              store_packed(dreg, kb, hi, lo);       // in place: the accumulator region becomes the next A operand
              TRE(kb, 8);
              signal_kb(kb);
              S.mask[mbase + kb * 4 + hq][row] = (uint16_t)sign_bits16(hi);   // after the hand-off: off the critical path
            } else {
              const uint32_t bits = S.mask[mbase + kb * 4 + hq][row];
              if constexpr (!X3) {
                // single-pass modes: convert first, then clear the masked halfwords of each packed word -- shift the
                // pair's two sign bits to the top of bytes 0 / 1, PRMT replicates them into a halfword mask, one AND:
                // 3 instructions per two elements instead of 4 (test + select each); same bits as select-then-convert
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float v0 = unscale<F16, X3>(rr[kb & 1][2 * i]), v1 = unscale<F16, X3>(rr[kb & 1][2 * i + 1]);
                  uint32_t w;
                  if (F16) w = cvt_rn_f16x2(v0, v1);
                  else { __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1); w = *reinterpret_cast<uint32_t*>(&h); }
                  uint32_t m;
                  asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(m) : "r"(bits << (7 - mask_pos(2 * i))));
                  hi[i] = w & m;
                }
                TRE(kb, 7);
                store_packed(dreg, kb, hi, lo);
              } else {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = (bits >> mask_pos(i)) & 1u ? unscale<F16, X3>(rr[kb & 1][i]) : 0.f;
                TRE(kb, 7);
                store_a(dreg, kb, v);
              }
              TRE(kb, 8);
              signal_kb(kb);
            }
            if (lane == 0 && (warp == 0 || warp == 15)) TR(2048 + (warp == 15 ? 2048 : 0) + sg * 16 + 1 + kb);
          }
          cur ^= 1;
        } else if (!X3 && sgm.kind == K_OUT) {          // (compile-time: only the bf16 instantiation has this form)
          uint32_t rr[8];
          tmem_ld8(d_addr, rr);
          constexpr float inv_out = F16 ? 1.f / (F16_SA * sw_of<F16, X3>()) : 1.f;       // D = SA * SW * (a8 . w_out)
          const float l0 = __uint_as_float(rr[0]) * inv_out + S.b_out[0], l1 = __uint_as_float(rr[1]) * inv_out + S.b_out[1],
                      l2 = __uint_as_float(rr[2]) * inv_out + S.b_out[2];
          if (hq == 0 && live && P.logits) {
            float* o = P.logits + r_glob * 3;
            o[0] = l0; o[1] = l1; o[2] = l2;
          }
          if (!P.backward) {
            // forward only: per-pair sums of the objective over this tile's rows -> score_part[p + tile]
            const float val = live ? coef * (P.obj.c[0] * l0 + P.obj.c[1] * l1 + P.obj.c[2] * l2 + P.obj.sq0 * l0 * l0) : 0.f;
            if (P.G == 1) {                         // explicit-row mode: one row per pair, nothing to reduce
              if (hq == 0 && live) P.score_part[pr + tile] = val;
            } else
            for (int64_t ps = p_first; ps <= p_last; ++ps) {
              float s = (pr == ps) ? val : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
              if (hq == 0 && lane == 0) S.red_s[q] = s;
              asm volatile("bar.sync 1, 512;" ::: "memory");
              if (tid == 0) P.score_part[ps + tile] = (S.red_s[0] + S.red_s[1]) + (S.red_s[2] + S.red_s[3]);
              asm volatile("bar.sync 1, 512;" ::: "memory");
            }
          } else {
            // seed of the backward pass: d8 = (dObj/dlogits . W_out) * 1[a_8 > 0], written over the logits' region
            const float cg = coef * P.gscale;
            const float dl0 = cg * (P.obj.c[0] + 2.f * P.obj.sq0 * l0), dl1 = cg * P.obj.c[1], dl2 = cg * P.obj.c[2];
            const int mbase = l1_hw + 6 * 16;
            quad_sync();                            // the sibling warps have read the logits too
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {
              const int col0 = kb * 64 + hq * 16;
              const uint32_t bits = S.mask[mbase + kb * 4 + hq][row];
              float v[16];
              const float4* w0 = reinterpret_cast<const float4*>(&S.w_out[0][col0]);
              const float4* w1 = reinterpret_cast<const float4*>(&S.w_out[1][col0]);
              const float4* w2 = reinterpret_cast<const float4*>(&S.w_out[2][col0]);
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 a0 = w0[i4], a1 = w1[i4], a2 = w2[i4];
                const float d[4] = {dl0 * a0.x + dl1 * a1.x + dl2 * a2.x, dl0 * a0.y + dl1 * a1.y + dl2 * a2.y,
                                    dl0 * a0.z + dl1 * a1.z + dl2 * a2.z, dl0 * a0.w + dl1 * a1.w + dl2 * a2.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) v[i4 * 4 + e] = (bits >> mask_pos(i4 * 4 + e)) & 1u ? d[e] : 0.f;
              }
              store_a(dreg, kb, v);
              signal_kb(kb);
            }
            cur ^= 1;
          }
        } else if (X3 && sgm.kind == K_FOUT) {          // (compile-time: only the fp32-grade instantiation)
          // Last hidden layer: a8 = relu(D + b) is consumed right here by the 3-wide output layer on CUDA cores
          // (fp32 FMAs against W_out in shared memory; the four 16-feature-group warps of a row add their partial
          // logits through shared memory in a fixed order), so a8 never becomes an MMA operand.
          const int mbase = l1_hw + (sgm.layer - 1) * 16;
          float pl0 = 0.f, pl1 = 0.f, pl2 = 0.f;
          {
            uint32_t rr[2][16];
            tmem_ld16_async(d_addr + (uint32_t)(hq * 16), rr[0]);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              tmem_ld_wait16(rr[kb & 1]);
              if (kb + 1 < 4) tmem_ld16_async(d_addr + (uint32_t)((kb + 1) * 64 + hq * 16), rr[(kb + 1) & 1]);
              const int col0 = kb * 64 + hq * 16;
              const float4* b4 = reinterpret_cast<const float4*>(&S.bias[sgm.layer - 1][col0]);
              const float4* w0 = reinterpret_cast<const float4*>(&S.w_out[0][col0]);
              const float4* w1 = reinterpret_cast<const float4*>(&S.w_out[1][col0]);
              const float4* w2 = reinterpret_cast<const float4*>(&S.w_out[2][col0]);
              uint32_t m = 0;                          // sign bits of a8 in the mask_pos layout
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 bb = b4[i4], a0 = w0[i4], a1 = w1[i4], a2 = w2[i4];
                const float zz[4] = {unscale_add<F16, X3>(rr[kb & 1][i4 * 4 + 0], bb.x), unscale_add<F16, X3>(rr[kb & 1][i4 * 4 + 1], bb.y),
                                     unscale_add<F16, X3>(rr[kb & 1][i4 * 4 + 2], bb.z), unscale_add<F16, X3>(rr[kb & 1][i4 * 4 + 3], bb.w)};
                const float wa[4] = {a0.x, a0.y, a0.z, a0.w}, wb[4] = {a1.x, a1.y, a1.z, a1.w}, wc[4] = {a2.x, a2.y, a2.z, a2.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a = fmaxf(zz[e], 0.f);
                  m |= (a > 0.f ? 1u : 0u) << mask_pos(i4 * 4 + e);
                  pl0 = fmaf(a, wa[e], pl0); pl1 = fmaf(a, wb[e], pl1); pl2 = fmaf(a, wc[e], pl2);
                }
              }
              S.mask[mbase + kb * 4 + hq][row] = (uint16_t)m;
            }
          }
          if (F16) { pl0 *= 1.f / F16_SA; pl1 *= 1.f / F16_SA; pl2 *= 1.f / F16_SA; }     // a8 above is SA * relu(z8)
          S.lg[hq][row][0] = pl0; S.lg[hq][row][1] = pl1; S.lg[hq][row][2] = pl2;
          quad_sync();
          const float l0 = S.b_out[0] + ((S.lg[0][row][0] + S.lg[1][row][0]) + (S.lg[2][row][0] + S.lg[3][row][0]));
          const float l1 = S.b_out[1] + ((S.lg[0][row][1] + S.lg[1][row][1]) + (S.lg[2][row][1] + S.lg[3][row][1]));
          const float l2 = S.b_out[2] + ((S.lg[0][row][2] + S.lg[1][row][2]) + (S.lg[2][row][2] + S.lg[3][row][2]));
          quad_sync();                              // lg may be rewritten (next tile) only after all four warps read it
          if (hq == 0 && live && P.logits) {
            float* o = P.logits + r_glob * 3;
            o[0] = l0; o[1] = l1; o[2] = l2;
          }
          if (!P.backward) {
            // forward only: per-pair sums of the objective over this tile's rows -> score_part[p + tile]
            const float val = live ? coef * (P.obj.c[0] * l0 + P.obj.c[1] * l1 + P.obj.c[2] * l2 + P.obj.sq0 * l0 * l0) : 0.f;
            if (P.G == 1) {                         // explicit-row mode: one row per pair, nothing to reduce
              if (hq == 0 && live) P.score_part[pr + tile] = val;
            } else
            for (int64_t ps = p_first; ps <= p_last; ++ps) {
              float s = (pr == ps) ? val : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
              if (hq == 0 && lane == 0) S.red_s[q] = s;
              asm volatile("bar.sync 1, 512;" ::: "memory");
              if (tid == 0) P.score_part[ps + tile] = (S.red_s[0] + S.red_s[1]) + (S.red_s[2] + S.red_s[3]);
              asm volatile("bar.sync 1, 512;" ::: "memory");
            }
          } else {
            // seed of the backward pass: d8 = (dObj/dlogits . W_out) * 1[a_8 > 0], written in place over the
            // accumulator columns this warp read (the region becomes the first backward layer's A operand)
            const float cg = coef * P.gscale;
            const float dl0 = cg * (P.obj.c[0] + 2.f * P.obj.sq0 * l0), dl1 = cg * P.obj.c[1], dl2 = cg * P.obj.c[2];
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {
              const int col0 = kb * 64 + hq * 16;
              const uint32_t bits = S.mask[mbase + kb * 4 + hq][row];
              float v[16];
              const float4* w0 = reinterpret_cast<const float4*>(&S.w_out[0][col0]);
              const float4* w1 = reinterpret_cast<const float4*>(&S.w_out[1][col0]);
              const float4* w2 = reinterpret_cast<const float4*>(&S.w_out[2][col0]);
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 a0 = w0[i4], a1 = w1[i4], a2 = w2[i4];
                const float d[4] = {dl0 * a0.x + dl1 * a1.x + dl2 * a2.x, dl0 * a0.y + dl1 * a1.y + dl2 * a2.y,
                                    dl0 * a0.z + dl1 * a1.z + dl2 * a2.z, dl0 * a0.w + dl1 * a1.w + dl2 * a2.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) v[i4 * 4 + e] = (bits >> mask_pos(i4 * 4 + e)) & 1u ? d[e] : 0.f;
              }
              store_a(dreg, kb, v);
              signal_kb(kb);
            }
            cur ^= 1;
          }
        } else {   // K_LAST: d1 = D * 1[a_1 > 0], summed over the rows of each pair present in the tile (K2)
          // warp (q, hq) reduces accumulator columns [64hq, 64hq+64) as two 32-column chunks
          if (P.G == 1) {                           // explicit-row mode: the row IS the pair; store it directly
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              const int cc = hq * 2 + c;              // 32-feature chunk index 0..7
              uint32_t rr[32];
              tmem_ld32(d_addr + (uint32_t)(cc * 32), rr);
              if (live) {
                const uint32_t b0 = S.mask[(sgm.half * 8 + cc) * 2][row], b1 = S.mask[(sgm.half * 8 + cc) * 2 + 1][row];
                float4* o = reinterpret_cast<float4*>(P.part + (pr + tile) * P.H1 + sgm.half * 256 + cc * 32);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  float e[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const int ii = i + j;
                    const uint32_t bit = ii < 16 ? (b0 >> mask_pos(ii)) & 1u : (b1 >> mask_pos(ii - 16)) & 1u;
                    e[j] = bit ? __uint_as_float(rr[ii]) : 0.f;
                  }
                  o[i / 4] = make_float4(e[0], e[1], e[2], e[3]);
                }
              }
            }
          } else
          for (int64_t ps = p_first; ps <= p_last; ++ps) {
            const bool mine = pr == ps;
            const bool any_mine = __any_sync(0xffffffffu, mine);   // a warp with no row of this pair contributes exact zeros
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              const int cc = hq * 2 + c;
              if (!any_mine) { S.red[q][cc * 32 + lane] = 0.f; continue; }
              uint32_t rr[32];
              tmem_ld32(d_addr + (uint32_t)(cc * 32), rr);
              const uint32_t b0 = mine ? S.mask[(sgm.half * 8 + cc) * 2][row] : 0u, b1 = mine ? S.mask[(sgm.half * 8 + cc) * 2 + 1][row] : 0u;
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i)
                v[i] = (i < 16 ? (b0 >> mask_pos(i)) & 1u : (b1 >> mask_pos(i - 16)) & 1u) ? __uint_as_float(rr[i]) : 0.f;
              S.red[q][cc * 32 + lane] = warp_transpose_sum(v, lane);
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (tid < 256) {
              const float s = (S.red[0][tid] + S.red[1][tid]) + (S.red[2][tid] + S.red[3][tid]);
              const int col = sgm.half * 256 + tid;
              if (!P.acc_mode) {
                P.part[(ps + tile) * P.H1 + col] = s;
              } else {
                const int64_t ta = (ps * P.G) / TILE_M, tb = ((ps + 1) * P.G - 1) / TILE_M;    // tiles of pair ps
                if (ta < tile0) {                     // the pair started in an earlier CTA: hand the tile's sum to the fix-up
                  P.cont[((int64_t)blockIdx.x * P.max_cont + (tile - tile0)) * P.H1 + col] = s;
                } else {
                  const float a = tile == ta ? s : S.acc[col] + s;       // ((s_ta + s_ta+1) + ...) : ascending tiles
                  if (tile == tb) P.dUp[ps * P.H1 + col] = a * P.inv_gscale;
                  else if (t == tiles_mine - 1) P.prefix[(int64_t)blockIdx.x * P.H1 + col] = a;
                  S.acc[col] = a;
                }
              }
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
          }
          if (!last_seg) {                          // 3D: second N-half reuses the same A operand
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) signal_kb(kb);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NEPI + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// =============================================================================================
// tc_trunk2_kernel: the single-pass modes (bf16 / fp16) with TWO row tiles in flight per CTA.
//
// In the kernel above a tile's layers form a serial chain (accumulate -> epilogue -> next layer's MMAs), and with one
// MMA per product the chain (~1000 cycles + fill) is half of a layer's 2048 cycles of tensor work: 0.70 of roofline.
// A second, independent tile hides it, but does not fit in TMEM next to TS-mode operands (2 x (128 + 256) columns).
// Here TMEM holds only the two fp32 accumulators (2 x 256 columns); the activations go through shared memory (SS-mode
// MMAs): while the tensor pipe works on one tile, the epilogue warps convert the other tile's accumulator into the A
// k-blocks of its next segment (canonical K-major SWIZZLE_128B, the same image as a weight tile).  The ReLU sign bits
// no longer fit in shared memory (2 x 36 KB) and live in a per-CTA global scratch (L2 resident; a thread only ever
// reads back words it wrote itself).
//
// What sets the pace is the ISSUING THREAD, not the pipe or shared memory (scripts/microbench/mma_2cta.cu): an MMA is
// taken from the thread only when the pipe can start it, so the pipe is busy only while the thread sits at its next
// tcgen05.mma, and every barrier operation between two groups of four MMAs beyond ~65 cycles of slack is lost tensor
// time (try_wait ~50, tcgen05.commit ~55 cycles each: two waits + two polls + two commits per k-block -> 186 cycles
// per MMA instead of 128).  Hence ONE ring whose stage holds a k-block of BOTH operands: one wait and one commit per
// k-block.  A segment is exactly NS2 = 4 k-blocks, so stage index = k-block index and every barrier flips once per
// work item.
//
//   warp 16   producer: the weight tile of every (item, k-block) of the MMA work list -> stage[kb].W (cp.async.bulk)
//   warp 17   MMA issuer: walks the MMA work list (tile slot, segment); waits d_free[slot] before overwriting an
//             accumulator; per k-block waits full[kb], issues 4 tcgen05.mma (SS), commits empty[kb]; commits d_ready[slot]
//   warps 0-15 epilogue: walk the epilogue work list; per item wait d_ready[slot], tcgen05.ld the accumulator, release it
//             (d_free) as soon as the last column is in registers, convert, and write the A k-blocks of the tile's next
//             segment into stage[kb].A (wait empty[kb]; st.shared + fence.proxy.async; arrive full[kb]); sign bits to the
//             scratch.  full[kb] completes on the 16 epilogue arrivals + the producer's arrive.expect_tx + the bytes.
// The two lists (host, tc2_schedule) interleave the tiles so that both sides fill / drain the ring in the same order
// and no wait is circular.  3D: the two N-halves of the last backward GEMM read the same A operand; the second half's
// stages are the same physical stages, so the epilogue only re-arrives on them (the A part is left in place) while the
// producer refills the W part.
// =============================================================================================
struct Smem2 {
  uint64_t full[NS2_MAX], empty[NS2_MAX], d_ready[2], d_free[2];
  uint32_t tmem_base, pad_;
  alignas(16) float bias[7][256];
  alignas(16) float w_out[3][256];
  float b_out[4];
  float red[4][256];            // per lane-quadrant partial column sums
  float red_s[4];
};

#ifdef DGDM_TRUNK_TRACE
// CTA 0, second tile pair: issuer stamps at item*16 + k, epilogue warp 0 at 2048 + item*16 + k (scripts/dev/trunk2_timeline.py)
#define TR2(slot_) do { if (blockIdx.x == 0 && t == 1 && lane == 0 && (warp == 0 || warp == NEPI + 1)) g_trace[slot_] = clock64(); } while (0)
#else
#define TR2(slot_) do { } while (0)
#endif

template <bool F16, int NCTA>
__global__ void __launch_bounds__(NTHREADS, 1) tc_trunk2_kernel(const __grid_constant__ TcParams P) {
  using R2 = Ring2<NCTA>;
  constexpr int NS2 = R2::NS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Smem2& S = *reinterpret_cast<Smem2*>(ring + NS2 * R2::STAGE);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  uint32_t rank = 0;                              // CTA pair: 0 = leader (issues the MMAs), 1 = peer
  if (NCTA == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

  if (tid == 0) {
    // CTA pair: the leader's `full` / `d_free` barriers also take one arrival per phase from the peer's relay warp
    const uint32_t extra = (NCTA == 2 && rank == 0) ? 1u : 0u;
    for (int s = 0; s < NS2; ++s) { mbar_init(&S.full[s], NEPI + 1 + extra); mbar_init(&S.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&S.d_ready[s], 1); mbar_init(&S.d_free[s], NEPI + extra); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 7 * 256; i += NTHREADS) S.bias[i / 256][i % 256] = P.bias[i / 256][i % 256] * (F16 ? F16_SA : 1.f);
  for (int i = tid; i < 3 * 256; i += NTHREADS) S.w_out[i / 256][i % 256] = P.w_out[i];
  if (tid < 3) S.b_out[tid] = P.b_out[tid];
  if (warp == NEPI + 1) {
    if (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (NCTA == 2) cluster_sync_all();               // both CTAs' barriers exist before anything remote touches them
  tc_fence_after();
  if (S.tmem_base != 0u) { if (tid == 0 && P.err) atomicExch(P.err, 9); __trap(); }   // literal TMEM addresses below

  // both CTAs of a pair walk the same number of tile pairs (the leader's: it owns the lower block index); a CTA whose
  // tile index runs past the end processes dead rows
  const int tiles_mine = (P.n_tiles - ((int)blockIdx.x - (int)rank) + (int)gridDim.x - 1) / (int)gridDim.x;
  const int pairs_mine = (tiles_mine + 1) / 2;

  if (warp == NEPI) {
    // =============================== producer ===============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < pairs_mine; ++t) {
        const bool has1 = 2 * t + 1 < tiles_mine;
        for (int i = 0; i < P.n_m; ++i) {
          const uint32_t it = P.m_items[i];
          if ((it >> 7) && !has1) continue;
          const Seg sgm = P.seg[it & 0x7Fu];
          const uint32_t tile_bytes = (uint32_t)sgm.n_rows * 128u, my_bytes = tile_bytes / NCTA;   // CTA pair: my half of the rows
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&S.empty[stage], phase ^ 1, P.err, 1);
            mbar_arrive_expect_tx(&S.full[stage], my_bytes);
            // image order inside a segment: kb0.hi, kb0.lo, kb1.hi, ...; the single-pass modes read the hi parts
            bulk_g2s(ring + stage * R2::STAGE, P.img + sgm.img_off + (size_t)(kb * 2) * tile_bytes + (size_t)rank * my_bytes, my_bytes,
                     &S.full[stage]);
            if (++stage == NS2) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == NEPI + 1 && rank == 0) {
    // =============================== MMA issuer (whole warp, uniform; lane 0 issues) ===============================
    uint32_t stage = 0, phase = 0, df_ph = 0;
    for (int t = 0; t < pairs_mine; ++t) {
      const bool has1 = 2 * t + 1 < tiles_mine;
      for (int i = 0; i < P.n_m; ++i) {
        const uint32_t it = P.m_items[i];
        const uint32_t slot = it >> 7, sg = it & 0x7Fu;
        if (slot && !has1) continue;
        const Seg sgm = P.seg[sg];
        TR2(i * 16 + 0);
        if (!sgm.accum) {                        // this segment overwrites the accumulator: the epilogue must have read it
          mbar_wait(&S.d_free[slot], ((df_ph >> slot) & 1u) ^ 1u, P.err, 5);
          df_ph ^= 1u << slot;
        }
        TR2(i * 16 + 1);
        const uint32_t d_base = slot * 256u;
        const uint32_t idesc = make_idesc<F16>(sgm.n_rows, TILE_M * NCTA);
        uint32_t accum = sgm.accum;
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&S.full[stage], phase, P.err, 3);
          tc_fence_after();
          TR2(i * 16 + 3 + kb * 3);
          // descriptors of the four K = 16 steps: + 32 bytes = + 2 in the 14-bit address field (no carry: < 227 KB), so
          // one add per descriptor instead of a shift / mask / or chain on the uniform datapath before every MMA
          const uint32_t b_addr = smem_u32(ring + stage * R2::STAGE);
          const uint64_t bd0 = make_b_desc(b_addr), ad0 = make_b_desc(b_addr + R2::WPART);
#pragma unroll
          for (int ks = 0; ks < KBLK / 16; ++ks) {
            if (NCTA == 2) tc_mma_ss2(d_base, ad0 + (uint64_t)(2 * ks), bd0 + (uint64_t)(2 * ks), idesc, accum);
            else tc_mma_ss(d_base, ad0 + (uint64_t)(2 * ks), bd0 + (uint64_t)(2 * ks), idesc, accum);
            accum = 1;
          }
          if (NCTA == 2) tc_commit2(&S.empty[stage]); else tc_commit(&S.empty[stage]);
          TR2(i * 16 + 4 + kb * 3);
          if (++stage == NS2) { stage = 0; phase ^= 1; }
        }
        if (sgm.kind != K_MID) { if (NCTA == 2) tc_commit2(&S.d_ready[slot]); else tc_commit(&S.d_ready[slot]); }
        TR2(i * 16 + 14);
      }
    }
  } else if (warp == NEPI + 1) {
    // =============================== relay (peer CTA of a pair) ===============================
    // forwards this CTA's "stage full" and "accumulator read" events to the leader's barriers, in the issuer's order
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, df_ph = 0, df_seen = 0;
      for (int t = 0; t < pairs_mine; ++t) {
        const bool has1 = 2 * t + 1 < tiles_mine;
        for (int i = 0; i < P.n_m; ++i) {
          const uint32_t it = P.m_items[i];
          const uint32_t slot = it >> 7, sg = it & 0x7Fu;
          if (slot && !has1) continue;
          const Seg sgm = P.seg[sg];
          if (!sgm.accum) {
            if ((df_seen >> slot) & 1u) {           // (the first use of an accumulator waits for nobody, here and in the leader)
              mbar_wait(&S.d_free[slot], (df_ph >> slot) & 1u, P.err, 7);
              df_ph ^= 1u << slot;
              mbar_arrive_remote(&S.d_free[slot], 0);
            }
            df_seen |= 1u << slot;
          }
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&S.full[stage], phase, P.err, 8);
            mbar_arrive_remote(&S.full[stage], 0);
            if (++stage == NS2) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // =============================== epilogue warps 0..15 ===============================
    const int q = warp & 3, hq = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int l1_hw = P.H1 / 16;
    uint32_t est = 0, eph = 0, dr_ph = 0, ekb = 0, est0 = 0;   // next stage to fill + its phase; k-block within the item; item's first stage
    uint16_t* const mask_cta = P.mask_scratch + (size_t)blockIdx.x * 2 * (2 * MASK_WORDS) * TILE_M + row;

    // this warp's 32 rows x 16 features of the next A k-block -> stage[est].A (canonical K-major SWIZZLE_128B);
    // `write` = false (one-CTA form only: 4 stages = one segment): the operand is already there (3D, second N-half of
    // the last GEMM, same physical stages): only hand the stages over again
    [[maybe_unused]] int tr_i = 0, tr_t = 0, tr_kb = 0;
    auto put_kb = [&](const uint32_t (&hi)[8], bool write) {
      mbar_wait(&S.empty[est], eph ^ 1u, P.err, 6);
#ifdef DGDM_TRUNK_TRACE
      { const int t = tr_t; TR2(2048 + tr_i * 16 + 8 + (tr_kb & 3)); }
#endif
      if (ekb == 0) est0 = est;
      if (write) {
        uint8_t* dst = ring + est * R2::STAGE + R2::WPART + row * 128;
        const int sw = row & 7;
        *reinterpret_cast<uint4*>(dst + (((2 * hq) ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + (((2 * hq + 1) ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      }
      // One proxy fence per ITEM (after the fourth k-block), then the four arrivals: fence.proxy.async costs the warp ~400
      // cycles, four of them per item made the epilogue the slowest role.  Tried instead: st.async (STAS, needs a cluster
      // launch) with transaction-count completion and no fence: 6600 cycles per item; the fence from lane 0 only, per
      // k-block: no gain.
      if (ekb == 3) {
        if (write) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          uint32_t st = est0;
#pragma unroll
          for (int k = 0; k < 4; ++k) { mbar_arrive(&S.full[st]); if (++st == NS2) st = 0; }
        }
      }
#ifdef DGDM_TRUNK_TRACE
      { const int t = tr_t; TR2(2048 + tr_i * 16 + 2 + (tr_kb & 3)); ++tr_kb; }
#endif
      ekb = (ekb + 1) & 3;
      if (++est == NS2) { est = 0; eph ^= 1u; }
    };

    for (int t = 0; t < pairs_mine; ++t) {
      const bool has1 = 2 * t + 1 < tiles_mine;
      for (int i = 0; i < P.n_e; ++i) {
        const uint32_t it = P.e_items[i];
        const uint32_t slot = it >> 7;
        const int code = (int)(it & 0x7Fu);
        if (slot && !has1) continue;
        tr_i = i; tr_t = t; tr_kb = 0;
        TR2(2048 + i * 16 + 0);
        const int tile = (int)blockIdx.x + (2 * t + (int)slot) * (int)gridDim.x;
        uint16_t* const mk = mask_cta + (size_t)slot * (2 * MASK_WORDS) * TILE_M;     // word w of this row: mk[w * TILE_M]

        // Row -> (pair, pose row): one 64-bit division per item that needs it, everything per row is 32-bit.
        const int64_t base = (int64_t)tile * TILE_M;
        const int64_t r_glob = base + row;
        const bool live = r_glob < P.n_rows;

        if (code >= IT_A1) {
          // ---- layer-1 operand, K-half kh: relu(Cst[obj] + U[design] + V[g]) for the 64 features this warp owns ----
          const int kh = code - IT_A1;
          const int64_t p_first = base / P.G;
          const uint32_t rem0 = (uint32_t)(base - p_first * P.G), Gu = (uint32_t)P.G;
          const uint32_t rofs = rem0 + (uint32_t)row, dq = rofs / Gu;
          const int64_t pr = live ? p_first + dq : -1;
          const int g = live ? (int)(rofs - dq * Gu) : 0;
          const float* u_row = nullptr; const float* c_row = nullptr;
          if (live) {
            const int64_t design = P.opd == 1 ? pr : (pr <= 0x7fffffffll ? (int64_t)((uint32_t)pr / (uint32_t)P.opd) : pr / P.opd);
            u_row = P.U + design * P.H1;
            c_row = P.Cst + (int64_t)pair_obj(pr, P.opd, P.n_designs, P.n_obj, P.pair_object) * P.H1;
          }
          auto load_z = [&](int col0, float (&z)[16]) {
            float4 pv[4];
            const char* vb = reinterpret_cast<const char*>(P.Vt);
            const uint32_t step = (uint32_t)P.G * 16u;
            uint32_t off = (uint32_t)(col0 >> 2) * step + (uint32_t)g * 16u;
#pragma unroll
            for (int k = 0; k < 4; ++k, off += step) pv[k] = live ? __ldg(reinterpret_cast<const float4*>(vb + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              if (live) {
                const float4 k4 = *reinterpret_cast<const float4*>(c_row + col0 + k);
                const float4 u4 = *reinterpret_cast<const float4*>(u_row + col0 + k);
                const float4 p4 = pv[k >> 2];
                a.x = (k4.x + u4.x) + p4.x; a.y = (k4.y + u4.y) + p4.y; a.z = (k4.z + u4.z) + p4.z; a.w = (k4.w + u4.w) + p4.w;
              }
              if (F16) { a.x *= F16_SA; a.y *= F16_SA; a.z *= F16_SA; a.w *= F16_SA; }
              z[k] = a.x; z[k + 1] = a.y; z[k + 2] = a.z; z[k + 3] = a.w;
            }
          };
          float zb[2][16];
          load_z(kh * 256 + hq * 16, zb[0]);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            if (kb + 1 < 4) load_z(kh * 256 + (kb + 1) * 64 + hq * 16, zb[(kb + 1) & 1]);
            uint32_t hi[8], lo[8];
            relu_split16<false, F16>(zb[kb & 1], hi, lo);
            put_kb(hi, true);
            __stcg(mk + (kh * 16 + kb * 4 + hq) * TILE_M, (uint16_t)sign_bits16(hi));
          }
          TR2(2048 + i * 16 + 6);
          continue;
        }

        const Seg sgm = P.seg[code];
        const uint32_t d_addr = lane_addr + slot * 256u;
        // the accumulator may be overwritten once every epilogue warp has its columns in registers
        auto release_d = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.d_free[slot]);
        };
        auto wait_d = [&]() {
          mbar_wait(&S.d_ready[slot], (dr_ph >> slot) & 1u, P.err, 4);
          dr_ph ^= 1u << slot;
          tc_fence_after();
        };

        if (sgm.kind == K_FWD) {
          // a = relu(D + b) -> next A operand; sign bits of this layer
          const int mbase = l1_hw + (sgm.layer - 1) * 16;
          wait_d();
          TR2(2048 + i * 16 + 1);
          uint32_t rr[2][16];
          tmem_ld16_async(d_addr + (uint32_t)(hq * 16), rr[0]);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            tmem_ld_wait16(rr[kb & 1]);
            if (kb + 1 < 4) tmem_ld16_async(d_addr + (uint32_t)((kb + 1) * 64 + hq * 16), rr[(kb + 1) & 1]);
            else release_d();
            const float4* b4 = reinterpret_cast<const float4*>(&S.bias[sgm.layer - 1][kb * 64 + hq * 16]);
            float z[16];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 bb = b4[i4];
              z[i4 * 4 + 0] = unscale_add<F16, false>(rr[kb & 1][i4 * 4 + 0], bb.x);
              z[i4 * 4 + 1] = unscale_add<F16, false>(rr[kb & 1][i4 * 4 + 1], bb.y);
              z[i4 * 4 + 2] = unscale_add<F16, false>(rr[kb & 1][i4 * 4 + 2], bb.z);
              z[i4 * 4 + 3] = unscale_add<F16, false>(rr[kb & 1][i4 * 4 + 3], bb.w);
            }
            uint32_t hi[8], lo[8];
            relu_split16<false, F16>(z, hi, lo);
            put_kb(hi, true);
            __stcg(mk + (mbase + kb * 4 + hq) * TILE_M, (uint16_t)sign_bits16(hi));
          }
          TR2(2048 + i * 16 + 6);
        } else if (sgm.kind == K_BWD) {
          // d = D * 1[a_layer > 0] -> next A operand
          const int mbase = sgm.layer == 0 ? 0 : l1_hw + (sgm.layer - 1) * 16;
          uint32_t bits[4];
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) bits[kb] = __ldcg(mk + (mbase + kb * 4 + hq) * TILE_M);   // in flight during the wait
          wait_d();
          TR2(2048 + i * 16 + 1);
          uint32_t rr[2][16];
          tmem_ld16_async(d_addr + (uint32_t)(hq * 16), rr[0]);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            tmem_ld_wait16(rr[kb & 1]);
            if (kb + 1 < 4) tmem_ld16_async(d_addr + (uint32_t)((kb + 1) * 64 + hq * 16), rr[(kb + 1) & 1]);
            else release_d();
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = (bits[kb] >> mask_pos(k)) & 1u ? unscale<F16, false>(rr[kb & 1][k]) : 0.f;
            uint32_t hi[8], lo[8];
            split_pack<false, F16>(v, hi, lo);
            put_kb(hi, true);
          }
          TR2(2048 + i * 16 + 6);
        } else {
          // K_OUT and K_LAST need the row's pair
          const int64_t p_first = base / P.G;
          const uint32_t rem0 = (uint32_t)(base - p_first * P.G), Gu = (uint32_t)P.G;
          const uint32_t rofs = rem0 + (uint32_t)row, dq = rofs / Gu;
          const int64_t pr = live ? p_first + dq : -1;
          int64_t last_row = base + TILE_M - 1;
          if (last_row >= P.n_rows) last_row = P.n_rows - 1;
          const int64_t p_last = base < P.n_rows ? p_first + (rem0 + (uint32_t)(last_row - base)) / Gu : p_first - 1;   // past the end: no pairs
          if (sgm.kind == K_OUT) {
            const float coef = (live && P.obj.row_coef) ? P.obj.row_coef[r_glob] : 1.f;
            const int mbase = l1_hw + 6 * 16;
            uint32_t bits[4] = {0u, 0u, 0u, 0u};
            if (P.backward) {
#pragma unroll
              for (int kb = 0; kb < 4; ++kb) bits[kb] = __ldcg(mk + (mbase + kb * 4 + hq) * TILE_M);
            }
            wait_d();
            TR2(2048 + i * 16 + 1);
            uint32_t rr[8];
            tmem_ld8(d_addr, rr);
            release_d();
            constexpr float inv_out = F16 ? 1.f / F16_SA : 1.f;
            const float l0 = __uint_as_float(rr[0]) * inv_out + S.b_out[0], l1 = __uint_as_float(rr[1]) * inv_out + S.b_out[1],
                        l2 = __uint_as_float(rr[2]) * inv_out + S.b_out[2];
            if (hq == 0 && live && P.logits) {
              float* o = P.logits + r_glob * 3;
              o[0] = l0; o[1] = l1; o[2] = l2;
            }
            if (!P.backward) {
              const float val = live ? coef * (P.obj.c[0] * l0 + P.obj.c[1] * l1 + P.obj.c[2] * l2 + P.obj.sq0 * l0 * l0) : 0.f;
              if (P.G == 1) {
                if (hq == 0 && live) P.score_part[pr + tile] = val;
              } else
              for (int64_t ps = p_first; ps <= p_last; ++ps) {
                float s = (pr == ps) ? val : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (hq == 0 && lane == 0) S.red_s[q] = s;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                if (tid == 0) P.score_part[ps + tile] = (S.red_s[0] + S.red_s[1]) + (S.red_s[2] + S.red_s[3]);
                asm volatile("bar.sync 1, 512;" ::: "memory");
              }
            } else {
              // seed of the backward pass: d8 = (dObj/dlogits . W_out) * 1[a_8 > 0]
              const float cg = coef * P.gscale;
              const float dl0 = cg * (P.obj.c[0] + 2.f * P.obj.sq0 * l0), dl1 = cg * P.obj.c[1], dl2 = cg * P.obj.c[2];
#pragma unroll 1
              for (int kb = 0; kb < 4; ++kb) {
                const int col0 = kb * 64 + hq * 16;
                float v[16];
                const float4* w0 = reinterpret_cast<const float4*>(&S.w_out[0][col0]);
                const float4* w1 = reinterpret_cast<const float4*>(&S.w_out[1][col0]);
                const float4* w2 = reinterpret_cast<const float4*>(&S.w_out[2][col0]);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const float4 a0 = w0[i4], a1 = w1[i4], a2 = w2[i4];
                  const float d[4] = {dl0 * a0.x + dl1 * a1.x + dl2 * a2.x, dl0 * a0.y + dl1 * a1.y + dl2 * a2.y,
                                      dl0 * a0.z + dl1 * a1.z + dl2 * a2.z, dl0 * a0.w + dl1 * a1.w + dl2 * a2.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[i4 * 4 + e] = (bits[kb] >> mask_pos(i4 * 4 + e)) & 1u ? d[e] : 0.f;
                }
                uint32_t hi[8], lo[8];
                split_pack<false, F16>(v, hi, lo);
                put_kb(hi, true);
              }
            }
            TR2(2048 + i * 16 + 6);
          } else {
            // ---- K_LAST: d1 = D * 1[a_1 > 0], summed over the rows of each pair present in the tile (K2) ----
            // warp (q, hq) reduces the 4 x 16 columns whose layer-1 sign bits it wrote itself: two 32-value chunks
            uint32_t bits[4];
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) bits[kb] = __ldcg(mk + (sgm.half * 16 + kb * 4 + hq) * TILE_M);
            wait_d();
            TR2(2048 + i * 16 + 1);
            if (code != P.n_seg - 1) {                // 3D, first N-half: hand the same operand over for the second half
              const uint32_t none[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll 1
              for (int kb = 0; kb < 4; ++kb) put_kb(none, false);
            }
            if (P.G == 1) {                           // explicit-row mode: the row IS the pair; store it directly
#pragma unroll 1
              for (int kb = 0; kb < 4; ++kb) {
                uint32_t rr[16];
                tmem_ld16_async(d_addr + (uint32_t)(kb * 64 + hq * 16), rr);
                tmem_ld_wait16(rr);
                if (live) {
                  float4* o = reinterpret_cast<float4*>(P.part + (pr + tile) * P.H1 + sgm.half * 256 + kb * 64 + hq * 16);
#pragma unroll
                  for (int k = 0; k < 16; k += 4) {
                    float e[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) e[j] = (bits[kb] >> mask_pos(k + j)) & 1u ? __uint_as_float(rr[k + j]) : 0.f;
                    o[k / 4] = make_float4(e[0], e[1], e[2], e[3]);
                  }
                }
              }
              release_d();
            } else
            for (int64_t ps = p_first; ps <= p_last; ++ps) {
              const bool mine = pr == ps;
#pragma unroll 1
              for (int c = 0; c < 2; ++c) {
                uint32_t r0[16], r1[16];
                tmem_ld16_async(d_addr + (uint32_t)((2 * c) * 64 + hq * 16), r0);
                tmem_ld16_async(d_addr + (uint32_t)((2 * c + 1) * 64 + hq * 16), r1);
                tmem_ld_wait16(r0);
                tmem_ld_wait16(r1);
                if (ps == p_last && c == 1) release_d();
                const uint32_t b0 = mine ? bits[2 * c] : 0u, b1 = mine ? bits[2 * c + 1] : 0u;
                float v[32];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  v[k] = (b0 >> mask_pos(k)) & 1u ? __uint_as_float(r0[k]) : 0.f;
                  v[16 + k] = (b1 >> mask_pos(k)) & 1u ? __uint_as_float(r1[k]) : 0.f;
                }
                const float sum = warp_transpose_sum(v, lane);
                S.red[q][(2 * c + (lane >> 4)) * 64 + hq * 16 + (lane & 15)] = sum;
              }
              asm volatile("bar.sync 1, 512;" ::: "memory");
              if (tid < 256) {
                const float s = (S.red[0][tid] + S.red[1][tid]) + (S.red[2][tid] + S.red[3][tid]);
                P.part[(ps + tile) * P.H1 + sgm.half * 256 + tid] = s;
              }
              asm volatile("bar.sync 1, 512;" ::: "memory");
            }
            TR2(2048 + i * 16 + 6);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (NCTA == 2) cluster_sync_all();               // the pair's MMAs touch both CTAs' TMEM and shared memory
  if (warp == NEPI + 1) {
    tc_fence_after();
    if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(0u) : "memory");
  }
}

// K2 tail: dUp[p,:] = sum over the tiles covering pair p of part[p + tile,:], ascending tile order.
__global__ void reduce_slots_kernel(float* __restrict__ dUp, const float* __restrict__ part, int64_t n_pairs, int G, int H1,
                                    float inv_gscale) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pairs * H1) return;
  int64_t p = idx / H1;
  int c = (int)(idx % H1);
  int64_t t0 = (p * G) / TILE_M, t1 = ((p + 1) * G - 1) / TILE_M;
  float s = 0.f;
  for (int64_t t = t0; t <= t1; ++t) s += part[(p + t) * H1 + c];
  dUp[idx] = s * inv_gscale;          // exact: the scale is a power of two
}
// acc_mode tail: a pair cut by a CTA boundary = the starting CTA's prefix + the following tiles' sums, ascending.
__global__ void fixup_pairs_kernel(float* __restrict__ dUp, const float* __restrict__ prefix, const float* __restrict__ cont,
                                   int n_tiles, int grid, int G, int H1, int max_cont, int64_t n_rows, float inv_gscale) {
  const int b = blockIdx.x;
  int tile0, count;
  cta_tiles(n_tiles, grid, b, tile0, count);
  if (count == 0) return;
  const int last_tile = tile0 + count - 1;
  int64_t last_row = (int64_t)last_tile * TILE_M + TILE_M - 1;
  if (last_row >= n_rows) last_row = n_rows - 1;
  const int64_t p = last_row / G;
  const int64_t ta = (p * G) / TILE_M, tb = ((p + 1) * G - 1) / TILE_M;
  if (tb <= last_tile || ta < tile0) return;       // completed inside the CTA / started in an earlier one (its block does it)
  for (int c = threadIdx.x; c < H1; c += blockDim.x) {
    float s = prefix[(int64_t)b * H1 + c];
    for (int64_t tile = last_tile + 1; tile <= tb; ++tile) {
      const int o = tile_owner(n_tiles, grid, (int)tile);
      int o0, oc;
      cta_tiles(n_tiles, grid, o, o0, oc);
      s += cont[((int64_t)o * max_cont + (tile - o0)) * H1 + c];
    }
    dUp[p * H1 + c] = s * inv_gscale;
  }
}
__global__ void reduce_score_slots_kernel(float* __restrict__ out, const float* __restrict__ part, int64_t n_pairs, int G) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  int64_t t0 = (p * G) / TILE_M, t1 = ((p + 1) * G - 1) / TILE_M;
  float s = 0.f;
  for (int64_t t = t0; t <= t1; ++t) s += part[p + t];
  out[p] = s;
}

// ---------------------------------------------------------------------------------------------
// weight image
// ---------------------------------------------------------------------------------------------
struct PackSeg { const float* src; int ld; int k0; int n_valid; int n_rows; uint32_t img_off; };

// one thread per 16-byte chunk (8 elements) of a tile; image order inside a segment: kb0.hi, kb0.lo, kb1.hi, ...
// `f16`: fp16 hi/lo instead of bf16 hi/lo.
__global__ void pack_tc_kernel(uint8_t* __restrict__ img, PackSeg ps, int f16, float scale) {
  const int tile_bytes = ps.n_rows * 128;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunks_per_tile = ps.n_rows * 8;
  if (idx >= 8 * chunks_per_tile) return;               // 4 k-blocks x (hi, lo)
  const int tl = idx / chunks_per_tile, rem = idx % chunks_per_tile;
  const int kb = tl >> 1, part = tl & 1;
  const int n = rem / 8, j = rem % 8;
  uint16_t out[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float w = n < ps.n_valid ? ps.src[(int64_t)n * ps.ld + ps.k0 + kb * KBLK + j * 8 + e] : 0.f;
    if (f16) {
      w *= scale;
      __half hi = __float2half_rn(w);
      __half v = part == 0 ? hi : __float2half_rn(w - __half2float(hi));
      out[e] = *reinterpret_cast<uint16_t*>(&v);
    } else {
      __nv_bfloat16 hi = __float2bfloat16_rn(w);
      __nv_bfloat16 v = part == 0 ? hi : __float2bfloat16_rn(w - __bfloat162float(hi));
      out[e] = *reinterpret_cast<uint16_t*>(&v);
    }
  }
  // 128B swizzle: 16-byte chunk j of row n lands at chunk (j ^ (n & 7))
  uint8_t* dst = img + ps.img_off + (size_t)tl * tile_bytes + (size_t)n * 128 + ((j ^ (n & 7)) * 16);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out);
}

struct Plan { int n_seg; Seg seg[MAX_SEG]; PackSeg pack[MAX_SEG]; size_t bytes; };

// Segment list of one tile (see the kernel header).  H1 = 256: 15 segments; H1 = 512: 17.
Plan make_plan(const dgdm_dyn_weights* w, int H1) {
  Plan pl{};
  size_t off = 0;
  auto add = [&](int kind, int layer, int half, int accum, const float* src, int ld, int k0, int n_valid, int n_rows) {
    Seg& s = pl.seg[pl.n_seg];
    s.img_off = (uint32_t)off; s.n_rows = (uint16_t)n_rows; s.kind = (uint8_t)kind; s.layer = (uint8_t)layer;
    s.half = (uint8_t)half; s.accum = (uint8_t)accum;
    pl.pack[pl.n_seg] = PackSeg{src, ld, k0, n_valid, n_rows, (uint32_t)off};
    off += (size_t)8 * n_rows * 128;
    ++pl.n_seg;
  };
  // forward, trunk layer index 1..7 = reference layers 2..8 (mask row `layer`, bias `layer-1`)
  if (H1 == 512) {
    add(K_MID, 1, 0, 0, w ? w->wl[0] : nullptr, 512, 0, 256, 256);
    add(K_FWD, 1, 0, 1, w ? w->wl[0] : nullptr, 512, 256, 256, 256);
  } else {
    add(K_FWD, 1, 0, 0, w ? w->wl[0] : nullptr, 256, 0, 256, 256);
  }
  for (int l = 2; l <= 7; ++l) add(K_FWD, l, 0, 0, w ? w->wl[l - 1] : nullptr, 256, 0, 256, 256);
  add(K_OUT, 0, 0, 0, w ? w->w_out : nullptr, 256, 0, 3, 16);
  // backward: B operand = W_l^T rows; after layer index l the mask of layer l-1 applies
  for (int l = 7; l >= 2; --l) add(K_BWD, l - 1, 0, 0, w ? w->wl_t[l - 1] : nullptr, 256, 0, 256, 256);
  add(K_LAST, 0, 0, 0, w ? w->wl_t[0] : nullptr, 256, 0, 256, 256);
  if (H1 == 512) add(K_LAST, 0, 1, 0, w ? w->wl_t[0] + 256 * 256 : nullptr, 256, 0, 256, 256);
  pl.bytes = off;
  return pl;
}

size_t smem_bytes() { return 1024 + (size_t)NSTAGE * WTILE_BYTES + sizeof(Smem); }
template <int NCTA> size_t smem2_bytes() { return 1024 + (size_t)Ring2<NCTA>::NS * Ring2<NCTA>::STAGE + sizeof(Smem2); }
constexpr size_t MASK_SCRATCH_BYTES = (size_t)MAX_CTAS2 * 2 * (2 * MASK_WORDS) * TILE_M * sizeof(uint16_t);

// DGDM_TRUNK2=1 runs the single-pass modes on the two-tile kernel, DGDM_TRUNK2=2 on its CTA-pair form (cta_group::2:
// each CTA stages half of every weight tile, which makes room for a ring of one and a half segments).  Off by default:
// see DESIGN.md 4.1 for the measurements.
int trunk2_mode() {
  static const int mode = [] { const char* e = getenv("DGDM_TRUNK2"); return (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0; }();
  return mode;
}
bool use_trunk2() { return trunk2_mode() != 0; }

// Work lists of the two-tile kernel for one pair of tiles (slot 0 = X, slot 1 = Y).  The epilogue list is the
// interleave E(X,sg), E(Y,sg); the MMA list is derived from it so that the issuer consumes the A ring in exactly the
// order the epilogue fills it: an item that produces an operand (A1, K_FWD, K_BWD, K_OUT when a backward follows)
// enables the next segment of its tile.  3D backward tail: the two N-halves of the last GEMM share one operand, which
// stays in the ring until the second half has run -- and the second half waits for the first half's epilogue -- so
// X's first half is reduced BEFORE Y's operand is produced (with the plain interleave Y's production would wait for
// ring slots X cannot release: a circular wait with a 5-slot ring).
void tc2_schedule(TcParams& P) {
  int ne = 0, nm = 0;
  auto E = [&](int slot, int code) { P.e_items[ne++] = (uint8_t)((slot << 7) | code); };
  const bool two_half = P.H1 == 512;
  const int n = P.n_seg, first = two_half ? 1 : 0;
  const bool tail3d = two_half && P.backward;
  for (int slot = 0; slot < 2; ++slot) {
    E(slot, IT_A1);
    if (two_half) E(slot, IT_A1 + 1);
  }
  for (int sg = first; sg < (tail3d ? n - 3 : n); ++sg) { E(0, sg); E(1, sg); }
  if (tail3d) { E(0, n - 3); E(0, n - 2); E(1, n - 3); E(0, n - 1); E(1, n - 2); E(1, n - 1); }
  for (int i = 0; i < ne; ++i) {
    const int slot = P.e_items[i] >> 7, code = P.e_items[i] & 0x7F;
    int next = -1;
    if (code >= IT_A1) next = code - IT_A1;
    else {
      const int kind = P.seg[code].kind;
      if (kind == K_FWD || kind == K_BWD || (kind == K_OUT && P.backward) || (kind == K_LAST && code != n - 1)) next = code + 1;
    }
    if (next >= 0) P.m_items[nm++] = (uint8_t)((slot << 7) | next);
  }
  P.n_e = ne; P.n_m = nm;
}

// event-pair timing of the trunk kernel (bench.py roofline)
// Off by default.  Launches made while the stream is being captured into a CUDA graph are not timed (events cannot
// be recorded into a capture and read back per replay); the events are destroyed when timing is switched off.
struct Timing {
  std::mutex mu;
  bool on = false;
  std::vector<cudaEvent_t> ev;   // start, stop, start, stop, ...
  std::vector<uint8_t> fwd_only; // per event pair: 1 = scoring launch (forward only), 0 = guidance launch (fwd + dgrad)
  std::vector<int64_t> pair_rows;
  size_t used = 0;
  int64_t rows = 0;
};
Timing g_timing;

}  // namespace

// DGDM_TRUNK_SLOTS=1 keeps the round-1 reduction (one H1-wide slot per (pair, tile) + reduce_slots_kernel) for A/B tests
static bool use_slots() {
  static const bool on = [] { const char* e = getenv("DGDM_TRUNK_SLOTS"); return e && e[0] == '1'; }();
  return on;
}
constexpr int MAX_GRID = 160;                  // acc_mode buffers are sized for this many CTAs
static bool acc_mode_for(int G) { return G >= TILE_M && !use_slots() && !use_trunk2(); }
static int max_cont_for(int G) { return G / TILE_M + 2; }     // tiles a pair can spill into later CTAs (bounded per CTA below)

size_t tc_trunk_workspace_bytes(int H1, int64_t n_pairs, int G) {
  int64_t n_tiles = (n_pairs * G + TILE_M - 1) / TILE_M;
  int64_t slots = n_pairs + n_tiles;
  const size_t sums = acc_mode_for(G) ? align_up((size_t)MAX_GRID * (1 + max_cont_for(G)) * H1 * sizeof(float), 256)
                                      : align_up((size_t)slots * H1 * sizeof(float), 256);
  return sums + align_up((size_t)slots * sizeof(float), 256) + 512 + align_up(MASK_SCRATCH_BYTES, 256);
}

int tc_trunk(const dgdm_dyn_weights* w, const float* U, const float* Cst, const float* V, int n_designs, int n_obj,
             int opd, const int32_t* pair_object, int G, const dgdm_objective* obj, bool backward, float* dUp,
             float* score_sum, float* logits, void* ws, size_t ws_bytes, int precision, cudaStream_t s) {
  const int H1 = w->H1;
  const int64_t n_pairs = (int64_t)n_designs * opd;
  const int64_t n_rows = n_pairs * G;
  const int64_t n_tiles = (n_rows + TILE_M - 1) / TILE_M;
  DGDM_CHECK_ARG(n_tiles < (1ll << 30), "tc_trunk: too many tiles");
  const bool acc_mode = acc_mode_for(G) && backward;
  const int max_cont = max_cont_for(G);
  Arena ar(ws, ws_bytes);
  float* score_part = ar.take<float>((size_t)(n_pairs + n_tiles));
  int* err = ar.take<int>(1);
  uint16_t* mask_scratch = ar.take<uint16_t>(MASK_SCRATCH_BYTES / sizeof(uint16_t));
  float *part = nullptr, *prefix = nullptr, *cont = nullptr;
  if (acc_mode) {
    prefix = ar.take<float>((size_t)MAX_GRID * H1);
    cont = ar.take<float>((size_t)MAX_GRID * max_cont * H1);
  } else if (backward) {
    part = ar.take<float>((size_t)(n_pairs + n_tiles) * H1);
  }
  if (!ar.ok) { set_error("tc_trunk: workspace too small"); return DGDM_EWORKSPACE; }

  // per-device one-time setup (one process normally drives one GPU, but do not assume it)
  static int sm_counts[64] = {0};
  int dev = 0;
  DGDM_CUDA(cudaGetDevice(&dev));
  DGDM_CHECK_ARG(dev >= 0 && dev < 64, "tc_trunk: device ordinal %d out of range", dev);
  if (sm_counts[dev] == 0) {
    int n = 0;
    DGDM_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    DGDM_CUDA(cudaFuncSetAttribute(tc_trunk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    DGDM_CUDA(cudaFuncSetAttribute(tc_trunk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    DGDM_CUDA(cudaFuncSetAttribute(tc_trunk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    DGDM_CUDA(cudaFuncSetAttribute(tc_trunk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    DGDM_CUDA(cudaFuncSetAttribute(tc_trunk2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2_bytes<1>()));
    DGDM_CUDA(cudaFuncSetAttribute(tc_trunk2_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2_bytes<1>()));
    DGDM_CUDA(cudaFuncSetAttribute(tc_trunk2_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2_bytes<2>()));
    DGDM_CUDA(cudaFuncSetAttribute(tc_trunk2_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2_bytes<2>()));
    sm_counts[dev] = n;
  }
  const int sm_count = sm_counts[dev];
  const Plan pl = make_plan(w, H1);
  TcParams P{};
  const bool f16 = precision == DGDM_PREC_FP16 || precision == DGDM_PREC_FP16X3;
  // the image buffer holds three images (dgdm_dyn_pack_tc): bf16 hi+lo, fp16 hi+lo scaled by F16_SW (fp16x3), and the
  // unscaled fp16 image of the single-pass fp16 mode
  const bool x3_mode = precision == DGDM_PREC_BF16X3 || precision == DGDM_PREC_FP16X3;
  P.img = (const uint8_t*)w->tc_image + (f16 ? (x3_mode ? pl.bytes : 2 * pl.bytes) : 0);
  P.gscale = f16 ? F16_GS : 1.f;
  DGDM_CHECK_ARG((int64_t)H1 * G * 4 < (1ll << 32), "tc_trunk: pose table of %d x %d exceeds 32-bit byte offsets", H1, G);
  P.U = U; P.Cst = Cst; P.Vt = V;
  for (int i = 0; i < 7; ++i) P.bias[i] = w->bl[i];
  P.w_out = w->w_out; P.b_out = w->b_out;
  P.pair_object = pair_object;
  P.part = part; P.score_part = score_part; P.logits = logits; P.err = err;
  P.obj = *obj;
  P.n_rows = n_rows; P.n_tiles = (int)n_tiles; P.G = G; P.H1 = H1; P.opd = opd; P.n_designs = n_designs; P.n_obj = n_obj;
  P.x3 = precision == DGDM_PREC_BF16X3 || precision == DGDM_PREC_FP16X3; P.backward = backward;
  // Segment list of this launch.  The image always holds the output-layer tile; the fp32-grade plan does not use
  // it: the last hidden layer becomes K_FOUT and the K_OUT segment is dropped.  Forward only stops after the output.
  int n_seg = 0;
  for (int i = 0; i < pl.n_seg; ++i) {
    Seg sg = pl.seg[i];
    const bool is_out = sg.kind == K_OUT;
    if (P.x3 && is_out) { P.seg[n_seg - 1].kind = K_FOUT; }
    else P.seg[n_seg++] = sg;
    if (is_out && !backward) break;
  }
  P.n_seg = n_seg;
  P.alt = (P.x3 && backward) ? 1 : 0;
  const bool two_tile = !P.x3 && use_trunk2();
  P.mask_scratch = mask_scratch;
  if (two_tile) tc2_schedule(P);
  {
    static const bool merge = [] { const char* e = getenv("DGDM_TRUNK_MERGE"); return !(e && e[0] == '0'); }();
    P.merge_ab = (!P.x3 && merge) ? 1 : 0;
  }
  P.acc_mode = acc_mode; P.max_cont = max_cont; P.dUp = dUp; P.prefix = prefix; P.cont = cont;
  P.inv_gscale = f16 ? 1.f / (F16_GS * (P.x3 ? F16_SW : 1.f)) : 1.f;

  DGDM_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s));
  int grid = (int)(n_tiles < sm_count ? n_tiles : sm_count);
  if (two_tile && grid > MAX_CTAS2) grid = MAX_CTAS2;
  if (acc_mode && grid > MAX_GRID) grid = MAX_GRID;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  std::unique_lock<std::mutex> timing_lock(g_timing.mu);
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (g_timing.on) DGDM_CUDA(cudaStreamIsCapturing(s, &cap));
  if (g_timing.on && cap == cudaStreamCaptureStatusNone) {
    if (g_timing.used + 2 > g_timing.ev.size()) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; DGDM_CUDA(cudaEventCreate(&e)); g_timing.ev.push_back(e); }
    }
    e0 = g_timing.ev[g_timing.used]; e1 = g_timing.ev[g_timing.used + 1];
    g_timing.used += 2; g_timing.rows += n_rows;
    g_timing.fwd_only.push_back(backward ? 0 : 1);
    g_timing.pair_rows.push_back(n_rows);
    DGDM_CUDA(cudaEventRecord(e0, s));
  }
  // CTA-pair form: not for the 3D backward launch (its last GEMM re-uses an operand in place, which needs the
  // four-stage ring of the one-CTA form) and not for a single tile
  const bool pair = two_tile && trunk2_mode() == 2 && !(H1 == 512 && backward) && grid >= 2;
  if (pair) {
    grid &= ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem2_bytes<2>(); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (f16) DGDM_CUDA(cudaLaunchKernelEx(&cfg, tc_trunk2_kernel<true, 2>, P));
    else DGDM_CUDA(cudaLaunchKernelEx(&cfg, tc_trunk2_kernel<false, 2>, P));
  }
  else if (two_tile && f16) tc_trunk2_kernel<true, 1><<<grid, NTHREADS, smem2_bytes<1>(), s>>>(P);
  else if (two_tile) tc_trunk2_kernel<false, 1><<<grid, NTHREADS, smem2_bytes<1>(), s>>>(P);
  else if (P.x3 && f16) tc_trunk_kernel<true, true><<<grid, NTHREADS, smem_bytes(), s>>>(P);
  else if (P.x3) tc_trunk_kernel<true, false><<<grid, NTHREADS, smem_bytes(), s>>>(P);
  else if (f16) tc_trunk_kernel<false, true><<<grid, NTHREADS, smem_bytes(), s>>>(P);
  else tc_trunk_kernel<false, false><<<grid, NTHREADS, smem_bytes(), s>>>(P);
  DGDM_LAUNCH_CHECK();
  if (e1) DGDM_CUDA(cudaEventRecord(e1, s));
  timing_lock.unlock();
  if (acc_mode) {
    fixup_pairs_kernel<<<grid, 256, 0, s>>>(dUp, prefix, cont, (int)n_tiles, grid, G, H1, max_cont, n_rows, P.inv_gscale);
  } else if (backward) {
    reduce_slots_kernel<<<(unsigned)((n_pairs * H1 + 255) / 256), 256, 0, s>>>(dUp, part, n_pairs, G, H1, P.inv_gscale);
  } else {
    reduce_score_slots_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, s>>>(score_sum, score_part, n_pairs, G);
  }
  DGDM_LAUNCH_CHECK();
  return DGDM_OK;
}

}  // namespace dgdm

#ifdef DGDM_TRUNK_TRACE
extern "C" int dgdm_trunk_trace_read(long long* out, int32_t n) {
  return cudaMemcpyFromSymbol(out, dgdm::g_trace, sizeof(long long) * (size_t)(n < 8192 ? n : 8192)) == cudaSuccess ? 0 : -3;
}
#endif

extern "C" int dgdm_trunk_timing(int32_t enable) {
  using namespace dgdm;
  std::lock_guard<std::mutex> lk(g_timing.mu);
  g_timing.on = enable != 0;
  g_timing.used = 0;
  g_timing.rows = 0;
  g_timing.fwd_only.clear();
  g_timing.pair_rows.clear();
  if (!g_timing.on) {
    for (cudaEvent_t e : g_timing.ev) cudaEventDestroy(e);
    g_timing.ev.clear();
  }
  return DGDM_OK;
}

extern "C" int dgdm_trunk_timing_read(double* total_ms, int64_t* launches, int64_t* rows) {
  using namespace dgdm;
  DGDM_CHECK_ARG(total_ms && launches && rows, "dgdm_trunk_timing_read: null pointer");
  std::lock_guard<std::mutex> lk(g_timing.mu);
  double tot = 0.0;
  for (size_t i = 0; i + 1 < g_timing.used; i += 2) {
    DGDM_CUDA(cudaEventSynchronize(g_timing.ev[i + 1]));
    float ms = 0.f;
    DGDM_CUDA(cudaEventElapsedTime(&ms, g_timing.ev[i], g_timing.ev[i + 1]));
    tot += ms;
  }
  *total_ms = tot; *launches = (int64_t)(g_timing.used / 2); *rows = g_timing.rows;
  return DGDM_OK;
}

extern "C" int dgdm_trunk_timing_read_split(double* ms, int64_t* launches, int64_t* rows) {
  using namespace dgdm;
  DGDM_CHECK_ARG(ms && launches && rows, "dgdm_trunk_timing_read_split: null pointer");
  std::lock_guard<std::mutex> lk(g_timing.mu);
  for (int k = 0; k < 2; ++k) { ms[k] = 0.0; launches[k] = 0; rows[k] = 0; }
  for (size_t i = 0; i + 1 < g_timing.used; i += 2) {
    DGDM_CUDA(cudaEventSynchronize(g_timing.ev[i + 1]));
    float t = 0.f;
    DGDM_CUDA(cudaEventElapsedTime(&t, g_timing.ev[i], g_timing.ev[i + 1]));
    const int k = g_timing.fwd_only[i / 2];
    ms[k] += t; launches[k] += 1; rows[k] += g_timing.pair_rows[i / 2];
  }
  return DGDM_OK;
}

extern "C" size_t dgdm_dyn_tc_image_bytes(int32_t H1) {
  if (H1 != 256 && H1 != 512) return 0;
  return 3 * dgdm::make_plan(nullptr, H1).bytes;          // bf16 image + fp16 x3 image + fp16 single-pass image
}

extern "C" int dgdm_dyn_pack_tc(const dgdm_dyn_weights* w, void* tc_image, void* stream) {
  using namespace dgdm;
  DGDM_CHECK_ARG(w && tc_image, "dgdm_dyn_pack_tc: null pointer");
  DGDM_CHECK_ARG(w->H1 == 256 || w->H1 == 512, "dgdm_dyn_pack_tc: H1=%d unsupported", w->H1);
  DGDM_CHECK_ARG(((uintptr_t)tc_image) % 128 == 0, "dgdm_dyn_pack_tc: image must be 128-byte aligned");
  Plan pl = make_plan(w, w->H1);
  for (int im = 0; im < 3; ++im) {             // 0: bf16, 1: fp16 x F16_SW (hi + lo), 2: fp16 unscaled (single pass)
    for (int i = 0; i < pl.n_seg; ++i) {
      const int chunks = 8 * pl.pack[i].n_rows * 8;
      pack_tc_kernel<<<(chunks + 255) / 256, 256, 0, (cudaStream_t)stream>>>((uint8_t*)tc_image + (size_t)im * pl.bytes,
                                                                            pl.pack[i], im > 0, im == 1 ? F16_SW : 1.f);
      DGDM_LAUNCH_CHECK();
    }
  }
  return DGDM_OK;
}
