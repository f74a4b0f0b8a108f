// Fused pair-scoring + top-K kernels for sm_100a (tcgen05 / TMEM / TMA engine).
//
// Replace, for a block of users against a shard of the catalogue, the loop of reference
// src/inference/recommender.py:97-106 around MultimodalRecommender.forward
// (src/models/multimodal.py:528-610) for fusion_type 'gated' (src/models/layers.py:195-225),
// 'concatenate' (multimodal.py:583-584) and 'attention' (layers.py:135-164) with the default prediction MLP
// [512, 256, 128] -> 1 (multimodal.py:366-386).  One kernel template, three front ends.
//
// Design (DESIGN.md §5 has the derivation and the measurements):
//   * One persistent CTA PAIR (cluster of 2, tcgen05 cta_group::2) per two SMs.  A tile is 256
//     (user, item) pairs: 128 rows per CTA = 8 users x 16 items.  Each CTA keeps HALF of every weight
//     matrix (split along N) resident in shared memory for the whole kernel, loaded once by the TMA
//     engine from pre-swizzled images: the pair together holds all 384 KB of 16-bit weights, which no
//     single SM could.
//   * Layer chain per tile, every accumulator and every hidden activation in tensor memory (512
//     columns, all used):
//       gated   A1 = fused vector (CUDA cores: gate softmax + weighted sum -> swizzled smem tile)
//               D1[c] = A1 . W1[c]^T, 8 N-chunks of 64        (tcgen05.mma SS, M=256, N=64, K=64)
//               H1[c] = 16bit(relu(D1[c] + b1)) in place      (tcgen05.ld / cvt.relu / tcgen05.st)
//       concat  H1[c] = 16bit(relu(Pu[user] + Pi[item]))      (CUDA cores straight into TMEM: layer 1 is
//               split into per-user / per-item partials; Pi tiles staged by TMA bulk copies)
//       attention  as gated, with A1 = sum over the 6 tokens of the LayerNormed attention rows (CUDA cores; the
//               item-item part of the attention lives in per-item records, see "attention fusion front end" below)
//       both    D2 += H1[c] . W2[:, c]^T                      (tcgen05.mma TS: A from TMEM, N=256)
//               H2  = 16bit(relu(D2 + b2)) in place;  D3 = H2 . W3^T (TS, N=128, K=256)
//               z   = w4 . relu(D3 + b3) + b4 ; score = final(z)   (layer-3 epilogue, CUDA cores)
//     Eval-mode BatchNorm is folded into the next Linear on load (pxr_load_weights).
//   * 4 H1 chunk buffers; layer-1 work for a chunk is issued 2-4 chunks ahead of the layer-2 MMA that
//     consumes it, so epilogue latency stays off the tensor pipe's critical path.
//   * The score never leaves the SM: rows that beat the user's running K-th best go through a small
//     shared-memory queue to a dedicated warp that keeps one sorted 64-slot list per user (ties ->
//     lower item index = the reference's stable sort) and writes K (score, index) per user at the end.
//   * Seen items (filter_seen, recommender.py:88-90): the user's ascending history is walked with a
//     cursor in step with the ascending item sweep -> a 16-bit mask per (user, tile); no per-pair search.
//   * Warp roles (16 warps): 0-3 front end (A1 tiles / Pi staging + Pu), 4 issues every MMA (one thread
//     of the leader CTA) and owns TMEM/TMA setup, 5 top-K (plus 6 for short units, template switch TK2),
//     8-11 / 12-15 two epilogue groups; attention adds a second front-end warpgroup (warps 16-19) and
//     moves registers between the roles with setmaxnreg.
//     All hand-offs are mbarriers; tcgen05.commit multicasts completion to both CTAs.
#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include <cuda_fp16.h>

#include "pxr_common.cuh"
#include "tc_ptx.cuh"

// This source is compiled once per fusion_activation (build.py: -DPXR_TC_TU=<pxr_act>), one object each, so that the kernel
// instantiations of the activations build in parallel.  TU 0 (ReLU, bf16 operands) also holds the one-off preparation kernels
// and the host side; TU a = 1..4 only exports pxr_tc_launch_act<a>(); TU 5 holds the ReLU kernels for fp16 operands.
#ifndef PXR_TC_TU
#define PXR_TC_TU 0
#endif

namespace tc {

constexpr int D = 64, H1 = 512, H2 = 256, H3 = 128;
constexpr int KCAP = 64;                  // slots of the per-user sorted list (K <= KCAP)
constexpr int QCAP = 512;                 // candidate queue entries
constexpr int THREADS = 512;                 // 16 warps
constexpr uint32_t IDX_MASK = 0x0FFFFFFFu;   // 28-bit item index inside a queue / list key
constexpr int FMT_BF16 = 0, FMT_FP16 = 1;
constexpr int F_CONCAT = 0, F_GATED = 1, F_ATTN = 2, F_GATEDW = 3;   // front ends of the kernel template
// F_GATEDW ("wide" gated fusion, any embedding_dim): layer 1 of a gated model is linear in the fused vector, so
//   W1 (g_0 E_u + sum_m g_m f_m) + b1 = g_0 (W1 E_u + b1) + sum_m g_m (W1 f_m + b1)        (the gate weights sum to 1)
// i.e. a gate-weighted sum of one per-user and M - 1 per-item PARTIALS of 512 columns, computed once per user / item
// (SURVEY.md A3's split applied to A4).  The kernel is the concat pipeline (layer-1 producers on CUDA cores, layers 2 / 3 on
// tcgen05) with the broadcast add replaced by that weighted sum; it does not depend on embedding_dim.
template <int FUS> __host__ __device__ constexpr bool l1_on_tensor_pipe() { return FUS == F_GATED || FUS == F_ATTN; }
constexpr int M_PLAIN = 0, M_PAGED = 1, M_SPREAD = 2;  // kernel modes (template parameter MODE)
// users x items of one CTA tile (128 rows).  gated / concat: 8 x 16, row = user * 16 + item.  attention: 16 x 8,
// row = item * 16 + user -- the 16 rows of an item are the M dimension of the front end's register MMAs
template <int FUS> __host__ __device__ constexpr int tile_users() { return FUS == F_ATTN ? 16 : 8; }
template <int FUS> __host__ __device__ constexpr int tile_items() { return FUS == F_ATTN ? 8 : 16; }
constexpr int THREADS_ATT = 640;             // attention: a second front-end warpgroup (warps 16-19)
template <int FUS> __host__ __device__ constexpr int n_threads() { return FUS == F_ATTN ? THREADS_ATT : THREADS; }
constexpr int NH = 4, DH = D / NH;                      // attention: heads x head dim (fast path: 4 x 16)
// attention: per-item record = 16-bit MMA B fragments in lane order (see item_attn_kernel), in 16-byte units
constexpr int ATT_TOKENS = 5;
constexpr int REC_S = 0;                                // [4 heads][32 lanes]      q / k score fragments
constexpr int REC_L = REC_S + NH * 32;                  // [4 quad lanes][2]        logsumexp of the item-item scores (fp32)
constexpr int REC_T = REC_L + 8;                        // [5 tokens][4][32 lanes]  tail K rows: Nc hi, xc hi / lo, Nc lo
constexpr int REC_U = REC_T + ATT_TOKENS * 128;         // [2 head pairs][4][32]    per-head out-projected item values
constexpr int ATT_REC_U4 = REC_U + 256;                 // 1032 x 16 B = 16 512 B per item
#ifndef PXR_TOPK_IDLE_NS
#define PXR_TOPK_IDLE_NS 200
#endif

enum {
  BAR_W = 0, BAR_A_FULL, BAR_A_EMPTY, BAR_D1_FULL0, BAR_D1_FULL1, BAR_D1_FULL2, BAR_D1_FULL3, BAR_H1_FULL0, BAR_H1_FULL1,
  BAR_H1_FULL2, BAR_H1_FULL3, BAR_H1_EMPTY0, BAR_H1_EMPTY1, BAR_H1_EMPTY2, BAR_H1_EMPTY3, BAR_PI_FULL0, BAR_PI_FULL1,
  BAR_PI_EMPTY0, BAR_PI_EMPTY1, BAR_D2_FULL, BAR_H2_FULL0, BAR_H2_FULL1, BAR_D3_FULL, BAR_D3_EMPTY, BAR_UNIT_DONE,
  BAR_UNIT_RESET, BAR_Q_FULL0, BAR_Q_FULL1, BAR_Q_FULL2, BAR_Q_EMPTY0, BAR_Q_EMPTY1, BAR_Q_EMPTY2, N_BARS
};

// Per-CTA scratch in shared memory.  (The attention variant keeps the epilogue biases in the kernel-parameter constant
// bank and stages its per-unit user constants in the idle A1 tile; its per-user operands then live in registers.)
template <int FUS>
struct MiscT {
  static constexpr bool ATT = (FUS == F_ATTN);
  static constexpr int TU = tile_users<FUS>();
  static constexpr int QC = ATT ? 256 : QCAP;          // candidate queue entries
  float b1[ATT ? 4 : H1]; float b2[ATT ? 4 : H2]; float b3[ATT ? 4 : H3]; float w4[ATT ? 4 : H3];   // contiguous
  float eu[ATT ? 1 : TU][D];
  float lu[ATT ? 1 : TU][8];
  unsigned long long list[TU][KCAP];
  unsigned long long upper[TU];                      // per user slot: only keys strictly below it are admitted (pages of a top_k > 64 call)
  int32_t slot_lo[TU], slot_hi[TU];                  // M_SPREAD: first row / end row of the slot's item sub-range (front end only)
  unsigned long long queue[QC];
  float thr[TU];
  uint32_t seen_mask[4][TU];
  uint32_t q_tail[2], q_head[2];                     // two candidate queues: lower half of the users -> top-K warp 5, upper half -> warp 6
  uint32_t tmem_base;
  float b4;
  unsigned long long bars[N_BARS];
};

// shared / tensor memory maps per front end
template <int FUS>
struct Map {
  static constexpr bool GATED = l1_on_tensor_pipe<FUS>();        // layer 1 on the tensor pipe from an A1 tile
  // shared memory (bytes from a 1024-aligned base); the weight image is the first WIMG bytes
  static constexpr uint32_t OFF_W1 = 0;                          // gated: 8 N-chunks x (32 rows x 128 B) = 32 KB
  static constexpr uint32_t OFF_W2 = GATED ? 32768u : 0u;        // 8 K-blocks x (128 rows x 128 B) = 128 KB
  static constexpr uint32_t OFF_W3 = OFF_W2 + 131072u;           // 4 K-blocks x (64 rows x 128 B)  = 32 KB
  static constexpr uint32_t WIMG = OFF_W3 + 32768u;
  static constexpr uint32_t OFF_A1 = WIMG;                       // gated: 128 rows x 128 B, SWIZZLE_128B
  static constexpr uint32_t PI_STRIDE = 1040, PI_BUF = tile_items<F_CONCAT>() * PI_STRIDE;   // concat: 16 item partials (512 x 16 bit)
  static constexpr uint32_t OFF_PI = WIMG;                       // padded by 16 B per row: conflict-free reads
  // wide gated: the item partials of a tile are staged one 64-column chunk at a time: per item M - 1 rows of 128 B + 16 B of
  // padding (conflict-free 16-byte reads), 16 items per stage, a ring of three stages in the Pi area
  static constexpr uint32_t Q_STAGES = 3, Q_STAGE_MAX = 16u * (5u * 128u + 16u);
  static constexpr uint32_t PU_STRIDE = 2064;                    // concat: 8 user partials (512 fp32), padded
  static constexpr uint32_t OFF_PU = OFF_PI + 2 * PI_BUF;
  static_assert(Q_STAGES * Q_STAGE_MAX <= 2 * PI_BUF, "the chunk ring of the wide gated front end lives in the Pi area");
  static constexpr uint32_t OFF_MISC = GATED ? OFF_A1 + 16384u : OFF_PU + tile_users<F_CONCAT>() * PU_STRIDE;
  static constexpr uint32_t SMEM = OFF_MISC + (uint32_t)sizeof(MiscT<FUS>) + 1024u;   // + alignment slack
  // tensor memory (columns)
  static constexpr uint32_t TM_D3 = 128, TM_D2 = 256;
  // H1 chunk buffer b.  gated: 64-col fp32 accumulators packed in place, buffers 2,3 share the D3 columns
  // (D3 is only live at the tile boundary).  concat: four 32-col packed buffers.
  __host__ __device__ static constexpr uint32_t h1buf(int b) {
    return GATED ? (b < 2 ? 64u * b : TM_D3 + 64u * (b - 2)) : 32u * b;
  }
};
static_assert(Map<F_GATED>::SMEM <= 232448 && Map<F_CONCAT>::SMEM <= 232448 && Map<F_ATTN>::SMEM <= 232448 && Map<F_GATEDW>::SMEM <= 232448, "shared memory budget");
static_assert(Map<F_CONCAT>::OFF_MISC % 16 == 0 && Map<F_GATED>::OFF_MISC % 16 == 0 && Map<F_GATEDW>::OFF_MISC % 16 == 0, "alignment");
static_assert(offsetof(MiscT<F_ATTN>, list) % 8 == 0, "alignment");
static_assert(offsetof(MiscT<F_GATED>, list) % 8 == 0 && offsetof(MiscT<F_GATED>, bars) % 8 == 0 && offsetof(MiscT<F_ATTN>, bars) % 8 == 0, "alignment");

struct Params {
  const uint8_t* wimg;          // [2][WIMG] pre-swizzled 16-bit operand images (rank 0, rank 1)
  const float* bias;            // b1[512] b2[256] b3[128] w4[128] b4   (b1 unused by concat: folded into Pi)
  const float* gate_w;          // gated: (M, M*D) fp32 row-major; the user part is the first D of each row
  const float* item_feats;      // gated: [rows][M-1][D] fp32 projected item-side modality vectors
  const float* item_logit;      // gated: [rows][8] fp32 item part of the gate logits (+ gate bias)
  const uint16_t* item_q;       // wide gated: 16-bit per-modality item partials of layer 1, Q_m = W1 f_m + b1, chunk-major:
                                // [tile of 16 rows][8 chunks][16 items][(M-1) x 64 columns + 8 of padding]  (q_stage_bytes per chunk)
  const uint16_t* item_pi;      // concat: [rows][512] 16-bit item partial of layer 1 (+ b1)
  const float* w1u_t;           // concat: [64][512] fp32, user columns of W1 transposed
  const uint4* attn_rec;        // attention: [rows][ATT_REC_U4] per-item records (16-bit MMA B fragments in lane order)
  const uint4* wo_frag;         // attention: centred out_proj weight as B fragments, [4 k-steps][4][32 lanes]
  uint4* xc0_scratch;           // attention: per CTA, the per-unit user operands of its 16 users in lane order (ATT_XC0_U4 x 16 B, SCR_*)
  const float* attn_in_wt;      // attention: in_proj weight transposed [64][192], bias [192]
  const float* attn_in_b;
  const float* attn_out_wt;     // attention: out_proj weight transposed [64][64], bias [64]
  const float* attn_out_b;
  const float* ln_w;            // attention: LayerNorm weight / bias [64]
  const float* ln_b;
  float bias_c[H1 + H2 + H3 + H3 + 4];   // attention: b1' b2 b3 w4 b4 read through the constant bank (no room in smem)
  const float* user_emb;        // (n_users_total, D) fp32 table
  const int64_t* user_idx;      // (n_users,)
  const int64_t* seen_indptr;   // (n_users + 1,) or NULL
  const int32_t* seen_idx;      // global item indices, ascending per user
  const uint8_t* item_missing;  // per item row: 1 = features missing, the score is 0.0 (recommender.py:229-230); NULL = none
  const unsigned long long* upper;   // [n_users] or NULL.  Page p > 0 of a top_k > 64 call: the key (ordered score << 32 | IDX_MASK - index)
                                // of the user's last entry of page p - 1; only strictly smaller keys are admitted (0: nothing is left)
  float* out_scores;            // [S][n_users][K]
  int32_t* out_idx;
  int64_t n_users, n_rows, item_base;
  int M, K, S, rows_per_split, n_units, final_act;
  int Dm;                       // embedding dim of the model (concat / wide gated: any multiple of 4 up to 512; gated / attention: 64)
  // M_SPREAD (small batches): n_users above counts VIRTUAL users v = r * n_real + u: real user u on item sub-range
  // r = rows [r * sub_rows, (r + 1) * sub_rows) of the shard; user_idx / seen_indptr are indexed by u = v % n_real
  int64_t n_real;
  int sub_rows;                 // multiple of the tile's item count
};

struct Unit { int g, s; int64_t row_lo, row_hi; int ntiles; };

template <int TI, bool SPREAD = false>
__device__ __forceinline__ Unit decode_unit(const Params& p, int w) {
  Unit u;
  u.g = w / p.S; u.s = w % p.S;
  u.row_lo = (int64_t)u.s * p.rows_per_split;
  u.row_hi = min(p.n_rows, u.row_lo + (int64_t)p.rows_per_split);
  u.ntiles = SPREAD ? p.sub_rows / TI : (int)((u.row_hi - u.row_lo + TI - 1) / TI);      // SPREAD: every slot sweeps one sub-range
  return u;
}

// ---------------------------------------------------------------------------------------------
// 16-bit operand format helpers
// ---------------------------------------------------------------------------------------------
// fusion_activation (multimodal.py:150-167) of the hidden layers inside the fused chain.  ReLU is one cvt.relu per pair of
// values; the others are evaluated on the fp32 accumulator before the 16-bit rounding: leaky_relu = max(x, 0.01 x),
// silu = x / (1 + e^-x) and tanh = 1 - 2 / (1 + e^2x) with MUFU exp / reciprocal (relative error ~1e-6, far below half a
// 16-bit ulp), gelu with erff as in the fp32 kernels.  act(0) = 0 for all five, so zero-padded hidden units stay exact.
template <int ACT> __device__ __forceinline__ float act_fast(float x) {
  if (ACT == PXR_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == PXR_ACT_LEAKY_RELU) return fmaxf(x, 0.01f * x);
  if (ACT == PXR_ACT_SILU) return __fdividef(x, 1.f + __expf(-x));
  if (ACT == PXR_ACT_TANH) return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x));
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));      // PXR_ACT_GELU (erf form, nn.GELU default)
}
template <int ACT, int FMT> __device__ __forceinline__ uint32_t act_pack(float lo, float hi) {
  uint32_t d;
  if (ACT == PXR_ACT_RELU) {
    if (FMT == FMT_BF16) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));   // saturate: no inf from fp16 range
  } else {
    lo = act_fast<ACT>(lo); hi = act_fast<ACT>(hi);
    if (FMT == FMT_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  }
  return d;
}
template <int FMT> __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (FMT == FMT_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <int FMT> __device__ __forceinline__ float2 unpack2(uint32_t x) {
  if (FMT == FMT_BF16) return make_float2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
  return __half22float2(*reinterpret_cast<const __half2*>(&x));
}
template <int FMT> __host__ __device__ constexpr uint32_t idesc(int M, int N) {
  // cute::UMMA::InstrDescriptor: D fp32 (bits 4-5 = 1), A/B format (bits 7-9 / 10-12: 0 = F16, 1 = BF16), K-major both
  return (1u << 4) | ((FMT == FMT_BF16 ? 1u : 0u) << 7) | ((FMT == FMT_BF16 ? 1u : 0u) << 10) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// epilogue pieces (one warp = 32 TMEM lanes = 32 rows; taddr already carries the lane base)
// ---------------------------------------------------------------------------------------------
template <int ACT, int FMT>
__device__ __forceinline__ void bias_act_pack32(const uint32_t* v, const float* bias, uint32_t* o) {
  const float4* bb = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = bb[q];
    o[2 * q] = act_pack<ACT, FMT>(__uint_as_float(v[4 * q]) + b.x, __uint_as_float(v[4 * q + 1]) + b.y);
    o[2 * q + 1] = act_pack<ACT, FMT>(__uint_as_float(v[4 * q + 2]) + b.z, __uint_as_float(v[4 * q + 3]) + b.w);
  }
}

// 64 fp32 accumulator columns -> 32 packed 16-bit columns written over the start of the same region
template <int ACT, int FMT>
__device__ __forceinline__ void epi_pack64(uint32_t t_src, uint32_t t_dst, const float* bias) {
  uint32_t v0[32], v1[32], o[32];
  ptx::tmem_ld32(t_src, v0);
  ptx::tmem_ld32(t_src + 32, v1);
  ptx::tc_wait_ld();
  bias_act_pack32<ACT, FMT>(v0, bias, o);
  bias_act_pack32<ACT, FMT>(v1, bias + 32, o + 16);
  ptx::tmem_st32(t_dst, o);
}

// concat layer 1 for one row and one 64-wide chunk: act(Pu[user] + Pi[item]) -> 32 packed columns
template <int ACT, int FMT>
__device__ __forceinline__ void concat_h1_chunk(const uint8_t* pi_row, const uint8_t* pu_row, uint32_t t_dst) {
  uint32_t o[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 pv = *reinterpret_cast<const uint4*>(pi_row + 16 * q);
    const float4 a = *reinterpret_cast<const float4*>(pu_row + 32 * q);
    const float4 b = *reinterpret_cast<const float4*>(pu_row + 32 * q + 16);
    const float2 p0 = unpack2<FMT>(pv.x), p1 = unpack2<FMT>(pv.y), p2 = unpack2<FMT>(pv.z), p3 = unpack2<FMT>(pv.w);
    o[4 * q + 0] = act_pack<ACT, FMT>(a.x + p0.x, a.y + p0.y);
    o[4 * q + 1] = act_pack<ACT, FMT>(a.z + p1.x, a.w + p1.y);
    o[4 * q + 2] = act_pack<ACT, FMT>(b.x + p2.x, b.y + p2.y);
    o[4 * q + 3] = act_pack<ACT, FMT>(b.z + p3.x, b.w + p3.y);
  }
  ptx::tmem_st32(t_dst, o);
}

// wide gated layer 1 for one row and one 64-wide chunk: act(g_0 Pu[user] + sum_m g_{m+1} Q_m[item]) -> 32 packed columns.
// `qi` = this row's item block of the staged chunk in shared memory ([M-1] rows of 64 columns, 16 bit; blocks are 16 B
// apart modulo 128, so the 16-byte reads of a quarter warp hit distinct banks; the two user rows of a warp share every read).
__host__ __device__ constexpr uint32_t q_item_bytes(int nm) { return (uint32_t)nm * 128u + 16u; }
__host__ __device__ constexpr uint32_t q_stage_bytes(int nm) { return 16u * q_item_bytes(nm); }       // one chunk of one 16-item tile
// PXR_GW_HALF2 (default): the item partials are stored in fp16 (11-bit significand, saturating convert) whatever the MMA
// operand format, and their gate-weighted sum runs on packed fp16 FMAs -- 5 HFMA2 per two columns instead of 10 unpack + 10
// FFMA: the producers are issue-bound (1.90 -> 2.51 G pairs/s, DESIGN K3w).  The fp16 sum carries <= 5 roundings of 2^-11 relative, a
// fifth of the 16-bit rounding of the activation that follows; the user term g_0 Pu stays fp32.
#ifndef PXR_GW_HALF2
#define PXR_GW_HALF2 1
#endif
constexpr bool GW_HALF2 = PXR_GW_HALF2 != 0;
template <int ACT, int FMT>
__device__ __forceinline__ void gatedw_h1_chunk_h2(const uint8_t* qi, int nm, float g0, const __half2 (&gh)[5], const uint8_t* pu_row, uint32_t t_dst) {
  uint32_t o[32];
#pragma unroll
  for (int s = 0; s < 4; ++s) {                      // 16 columns per step
    uint4 pv[5][2];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      if (m < nm) {
        const uint4* src = reinterpret_cast<const uint4*>(qi + m * 128 + 32 * s);
        pv[m][0] = src[0]; pv[m][1] = src[1];
      } else {
        pv[m][0] = pv[m][1] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const float4 pa = *reinterpret_cast<const float4*>(pu_row + 64 * s + 32 * hh);
      const float4 pb = *reinterpret_cast<const float4*>(pu_row + 64 * s + 32 * hh + 16);
      const float pu[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
      uint32_t w[5][4];
#pragma unroll
      for (int m = 0; m < 5; ++m) { w[m][0] = pv[m][hh].x; w[m][1] = pv[m][hh].y; w[m][2] = pv[m][hh].z; w[m][3] = pv[m][hh].w; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 a = __hmul2(gh[0], *reinterpret_cast<const __half2*>(&w[0][j]));
#pragma unroll
        for (int m = 1; m < 5; ++m) a = __hfma2(gh[m], *reinterpret_cast<const __half2*>(&w[m][j]), a);
        const float2 f = __half22float2(a);
        o[8 * s + 4 * hh + j] = act_pack<ACT, FMT>(fmaf(g0, pu[2 * j], f.x), fmaf(g0, pu[2 * j + 1], f.y));
      }
    }
  }
  ptx::tmem_st32(t_dst, o);
}
template <int ACT, int FMT>
__device__ __forceinline__ void gatedw_h1_chunk(const uint8_t* qi, int nm, const float (&g)[6], const uint8_t* pu_row, uint32_t t_dst) {
  uint32_t o[32];
#pragma unroll
  for (int s = 0; s < 4; ++s) {                      // 16 columns per step
    uint4 pv[5][2];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      if (m < nm) {
        const uint4* src = reinterpret_cast<const uint4*>(qi + m * 128 + 32 * s);
        pv[m][0] = src[0]; pv[m][1] = src[1];
      } else {
        pv[m][0] = pv[m][1] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    float acc[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(pu_row + 64 * s + 16 * q);
      acc[4 * q + 0] = g[0] * a.x; acc[4 * q + 1] = g[0] * a.y; acc[4 * q + 2] = g[0] * a.z; acc[4 * q + 3] = g[0] * a.w;
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const float gm = g[m + 1];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float2 p0 = unpack2<FMT>(pv[m][hh].x), p1 = unpack2<FMT>(pv[m][hh].y), p2 = unpack2<FMT>(pv[m][hh].z), p3 = unpack2<FMT>(pv[m][hh].w);
        float* ac = acc + 8 * hh;
        ac[0] = fmaf(gm, p0.x, ac[0]); ac[1] = fmaf(gm, p0.y, ac[1]); ac[2] = fmaf(gm, p1.x, ac[2]); ac[3] = fmaf(gm, p1.y, ac[3]);
        ac[4] = fmaf(gm, p2.x, ac[4]); ac[5] = fmaf(gm, p2.y, ac[5]); ac[6] = fmaf(gm, p3.x, ac[6]); ac[7] = fmaf(gm, p3.y, ac[7]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[8 * s + i] = act_pack<ACT, FMT>(acc[2 * i], acc[2 * i + 1]);
  }
  ptx::tmem_st32(t_dst, o);
}


// ---------------------------------------------------------------------------------------------
// attention fusion front end (src/models/layers.py:135-164 as documented; SURVEY.md A5 split) on register-level MMAs
//
// Tokens of pair (u, i): x_0 = E_u, x_1.. = the item-side modality vectors; o_a = concatenated per-head attention
// outputs of row a; y_a = x_a + b_o + W_o o_a; fused = mean_a LayerNorm(y_a).  LayerNorm only needs y - mean(y), so
// with P = I - 11^T / D everything is carried centred over d: Wc = P W_o, xc_a = P (x_a + b_o), yc_a = xc_a + Wc o_a.
//   item row a >= 1: the user column gets w_ah = sigmoid(s_a0,h - L_ah), L_ah = logsumexp_b>=1 s_ab,h (per item), and
//       yc_a = xc_a + sum_h (1 - w_ah) Nc_ah + Wc (w_a (.) v_u),      Nc_ah = Wc[:, head h] vbar_ah   (per item)
//   user row:    softmax over [q_u.k_u, q_u.k_b ...] per head,
//       yc_0 = xc_0 + Wc (p_00 (.) v_u) + sum_h sum_b p_0b,h Uc_bh,   Uc_bh = Wc[:, head h] v_b,h     (per item)
//   fused = ln_b + ln_w / M * sum_a yc_a rstd_a; the affine part is folded into layer 1 (W1' = W1 diag(ln_w / M),
//   b1' = b1 + W1 ln_b), so the A1 operand is the 16-bit rounding of acc = sum_a yc_a rstd_a.
// Every "coefficients x vectors" product above is a small GEMM whose B operand depends on the ITEM only (or on nothing:
// Wc), so for one item and 16 users it is a chain of mma.sync.m16n8k16 (M = the 16 users of the CTA, fp32 accumulators in
// registers -- the tcgen05 chain owns all 512 TMEM columns, and these GEMMs are 1/6 of its FLOPs): scores (K = 16 per
// head), per token [w_a (.) v_u | 1 - w_a, 1] (K = 80) x [Wc ; Nc_a hi, xc_a hi, xc_a lo, Nc_a lo] -> yc_a (16 x 64),
// the user row with K = 64 + 32.  The per-item B fragments are stored in the record in lane order (one coalesced
// 16-byte load per lane and fragment pair), Wc's fragments stay in L1.  CUDA cores only produce coefficients (20 sigmoids, a
// 6-way softmax per head), the A fragments (packed 16-bit multiplies by the per-user value fragments kept in registers),
// sum y^2, and the rstd-weighted token sum: ~95 warp instructions per pair instead of ~360 for the all-CUDA-core form.
// A front-end warp owns one item x 16 users at a time (tile = 8 items x 16 users, two items per warp).
// ---------------------------------------------------------------------------------------------
// experiment switches of the attention front end (scripts/ab_build.sh -DPXR_ATT_...=1)
#ifndef PXR_ATT_REC_L1
#define PXR_ATT_REC_L1 0        // 1: record fragments allocate in L1 (plain ld.global.nc) instead of streaming past it
#endif
#ifndef PXR_ATT_WCBATCH
#define PXR_ATT_WCBATCH 0       // 1: the four Wc fragment loads of a k-step are issued before its eight MMAs
#endif
#ifndef PXR_ATT_NOPF
#define PXR_ATT_NOPF 0          // 1: no L1 / L2 prefetches
#endif
#ifndef PXR_ATT_PFS
#define PXR_ATT_PFS 0           // 1: the score fragments of the next tile's item are prefetched into L1 at the end of a step
#endif
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {      // item records: read once per (item, 16 users), keep them out of L1
#if PXR_ATT_REC_L1
  return __ldg(p);
#else
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#endif
}
template <int FMT>
__device__ __forceinline__ void hmma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FMT == FMT_BF16)
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int FMT> __device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b) {     // packed 16-bit multiply, rounded once
  uint32_t d;
  if (FMT == FMT_BF16) asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t dup_lo(uint32_t x) { return __byte_perm(x, x, 0x1010); }
__device__ __forceinline__ uint32_t dup_hi(uint32_t x) { return __byte_perm(x, x, 0x3232); }

// per-unit constants of a front-end thread: A fragments (rows = the CTA's 16 users) of k_u, q_u / sqrt(dh), v_u per head,
// xc_0 = P (E_u + b_o) in accumulator layout, q_u.k_u / sqrt(dh) of rows g and g + 8
struct AttnUserFrag {            // what a front-end thread keeps in registers for the whole unit
  uint32_t vu[NH][4];
  float s00[2][NH];
};
// Per-CTA global scratch (L2) with the per-unit user operands in lane order, written by warp 0 of the front end during the
// unit setup and read by all eight front-end warps: the k_u / q_u fragments and xc_0 are only needed once per item step,
// so they are fetched per step instead of occupying 64 of the 128 registers a front-end thread has.  16-byte units:
constexpr int SCR_XC0 = 0;                  // [8 n-tiles][32 lanes]  xc_0 in accumulator layout
constexpr int SCR_KU = 8 * 32;              // [4 heads][32 lanes]    A fragments of k_u
constexpr int SCR_QU = SCR_KU + 4 * 32;     // [4 heads][32 lanes]    A fragments of q_u / sqrt(dh)
constexpr int SCR_VU = SCR_QU + 4 * 32;     // [4 heads][32 lanes]    A fragments of v_u
constexpr int SCR_S00 = SCR_VU + 4 * 32;    // [2][32 lanes]          q_u.k_u / sqrt(dh) of rows g / g + 8 x 4 heads
constexpr int ATT_XC0_U4 = SCR_S00 + 2 * 32;

// `stage` = the idle A1 tile (16 KB): [16][192] fp32 in_proj outputs, then [16][64] fp32 E_u.  128 threads, named barrier 1.
template <int FMT>
__device__ __forceinline__ void attn_user_setup(const Params& p, float* stage, int tid, int64_t ubase, uint4* scr) {
  constexpr int TUA = tile_users<F_ATTN>();
  float (*qkv)[3 * D] = reinterpret_cast<float (*)[3 * D]>(stage);
  float (*eu)[D] = reinterpret_cast<float (*)[D]>(stage + TUA * 3 * D);
  for (int i = tid; i < TUA * (D / 4); i += 128) {
    const int u = i >> 4, d4 = (i & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ubase + u < p.n_users) v = *reinterpret_cast<const float4*>(p.user_emb + p.user_idx[ubase + u] * D + d4);
    *reinterpret_cast<float4*>(&eu[u][d4]) = v;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  for (int n = tid; n < 3 * D; n += 128) {                 // in_proj of the user token: q | k | v
    float acc[TUA];
    const float b = p.attn_in_b[n];
#pragma unroll
    for (int u = 0; u < TUA; ++u) acc[u] = b;
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
      const float w = p.attn_in_wt[k * 3 * D + n];
#pragma unroll
      for (int u = 0; u < TUA; ++u) acc[u] = fmaf(w, eu[u][k], acc[u]);
    }
    const float sc = n < D ? 0.25f : 1.f;                  // q / sqrt(dh), dh = 16
#pragma unroll
    for (int u = 0; u < TUA; ++u) qkv[u][n] = acc[u] * sc;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (tid < 32) {                                             // identical for every front-end warp: warp 0 writes the scratch
    const int lane = tid, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int c = DH * h + 2 * t;
#pragma unroll
      for (int m = 0; m < 3; ++m) {                          // 0: q, 1: k, 2: v
        const int o = m * D + c;
        uint4 f;
        f.x = pack2<FMT>(qkv[g][o], qkv[g][o + 1]);         f.y = pack2<FMT>(qkv[g + 8][o], qkv[g + 8][o + 1]);
        f.z = pack2<FMT>(qkv[g][o + 8], qkv[g][o + 9]);     f.w = pack2<FMT>(qkv[g + 8][o + 8], qkv[g + 8][o + 9]);
        scr[(m == 0 ? SCR_QU : (m == 1 ? SCR_KU : SCR_VU)) + h * 32 + lane] = f;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float dots[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < DH; ++e) dot = fmaf(qkv[g + 8 * r][DH * h + e], qkv[g + 8 * r][D + DH * h + e], dot);
        dots[h] = dot;
      }
      scr[SCR_S00 + r * 32 + lane] = make_uint4(__float_as_uint(dots[0]), __float_as_uint(dots[1]), __float_as_uint(dots[2]), __float_as_uint(dots[3]));
      // xc_0 = (E_u + b_o) centred over d
      const float* e = eu[g + 8 * r];
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) part += e[16 * t + i] + p.attn_out_b[16 * t + i];
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      const float mean = part * (1.f / D);
#pragma unroll
      for (int nn = 0; nn < 8; ++nn) {
        const int c = 8 * nn + 2 * t;
        reinterpret_cast<float2*>(scr + SCR_XC0 + nn * 32 + lane)[r] = make_float2(e[c] + p.attn_out_b[c] - mean, e[c + 1] + p.attn_out_b[c + 1] - mean);
      }
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");              // the stage (A1 tile) may be overwritten now
}

// Y += [x_h (.) v_u]_h . Wc^T: x[h] = (coefficient of row g, coefficient of row g + 8) packed; one k-step per head
template <int FMT>
__device__ __forceinline__ void attn_wo_pass(float (&Y)[8][4], const uint32_t (&x)[NH], const AttnUserFrag& U,
                                             const uint4* __restrict__ wo, int lane) {
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const uint32_t lo = dup_lo(x[h]), hi = dup_hi(x[h]);
    uint32_t a[4];
    a[0] = mul2<FMT>(lo, U.vu[h][0]); a[1] = mul2<FMT>(hi, U.vu[h][1]);
    a[2] = mul2<FMT>(lo, U.vu[h][2]); a[3] = mul2<FMT>(hi, U.vu[h][3]);
#if PXR_ATT_WCBATCH
    uint4 w[4];
#pragma unroll
    for (int np = 0; np < 4; ++np) w[np] = __ldg(wo + (h * 4 + np) * 32 + lane);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      hmma<FMT>(Y[2 * np], a, w[np].x, w[np].y);
      hmma<FMT>(Y[2 * np + 1], a, w[np].z, w[np].w);
    }
#else
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const uint4 w = __ldg(wo + (h * 4 + np) * 32 + lane);
      hmma<FMT>(Y[2 * np], a, w.x, w.y);
      hmma<FMT>(Y[2 * np + 1], a, w.z, w.w);
    }
#endif
  }
}

// acc (+)= Y * rstd(Y) for the rows g (entries 0, 1) and g + 8 (entries 2, 3); Y is centred by construction
template <bool FIRST>
__device__ __forceinline__ void attn_norm_acc(const float (&Y)[8][4], float (&acc)[8][4]) {
  float p0[4] = {0.f, 0.f, 0.f, 0.f}, p1[4] = {0.f, 0.f, 0.f, 0.f};      // independent partial sums: short dependency chains
#pragma unroll
  for (int nn = 0; nn < 8; ++nn) {
    p0[nn & 3] = fmaf(Y[nn][0], Y[nn][0], fmaf(Y[nn][1], Y[nn][1], p0[nn & 3]));
    p1[nn & 3] = fmaf(Y[nn][2], Y[nn][2], fmaf(Y[nn][3], Y[nn][3], p1[nn & 3]));
  }
  float s0 = (p0[0] + p0[1]) + (p0[2] + p0[3]), s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float r0 = rsqrtf(fmaf(s0, 1.f / D, 1e-5f)), r1 = rsqrtf(fmaf(s1, 1.f / D, 1e-5f));
#pragma unroll
  for (int nn = 0; nn < 8; ++nn) {
    if (FIRST) { acc[nn][0] = r0 * Y[nn][0]; acc[nn][1] = r0 * Y[nn][1]; acc[nn][2] = r1 * Y[nn][2]; acc[nn][3] = r1 * Y[nn][3]; }
    else {
      acc[nn][0] = fmaf(r0, Y[nn][0], acc[nn][0]); acc[nn][1] = fmaf(r0, Y[nn][1], acc[nn][1]);
      acc[nn][2] = fmaf(r1, Y[nn][2], acc[nn][2]); acc[nn][3] = fmaf(r1, Y[nn][3], acc[nn][3]);
    }
  }
}

// Y (+)= A . B for the 8 n-tiles, B fragments streamed from the record two n-tiles (16 bytes per lane) at a time
template <int FMT>
__device__ __forceinline__ void attn_rec_mma(float (&Y)[8][4], const uint32_t (&a)[4], const uint4* __restrict__ frag, int lane) {
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    const uint4 f = ldg_stream_u4(frag + np * 32 + lane);
    hmma<FMT>(Y[2 * np], a, f.x, f.y);
    hmma<FMT>(Y[2 * np + 1], a, f.z, f.w);
  }
}

// acc[nn][..] (accumulator layout: rows g / g + 8 of the 16 users, columns 8 nn + 2 t, + 1) = sum over the 1 + nt tokens of
// the normalised rows of pair (user row, this item).  Eight front-end warps (two per scheduler) cover each other's
// latencies, so the step is written for few registers (128 per thread): one accumulator set for the row being built, one for
// the token sum, record fragments consumed as they arrive; what the next token / the next item needs is pulled into L1 / L2
// by prefetches issued a token / a tile ahead.
template <int FMT>
__device__ __forceinline__ void attn_item_step(const AttnUserFrag& U, const uint4* __restrict__ rec, const uint4* __restrict__ rec_next,
                                               const uint4* __restrict__ wo, const uint4* __restrict__ scr, int nt, int lane,
                                               float (&acc)[8][4]) {
  const int t = lane & 3;
  if (rec_next && !PXR_ATT_NOPF) {                                                       // next tile's record -> L2 (129 lines of 128 B)
    const char* pn = reinterpret_cast<const char*>(rec_next) + lane * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + i * 4096));
    if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + 16384));
  }
  auto prefetch_tail = [&](int a) {                                                      // token a's tail fragments (2 KB) -> L1
    if (lane < 16 && !PXR_ATT_NOPF) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(rec + REC_T + a * 128) + lane * 128));
  };
  uint32_t wpk[2][NH];                                     // sigmoid weights of tokens 2t / 2t + 1: (row g, row g + 8) packed
  float Y[8][4];
  {
    // ---- scores: c1 = k_u,h . q_a,h (item rows, their user column), c0 = q_u,h . k_b,h (user row); entry 0 / 1: row g,
    //      token 2t / 2t + 1; entry 2 / 3: row g + 8
    uint32_t x0[NH], pa[2][4];
    {
      const uint4 La4 = __ldg(rec + REC_L + 2 * t), Lb4 = __ldg(rec + REC_L + 2 * t + 1);   // L of tokens 2t, 2t + 1 x 4 heads
      const float La[NH] = {__uint_as_float(La4.x), __uint_as_float(La4.y), __uint_as_float(La4.z), __uint_as_float(La4.w)};
      const float Lb[NH] = {__uint_as_float(Lb4.x), __uint_as_float(Lb4.y), __uint_as_float(Lb4.z), __uint_as_float(Lb4.w)};
      const bool v0 = 2 * t < nt, v1 = 2 * t + 1 < nt;
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const uint4 sf = ldg_stream_u4(rec + REC_S + h * 32 + lane);
        const uint4 kf = __ldcg(scr + SCR_KU + h * 32 + lane), qf = __ldcg(scr + SCR_QU + h * 32 + lane);
        const uint32_t ku[4] = {kf.x, kf.y, kf.z, kf.w}, qu[4] = {qf.x, qf.y, qf.z, qf.w};
        float c1[4] = {0.f, 0.f, 0.f, 0.f}, c0[4] = {0.f, 0.f, 0.f, 0.f};
        hmma<FMT>(c1, ku, sf.x, sf.y);
        hmma<FMT>(c0, qu, sf.z, sf.w);
        wpk[0][h] = pack2<FMT>(__fdividef(1.f, 1.f + __expf(La[h] - c1[0])), __fdividef(1.f, 1.f + __expf(La[h] - c1[2])));
        wpk[1][h] = pack2<FMT>(__fdividef(1.f, 1.f + __expf(Lb[h] - c1[1])), __fdividef(1.f, 1.f + __expf(Lb[h] - c1[3])));
        // user row: softmax over [q_u.k_u, q_u.k_b ...] of head h, rows g and g + 8 (the tokens are spread over the quad)
        const float a0 = v0 ? c0[0] : -INFINITY, a1 = v1 ? c0[1] : -INFINITY, b0 = v0 ? c0[2] : -INFINITY, b1 = v1 ? c0[3] : -INFINITY;
        float mg = fmaxf(fmaxf(a0, a1), U.s00[0][h]), mh = fmaxf(fmaxf(b0, b1), U.s00[1][h]);
        mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 1)); mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, 1));
        mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 2)); mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, 2));
        const float ea0 = __expf(a0 - mg), ea1 = __expf(a1 - mg), eb0 = __expf(b0 - mh), eb1 = __expf(b1 - mh);
        float sg = ea0 + ea1, sh = eb0 + eb1;
        sg += __shfl_xor_sync(0xffffffffu, sg, 1); sh += __shfl_xor_sync(0xffffffffu, sh, 1);
        sg += __shfl_xor_sync(0xffffffffu, sg, 2); sh += __shfl_xor_sync(0xffffffffu, sh, 2);
        const float e0g = __expf(U.s00[0][h] - mg), e0h = __expf(U.s00[1][h] - mh);
        const float ig = __fdividef(1.f, sg + e0g), ih = __fdividef(1.f, sh + e0h);
        x0[h] = pack2<FMT>(e0g * ig, e0h * ih);
        pa[h >> 1][(h & 1) * 2] = pack2<FMT>(ea0 * ig, ea1 * ig);          // A fragment: K column 8 (h & 1) + token, k-step h >> 1
        pa[h >> 1][(h & 1) * 2 + 1] = pack2<FMT>(eb0 * ih, eb1 * ih);
      }
    }
    prefetch_tail(0);
    // ---- user row: xc_0 + [p_0b] . Uc + Wc (p_00 (.) v_u)
#pragma unroll
    for (int nn = 0; nn < 8; ++nn) {
      const uint4 v = __ldcg(scr + SCR_XC0 + nn * 32 + lane);
      Y[nn][0] = __uint_as_float(v.x); Y[nn][1] = __uint_as_float(v.y); Y[nn][2] = __uint_as_float(v.z); Y[nn][3] = __uint_as_float(v.w);
    }
    attn_rec_mma<FMT>(Y, pa[0], rec + REC_U, lane);
    attn_rec_mma<FMT>(Y, pa[1], rec + REC_U + 128, lane);
    attn_wo_pass<FMT>(Y, x0, U, wo, lane);
    attn_norm_acc<true>(Y, acc);
  }
  // ---- item rows
  const uint32_t one2 = FMT == FMT_BF16 ? 0x3F803F80u : 0x3C003C00u;
#pragma unroll 1
  for (int a = 0; a < nt; ++a) {
    if (a + 1 < nt) prefetch_tail(a + 1);
    uint32_t x[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) x[h] = __shfl_sync(0xffffffffu, (a & 1) ? wpk[1][h] : wpk[0][h], (lane & ~3) | (a >> 1));
    // tail k-step, K columns: 0-3 (1 - w_h) -> Nc_h hi, 4 / 5: 1 -> xc hi / lo, 8-11 (1 - w_h) -> Nc_h lo
    {
      const float2 fa = unpack2<FMT>((t & 1) ? x[2] : x[0]), fb = unpack2<FMT>((t & 1) ? x[3] : x[1]);   // .x: row g, .y: row g + 8
      uint32_t r0 = pack2<FMT>(1.f - fa.x, 1.f - fb.x), r1 = pack2<FMT>(1.f - fa.y, 1.f - fb.y);
      if (t == 2) { r0 = one2; r1 = one2; }
      if (t == 3) { r0 = 0u; r1 = 0u; }
      const uint32_t at[4] = {r0, r1, t < 2 ? r0 : 0u, t < 2 ? r1 : 0u};
#pragma unroll
      for (int nn = 0; nn < 8; ++nn) { Y[nn][0] = 0.f; Y[nn][1] = 0.f; Y[nn][2] = 0.f; Y[nn][3] = 0.f; }
      attn_rec_mma<FMT>(Y, at, rec + REC_T + a * 128, lane);
    }
    attn_wo_pass<FMT>(Y, x, U, wo, lane);
    attn_norm_acc<false>(Y, acc);
  }
#if PXR_ATT_PFS
  if (rec_next && lane < 17) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(rec_next + REC_S) + lane * 128));
#endif
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// FUS selects the front end.  GATED below means "layer 1 runs on the tensor pipe from an A1 tile in shared memory"
// (gated and attention fusion: the fused vector depends on the pair); concat feeds layer-1 partial sums instead.
// TK2: two top-K warps (short units, where list updates are a visible share of the work) instead of one.
// ACT: fusion_activation of the hidden layers (pxr_act; ReLU is the fast default, see act_pack).
// MODE: M_PAGED = pages p > 0 of a top_k > 64 call (Params::upper); M_SPREAD = small user batches (<= 8 users, gated front
// end): the 16 user slots of a unit are (user, item sub-range) pairs, see "small batches" at the gated front end.  Separate
// instantiations, so that this code costs the plain kernels nothing (their register allocation is tight: any extra live value
// shows up as spills in the epilogue).
template <int FUS, int FMT, bool TK2, int ACT, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(n_threads<FUS>(), 1)
score_fused_kernel(const __grid_constant__ Params p) {
  constexpr int NT = n_threads<FUS>();
  constexpr bool GATED = l1_on_tensor_pipe<FUS>();
  constexpr bool ATT = (FUS == F_ATTN);
  constexpr bool WIDE = (FUS == F_GATEDW);       // gated fusion on the concat pipeline: gate-weighted layer-1 partials
  constexpr bool PAGED = (MODE == M_PAGED), SPREAD = (MODE == M_SPREAD);
  static_assert(!SPREAD || FUS == F_GATED || FUS == F_CONCAT, "small-batch mode exists for the gated and concat front ends");
  constexpr int TU = tile_users<FUS>(), TI = tile_items<FUS>();
  using MP = Map<FUS>;
  using MiscF = MiscT<FUS>;
  constexpr int QC = MiscF::QC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);
  MiscF& ms = *reinterpret_cast<MiscF*>(sm + MP::OFF_MISC);
  const uint32_t bar0 = ptx::smem_u32(&ms.bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  // ------------------------------------------------------------------ setup
  if (!ATT) for (int i = threadIdx.x; i < H1 + H2 + H3 + H3; i += NT) ms.b1[i] = p.bias[i];   // b1,b2,b3,w4 are contiguous
  // epilogue constants: shared memory, or (attention) the kernel-parameter constant bank
  // (expressions, not variables: a pointer kept live across the whole kernel costs the epilogue registers)
#define PXR_CB1 (ATT ? p.bias_c : ms.b1)
#define PXR_CB2 (ATT ? p.bias_c + H1 : ms.b2)
#define PXR_CB3 (ATT ? p.bias_c + H1 + H2 : ms.b3)
#define PXR_CW4 (ATT ? p.bias_c + H1 + H2 + H3 : ms.w4)
  if (threadIdx.x == 0) {
    ms.b4 = ATT ? p.bias_c[H1 + H2 + H3 + H3] : p.bias[H1 + H2 + H3 + H3];
    ms.q_tail[0] = ms.q_tail[1] = 0; ms.q_head[0] = ms.q_head[1] = 0;
    ptx::mbar_init(BAR(BAR_W), 1);
    ptx::mbar_init(BAR(BAR_A_FULL), ATT ? 16 : 8);        // one arrival per front-end warp of both CTAs
    ptx::mbar_init(BAR(BAR_A_EMPTY), 1);
    for (int b = 0; b < 4; ++b) {
      ptx::mbar_init(BAR(BAR_D1_FULL0 + b), 1);
      ptx::mbar_init(BAR(BAR_H1_FULL0 + b), 8);
      ptx::mbar_init(BAR(BAR_H1_EMPTY0 + b), 1);
    }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(BAR(BAR_PI_FULL0 + b), 1); ptx::mbar_init(BAR(BAR_PI_EMPTY0 + b), 8); }
    for (int b = 0; b < 3; ++b) { ptx::mbar_init(BAR(BAR_Q_FULL0 + b), 1); ptx::mbar_init(BAR(BAR_Q_EMPTY0 + b), 4); }   // wide gated: 4 warps read a chunk
    ptx::mbar_init(BAR(BAR_D2_FULL), 1);
    ptx::mbar_init(BAR(BAR_H2_FULL0), 8); ptx::mbar_init(BAR(BAR_H2_FULL1), 8);
    ptx::mbar_init(BAR(BAR_D3_FULL), 1);
    ptx::mbar_init(BAR(BAR_D3_EMPTY), 8);
    ptx::mbar_init(BAR(BAR_UNIT_DONE), 4);
    ptx::mbar_init(BAR(BAR_UNIT_RESET), TK2 ? 2 : 1);   // one arrival per active top-K warp
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < TU * KCAP; i += NT) (&ms.list[0][0])[i] = 0ull;
  for (int i = threadIdx.x; i < QC; i += NT) ms.queue[i] = 0ull;
  if (threadIdx.x < TU) {
    ms.thr[threadIdx.x] = -INFINITY;
    if (PAGED) {
      const int64_t ord = ((int64_t)(pair / p.S) * 2 + rank) * TU + threadIdx.x;     // user slot of this CTA's first unit
      ms.upper[threadIdx.x] = ord < p.n_users ? p.upper[ord] : ~0ull;
    }
  }
  __syncthreads();
  if (warp == 4) {
    if (lane == 0) {   // this CTA's half of every weight matrix: 16 KB bulk copies through the TMA engine
      ptx::mbar_expect_tx(BAR(BAR_W), MP::WIMG);
      const uint8_t* src = p.wimg + (size_t)rank * MP::WIMG;
      for (uint32_t o = 0; o < MP::WIMG; o += 16384) ptx::bulk_g2s(base + o, src + o, 16384, BAR(BAR_W));
    }
    __syncwarp();
    ptx::tmem_alloc_2cta(ptx::smem_u32(&ms.tmem_base), 512);
    if (lane == 0) ptx::mbar_wait(BAR(BAR_W), 0);
    __syncwarp();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();          // both CTAs: barriers initialised, weights resident, TMEM allocated
  ptx::tc_fence_after();
  const uint32_t tmem = ms.tmem_base;

  // units of this pair: w = pair, pair + n_pairs, ...; every role walks the same (unit, tile) sequence;
  // T counts tiles over all units of the pair

  // attention: register budget per warpgroup.  640 threads are launched with 96 registers each and setmaxnreg can only move
  // registers inside that pool (61 440): the MMA / top-K warpgroup drops to 40, the two front-end warpgroups take 128, the
  // epilogue warpgroups drop to 88  (2*128*128 + 128*40 + 2*128*88 = 60 416).  setmaxnreg sits at the top of each
  // warpgroup's branch so that ptxas allocates per role.
  if (warp < 4 || (ATT && warp >= 16)) {
    if (ATT) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    // =============================================================== front end
    // attention: eight front-end warps (warpgroups A = warps 0-3 and B = warps 16-19), one item of the tile each, two per
    // scheduler; warpgroup A also builds the per-unit user operands; named barriers 2 / 3 fence them between units.
    const bool feB = ATT && warp >= 16;
    const int fw = feB ? warp - 12 : warp;  // front-end warp 0..7 (attention) / 0..3
    const int tid = threadIdx.x & 127;      // 0..127 inside the front-end warpgroup
    const int Mm = p.M;
    int T = 0;
    AttnUserFrag UF;                        // attention: per-unit user operands of this thread (registers)
    uint4* const scr = ATT ? p.xc0_scratch + (size_t)blockIdx.x * ATT_XC0_U4 : nullptr;
    for (int w = pair; w < p.n_units; w += n_pairs) {
      const Unit un = decode_unit<TI, SPREAD>(p, w);
      const int64_t ubase = ((int64_t)un.g * 2 + rank) * TU;     // first user ordinal of this CTA's group
      if (!GATED && T > 0) {
        // Pu is read by the layer-1 producers of the previous unit's last tile: wait until they are done with it
        ptx::mbar_wait(BAR(BAR_PI_EMPTY0 + ((T - 1) & 1)), ((T - 1) >> 1) & 1);
      }
      if (ATT) asm volatile("bar.sync 2, 256;" ::: "memory");      // both warpgroups are done with the previous unit's scratch
      if (!feB) {
      asm volatile("bar.sync 1, 128;" ::: "memory");              // previous unit's readers of eu/lu are done
      if (ATT) {
        // the per-unit user operands are staged in the A1 tile: wait until the layer-1 MMAs of the previous tile have read it
        if (T > 0) ptx::mbar_wait(BAR(BAR_A_EMPTY), (T - 1) & 1);
        attn_user_setup<FMT>(p, reinterpret_cast<float*>(sm + MP::OFF_A1), tid, ubase, scr);
      } else {
        const int u = tid >> 4, d4 = (tid & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t uo = SPREAD ? (ubase + u) % p.n_real : ubase + u;      // SPREAD: slot (ubase + u) = real user uo on sub-range (ubase + u) / n_real
        if (ubase + u < p.n_users && d4 < p.Dm) v = *reinterpret_cast<const float4*>(p.user_emb + p.user_idx[uo] * p.Dm + d4);
        *reinterpret_cast<float4*>(&ms.eu[u][d4]) = v;
        if (SPREAD && tid < TU) {            // the slot's item sub-range (rows of the shard)
          const int64_t r = (ubase + tid) / p.n_real;
          const int64_t lo = un.row_lo + r * p.sub_rows;
          ms.slot_lo[tid] = (int32_t)min(lo, un.row_hi);
          ms.slot_hi[tid] = (ubase + tid < p.n_users) ? (int32_t)min(lo + p.sub_rows, un.row_hi) : (int32_t)min(lo, un.row_hi);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (FUS == F_ATTN) {
      } else if (FUS == F_GATED) {
        if (tid < 64) {                      // user part of the gate logits (layers.py:207 split per SURVEY A4)
          const int u = tid >> 3, m = tid & 7;
          float acc = 0.f;
          if (m < Mm) {
            const float* wr = p.gate_w + (size_t)m * Mm * D;
#pragma unroll 8
            for (int d = 0; d < D; ++d) acc += wr[d] * ms.eu[u][d];
          }
          ms.lu[u][m] = acc;
        }
      } else {
        // per-user partial of layer 1: Pu[u][n] = sum_{k < D} W1[n][k] Eu[u][k]   (SURVEY.md A3), fp32.  The user
        // embedding is staged 64 dims at a time, so any embedding_dim works (the MMA chain does not depend on it).
        float acc[TU][4];
#pragma unroll
        for (int u = 0; u < TU; ++u) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f; }
        const float* wcol = p.w1u_t + 4 * tid;
        float lacc = 0.f;                    // wide gated: user part of gate logit m of user u (layers.py:207 split), tid = 8 u + m
        for (int k0 = 0;; k0 += D) {
          const int kn = min(D, p.Dm - k0);
          if (WIDE && tid < 64 && (tid & 7) < Mm) {
            const float* wr = p.gate_w + (size_t)(tid & 7) * Mm * p.Dm + k0;
            for (int k = 0; k < kn; ++k) lacc = fmaf(wr[k], ms.eu[tid >> 3][k], lacc);
          }
#pragma unroll 4
          for (int k = 0; k < kn; ++k) {
            const float4 wv = *reinterpret_cast<const float4*>(wcol + (size_t)(k0 + k) * H1);
#pragma unroll
            for (int u = 0; u < TU; ++u) {
              const float e = ms.eu[u][k];
              acc[u][0] = fmaf(e, wv.x, acc[u][0]); acc[u][1] = fmaf(e, wv.y, acc[u][1]);
              acc[u][2] = fmaf(e, wv.z, acc[u][2]); acc[u][3] = fmaf(e, wv.w, acc[u][3]);
            }
          }
          if (k0 + D >= p.Dm) break;
          asm volatile("bar.sync 1, 128;" ::: "memory");            // everyone is done with this 64-dim slice of E_u
          {
            const int u = tid >> 4, d4 = k0 + D + (tid & 15) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int64_t uo2 = SPREAD ? (ubase + u) % p.n_real : ubase + u;
            if (ubase + u < p.n_users && d4 < p.Dm) v = *reinterpret_cast<const float4*>(p.user_emb + p.user_idx[uo2] * p.Dm + d4);
            *reinterpret_cast<float4*>(&ms.eu[u][(tid & 15) * 4]) = v;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (WIDE) {                          // the per-user partial carries b1 (every partial does: the gate weights sum to 1)
          const float4 bb = *reinterpret_cast<const float4*>(&ms.b1[4 * tid]);
#pragma unroll
          for (int u = 0; u < TU; ++u) { acc[u][0] += bb.x; acc[u][1] += bb.y; acc[u][2] += bb.z; acc[u][3] += bb.w; }
          if (tid < 64) ms.lu[tid >> 3][tid & 7] = lacc;
        }
#pragma unroll
        for (int u = 0; u < TU; ++u)
          *reinterpret_cast<float4*>(sm + MP::OFF_PU + u * MP::PU_STRIDE + 16 * tid) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (ATT) {
        asm volatile("bar.sync 3, 256;" ::: "memory");            // the user operands of this unit are in the scratch
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const uint4 f = __ldcg(scr + SCR_VU + h * 32 + lane);
          UF.vu[h][0] = f.x; UF.vu[h][1] = f.y; UF.vu[h][2] = f.z; UF.vu[h][3] = f.w;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint4 f = __ldcg(scr + SCR_S00 + r * 32 + lane);
          UF.s00[r][0] = __uint_as_float(f.x); UF.s00[r][1] = __uint_as_float(f.y); UF.s00[r][2] = __uint_as_float(f.z); UF.s00[r][3] = __uint_as_float(f.w);
        }
      }
      // seen-item cursors: lanes 0..TU-1 of warp 0 walk user u's ascending history with the item sweep
      int64_t cur = 0, cend = 0; int32_t nextv = 0x7fffffff;
      if (warp == 0 && lane < TU && p.seen_indptr && ubase + lane < p.n_users) {
        const int64_t uo = SPREAD ? (ubase + lane) % p.n_real : ubase + lane;
        cur = p.seen_indptr[uo]; cend = p.seen_indptr[uo + 1];
        const int32_t first = (int32_t)(p.item_base + (SPREAD ? (int64_t)ms.slot_lo[lane] : un.row_lo));
        int64_t lo = cur, hi = cend;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (p.seen_idx[mid] < first) lo = mid + 1; else hi = mid; }
        cur = lo;
        nextv = cur < cend ? p.seen_idx[cur] : 0x7fffffff;
      }
      if (!GATED && warp != 0) { T += un.ntiles; continue; }     // concat: warp 0 alone stages the tiles
      for (int t = 0; t < un.ntiles; ++t, ++T) {
        const int64_t row0 = un.row_lo + (int64_t)t * TI;
        auto write_seen_mask = [&]() {         // lanes 0..TU-1 of warp 0: the TI-bit seen mask of this tile per user
          if (warp == 0 && lane < TU) {
            uint32_t mask = 0;
            const int32_t i0 = (int32_t)(p.item_base + (SPREAD ? (int64_t)ms.slot_lo[lane] + (int64_t)t * TI : row0));
            while (nextv < i0 + TI) {
              if (nextv >= i0) mask |= 1u << (nextv - i0);
              ++cur;
              nextv = cur < cend ? p.seen_idx[cur] : 0x7fffffff;
            }
            ms.seen_mask[T & 3][lane] = mask;
          }
        };
        if (FUS == F_ATTN) {
          write_seen_mask();
          {                                                           // this warp's item of the tile x the CTA's 16 users
            const int j = fw;
            const int64_t row = row0 + j;
            const int64_t rr = row < un.row_hi ? row : un.row_lo;    // padding rows recompute a valid item (discarded later)
            const int64_t rn = row + TI;                             // the item this warp takes in the next tile
            float acc[8][4];
            attn_item_step<FMT>(UF, p.attn_rec + rr * ATT_REC_U4, rn < un.row_hi ? p.attn_rec + rn * ATT_REC_U4 : nullptr, p.wo_frag,
                                scr, Mm - 1, lane, acc);
            if (T > 0) ptx::mbar_wait(BAR(BAR_A_EMPTY), (T - 1) & 1);   // layer-1 MMAs of the previous tile have read A1
            // 16-bit pack into the swizzled A1 rows: row = item * 16 + user, user rows g and g + 8 of this lane
            const int g = lane >> 2, t4 = (lane & 3) * 4;
            uint8_t* a1 = sm + MP::OFF_A1 + (2 * j) * 1024 + g * 128 + t4;
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) {
              *reinterpret_cast<uint32_t*>(a1 + ((nn ^ g) << 4)) = pack2<FMT>(acc[nn][0], acc[nn][1]);
              *reinterpret_cast<uint32_t*>(a1 + 1024 + ((nn ^ g) << 4)) = pack2<FMT>(acc[nn][2], acc[nn][3]);
            }
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster_release(BAR(BAR_A_FULL), 0);
        } else if (FUS == F_GATED && SPREAD) {
          // ---- small batches: slot u of the CTA sweeps its own item sub-range, so the 16 items of a tile differ per slot:
          // thread (j, s) computes the gate of pair (slot s, that slot's item j) and then, per slot u, fetches item j of
          // slot u (software-pipelined one slot ahead: the records come from L2 / HBM once per pair here, not once per 8 pairs)
          const int j = tid >> 3, s = tid & 7;
          const int toff = t * TI + j;
          auto load_item = [&](int u, float (&f)[5][8]) {
            const int row = ms.slot_lo[u] + toff;
            const bool valid = row < ms.slot_hi[u];
#pragma unroll
            for (int m = 0; m < 5; ++m) {
              if (m < Mm - 1 && valid) {
                const float4* src = reinterpret_cast<const float4*>(p.item_feats + ((int64_t)row * (Mm - 1) + m) * D + 8 * s);
                const float4 a = __ldg(src), b = __ldg(src + 1);
                f[m][0] = a.x; f[m][1] = a.y; f[m][2] = a.z; f[m][3] = a.w; f[m][4] = b.x; f[m][5] = b.y; f[m][6] = b.z; f[m][7] = b.w;
              } else {
#pragma unroll
                for (int d = 0; d < 8; ++d) f[m][d] = 0.f;
              }
            }
          };
          float fa[5][8], fb[5][8];
          load_item(0, fa);
          float g[6];
          {
            const int row = ms.slot_lo[s] + toff;
            const int64_t rr = row < ms.slot_hi[s] ? row : un.row_lo;
            const float4 l0 = __ldg(reinterpret_cast<const float4*>(p.item_logit + rr * 8));
            const float4 l1 = __ldg(reinterpret_cast<const float4*>(p.item_logit + rr * 8 + 4));
            const float li[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            float mx = -INFINITY;
#pragma unroll
            for (int m = 0; m < 6; ++m) { g[m] = m < Mm ? li[m] + ms.lu[s][m] : -INFINITY; mx = fmaxf(mx, g[m]); }
            float sum = 0.f;
#pragma unroll
            for (int m = 0; m < 6; ++m) { g[m] = m < Mm ? expf(g[m] - mx) : 0.f; sum += g[m]; }
            const float inv = 1.f / sum;
#pragma unroll
            for (int m = 0; m < 6; ++m) g[m] *= inv;
          }
          write_seen_mask();
          if (T > 0) ptx::mbar_wait(BAR(BAR_A_EMPTY), (T - 1) & 1);   // layer-1 MMAs of the previous tile have read A1
          auto emit = [&](int u, const float (&f)[5][8]) {
            float gm[6];
#pragma unroll
            for (int m = 0; m < 6; ++m) gm[m] = __shfl_sync(0xffffffffu, g[m], (lane & ~7) | u);
            const float4 e0 = *reinterpret_cast<const float4*>(&ms.eu[u][8 * s]);
            const float4 e1 = *reinterpret_cast<const float4*>(&ms.eu[u][8 * s + 4]);
            float acc[8] = {gm[0] * e0.x, gm[0] * e0.y, gm[0] * e0.z, gm[0] * e0.w, gm[0] * e1.x, gm[0] * e1.y, gm[0] * e1.z, gm[0] * e1.w};
#pragma unroll
            for (int m = 0; m < 5; ++m)
#pragma unroll
              for (int d = 0; d < 8; ++d) acc[d] = fmaf(gm[m + 1], f[m][d], acc[d]);
            uint4 pk;
            pk.x = pack2<FMT>(acc[0], acc[1]); pk.y = pack2<FMT>(acc[2], acc[3]);
            pk.z = pack2<FMT>(acc[4], acc[5]); pk.w = pack2<FMT>(acc[6], acc[7]);
            const int r = u * TI + j;
            *reinterpret_cast<uint4*>(sm + MP::OFF_A1 + (r >> 3) * 1024 + (r & 7) * 128 + ((s ^ (r & 7)) << 4)) = pk;
          };
#pragma unroll 1
          for (int u = 0; u < TU; u += 2) {
            load_item(u + 1, fb);
            emit(u, fa);
            if (u + 2 < TU) load_item(u + 2, fa);
            emit(u + 1, fb);
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster_release(BAR(BAR_A_FULL), 0);
        } else if (FUS == F_GATED) {
          const int j = tid >> 3, s = tid & 7;    // item of the tile, 8-wide slice of D
          const int64_t row = row0 + j;
          const bool valid = row < un.row_hi;
          const int64_t rr = valid ? row : un.row_lo;
          // item-side modality vectors for dims [8s, 8s+8) and the item part of the gate logits
          float f[5][8];
#pragma unroll
          for (int m = 0; m < 5; ++m) {
            if (m < Mm - 1 && valid) {
              const float4* src = reinterpret_cast<const float4*>(p.item_feats + (rr * (Mm - 1) + m) * D + 8 * s);
              const float4 a = src[0], b = src[1];
              f[m][0] = a.x; f[m][1] = a.y; f[m][2] = a.z; f[m][3] = a.w; f[m][4] = b.x; f[m][5] = b.y; f[m][6] = b.z; f[m][7] = b.w;
            } else {
#pragma unroll
              for (int d = 0; d < 8; ++d) f[m][d] = 0.f;
            }
          }
          // gate of pair (user s, item j): softmax over the M modality logits (layers.py:207-211)
          float g[6];
          {
            const float4 l0 = *reinterpret_cast<const float4*>(p.item_logit + rr * 8);
            const float4 l1 = *reinterpret_cast<const float4*>(p.item_logit + rr * 8 + 4);
            const float li[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            float mx = -INFINITY;
#pragma unroll
            for (int m = 0; m < 6; ++m) { g[m] = m < Mm ? li[m] + ms.lu[s][m] : -INFINITY; mx = fmaxf(mx, g[m]); }
            float sum = 0.f;
#pragma unroll
            for (int m = 0; m < 6; ++m) { g[m] = m < Mm ? expf(g[m] - mx) : 0.f; sum += g[m]; }
            const float inv = 1.f / sum;
#pragma unroll
            for (int m = 0; m < 6; ++m) g[m] *= inv;
          }
          write_seen_mask();
          if (T > 0) ptx::mbar_wait(BAR(BAR_A_EMPTY), (T - 1) & 1);   // layer-1 MMAs of the previous tile have read A1
#pragma unroll
          for (int u = 0; u < TU; ++u) {
            float gm[6];
#pragma unroll
            for (int m = 0; m < 6; ++m) gm[m] = __shfl_sync(0xffffffffu, g[m], (lane & ~7) | u);
            const float4 e0 = *reinterpret_cast<const float4*>(&ms.eu[u][8 * s]);
            const float4 e1 = *reinterpret_cast<const float4*>(&ms.eu[u][8 * s + 4]);
            float acc[8] = {gm[0] * e0.x, gm[0] * e0.y, gm[0] * e0.z, gm[0] * e0.w, gm[0] * e1.x, gm[0] * e1.y, gm[0] * e1.z, gm[0] * e1.w};
#pragma unroll
            for (int m = 0; m < 5; ++m)
#pragma unroll
              for (int d = 0; d < 8; ++d) acc[d] = fmaf(gm[m + 1], f[m][d], acc[d]);
            uint4 pk;
            pk.x = pack2<FMT>(acc[0], acc[1]); pk.y = pack2<FMT>(acc[2], acc[3]);
            pk.z = pack2<FMT>(acc[4], acc[5]); pk.w = pack2<FMT>(acc[6], acc[7]);
            const int r = u * TI + j;
            *reinterpret_cast<uint4*>(sm + MP::OFF_A1 + (r >> 3) * 1024 + (r & 7) * 128 + ((s ^ (r & 7)) << 4)) = pk;
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster_release(BAR(BAR_A_FULL), 0);
        } else {
          // stage the 16 item partials of this tile: one 1 KB TMA bulk copy per item row (padded rows in smem).
          // The whole warp waits for the buffer, so the seen masks never run more than 2 tiles ahead either.
          const int buf = T & 1;
          if (T >= 2) ptx::mbar_wait(BAR(BAR_PI_EMPTY0 + buf), ((T >> 1) - 1) & 1);
          write_seen_mask();
          if (WIDE) {
            // M - 1 partials per item = 80 KB per tile: staged one 64-column chunk (10 KB, contiguous in the chunk-major
            // layout) at a time through a ring of three buffers, one TMA bulk copy each; the tile itself (seen masks, the
            // unit's user constants) is announced first
            if (tid == 0) {
              ptx::mbar_arrive_local(BAR(BAR_PI_FULL0 + buf));
              const uint32_t stage = q_stage_bytes(Mm - 1);
              const uint8_t* src = reinterpret_cast<const uint8_t*>(p.item_q) + (size_t)(row0 >> 4) * 8 * stage;
#pragma unroll 1
              for (int c = 0; c < 8; ++c) {
                const uint32_t G = (uint32_t)T * 8u + c, slot = G % MP::Q_STAGES, n = G / MP::Q_STAGES;
                if (n > 0) ptx::mbar_wait(BAR(BAR_Q_EMPTY0 + slot), (n - 1) & 1);      // its four readers are done with the slot
                ptx::mbar_expect_tx(BAR(BAR_Q_FULL0 + slot), stage);
                ptx::bulk_g2s(base + MP::OFF_PI + slot * MP::Q_STAGE_MAX, src + (size_t)c * stage, stage, BAR(BAR_Q_FULL0 + slot));
              }
            }
            __syncwarp();
          } else if (SPREAD) {
            // small batches: the 128 rows of a tile are 128 different items, more than the staging buffers hold; the
            // layer-1 producers read each row's partial straight from L2 (prefetched a tile ahead), nothing to stage
            if (tid == 0) ptx::mbar_arrive_local(BAR(BAR_PI_FULL0 + buf));
          } else if (tid == 0) {
            const int nvalid = (int)min((int64_t)TI, un.row_hi - row0);
            ptx::mbar_expect_tx(BAR(BAR_PI_FULL0 + buf), (uint32_t)nvalid * 1024u);
            for (int jj = 0; jj < nvalid; ++jj)
              ptx::bulk_g2s(base + MP::OFF_PI + buf * MP::PI_BUF + jj * MP::PI_STRIDE, p.item_pi + (row0 + jj) * H1, 1024,
                            BAR(BAR_PI_FULL0 + buf));
          }
        }
      }
    }
  } else if (warp < 8) {
   if (ATT) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
   if (warp == 4) {
    // =============================================================== MMA issuer (leader CTA, one thread)
    if (rank == 0 && lane == 0) {
      int NT = 0;
      for (int w = pair; w < p.n_units; w += n_pairs) NT += decode_unit<TI, SPREAD>(p, w).ntiles;
      const uint64_t dA1 = ptx::smem_desc_sw128(base + MP::OFF_A1);
      const uint64_t dW1 = ptx::smem_desc_sw128(base + MP::OFF_W1);
      const uint64_t dW2 = ptx::smem_desc_sw128(base + MP::OFF_W2);
      const uint64_t dW3 = ptx::smem_desc_sw128(base + MP::OFF_W3);
      constexpr uint32_t I1 = idesc<FMT>(256, 64), I2 = idesc<FMT>(256, 256), I3 = idesc<FMT>(256, 128);
      uint32_t h1ph[4] = {0, 0, 0, 0};
      auto issue_m1 = [&](int c) {        // gated: D1[c % 4] = A1 . W1[chunk c]^T   (K = 64: 4 steps of 16)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::mma2_ss(tmem + MP::h1buf(c & 3), dA1 + 2 * k, dW1 + (uint64_t)(c * 4096 >> 4) + 2 * k, I1, k > 0);
        ptx::commit2_mc(BAR(BAR_D1_FULL0 + (c & 3)), 3);
      };
      auto issue_m2 = [&](int c) {        // D2 += H1[c] . W2[:, 64c .. 64c+63]^T
        const int b = c & 3;
        ptx::mbar_wait(BAR(BAR_H1_FULL0 + b), h1ph[b]); h1ph[b] ^= 1;
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::mma2_ts(tmem + MP::TM_D2, tmem + MP::h1buf(b) + 8 * k, dW2 + (uint64_t)(c * 16384 >> 4) + 2 * k, I2, (c > 0 || k > 0));
        if (!GATED) ptx::commit2_mc(BAR(BAR_H1_EMPTY0 + b), 3);   // concat: CUDA cores refill the buffer directly
      };
      auto issue_m3 = [&](int Tprev) {    // D3 = H2 . W3^T   (K = 256 in two halves as the H2 halves arrive)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          ptx::mbar_wait(BAR(BAR_H2_FULL0 + half), Tprev & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int kk = half * 8 + k;
            ptx::mma2_ts(tmem + MP::TM_D3, tmem + MP::TM_D2 + half * 128 + 8 * k,
                         dW3 + (uint64_t)((kk >> 2) * 8192 >> 4) + 2 * (kk & 3), I3, kk > 0);
          }
        }
        ptx::commit2_mc(BAR(BAR_D3_FULL), 3);
      };
      // Issue order per tile (the tensor pipe executes in issue order).
      // gated: chunk buffers 2,3 share the D3 columns, so layer-1 chunks 2,3 of tile T wait until the layer-3
      // epilogue of tile T-1 has drained D3; every other layer-1 chunk is issued as soon as its buffer's previous
      // H1 chunk has been consumed, 2-4 chunks ahead of the layer-2 MMA that needs it.
      for (int T = 0; T < NT; ++T) {
        if (GATED) {
          ptx::mbar_wait_cluster(BAR(BAR_A_FULL), T & 1);
          ptx::tc_fence_after();
          issue_m1(0);
          issue_m1(1);
          if (T > 0) issue_m3(T - 1);
          issue_m2(0); issue_m1(4);
          if (T > 0) { ptx::mbar_wait(BAR(BAR_D3_EMPTY), (T - 1) & 1); ptx::tc_fence_after(); }
          issue_m1(2);
          issue_m1(3);
          issue_m2(1); issue_m1(5);
          issue_m2(2); issue_m1(6);
          issue_m2(3); issue_m1(7);
          ptx::commit2_mc(BAR(BAR_A_EMPTY), 3);
          issue_m2(4); issue_m2(5); issue_m2(6); issue_m2(7);
        } else {
          if (T > 0) {
            if (T > 1) { ptx::mbar_wait(BAR(BAR_D3_EMPTY), (T - 2) & 1); ptx::tc_fence_after(); }
            issue_m3(T - 1);
          }
#pragma unroll 1
          for (int c = 0; c < 8; ++c) issue_m2(c);
        }
        ptx::commit2_mc(BAR(BAR_D2_FULL), 3);
      }
      if (NT > 0) {
        if (!GATED && NT > 1) { ptx::mbar_wait(BAR(BAR_D3_EMPTY), (NT - 2) & 1); ptx::tc_fence_after(); }
        issue_m3(NT - 1);
      }
    }
   } else if (warp == 5 || (warp == 6 && TK2)) {
    // =============================================================== top-K warp(s)
    // TK2: warp 5 owns users 0-3 and warp 6 users 4-7, one queue each (see launch_fused for when)
    const int qh = warp - 5;
    constexpr uint32_t QM = (TK2 ? QC / 2 : QC) - 1;      // queue capacities are powers of two
    const int u_lo = TK2 ? (TU / 2) * qh : 0, u_hi = TK2 ? (TU / 2) * (qh + 1) : TU;
    unsigned long long* const queue = ms.queue + qh * (QC / 2);
    uint32_t head = 0, done_ph = 0;
    for (int w = pair; w < p.n_units; w += n_pairs) {
      const Unit un = decode_unit<TI, SPREAD>(p, w);
      if (un.ntiles == 0) continue;
      const int64_t ubase = ((int64_t)un.g * 2 + rank) * TU;
      bool finished = false;
      while (true) {
        unsigned long long e = 0ull;
        if (lane == 0) e = *reinterpret_cast<volatile unsigned long long*>(&queue[head & QM]);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e != 0ull) {
          __syncwarp();
          if (lane == 0) {
            *reinterpret_cast<volatile unsigned long long*>(&queue[head & QM]) = 0ull;
            *reinterpret_cast<volatile uint32_t*>(&ms.q_head[qh]) = head + 1;
          }
          ++head;
          const int u = (int)((e >> 28) & (unsigned long long)(TU - 1));
          const unsigned long long key = e & 0xFFFFFFFF0FFFFFFFull;
          // sorted (descending) insertion into list[u]: lanes hold slots lane and lane + 32
          unsigned long long* L = ms.list[u];
          const unsigned long long a = L[lane], b = L[lane + 32];
          const unsigned ba = __ballot_sync(0xffffffffu, key > a), bb = __ballot_sync(0xffffffffu, key > b);
          const int pos = ba ? 32 - __popc(ba) : 64 - __popc(bb);     // entries >= key come first
          // page bound (top_k > 64): entries of the previous pages pass the score threshold again and are dropped here
          if (pos < p.K && (!PAGED || key < ms.upper[u])) {
            const unsigned long long a_up = __shfl_up_sync(0xffffffffu, a, 1), b_up = __shfl_up_sync(0xffffffffu, b, 1);
            const unsigned long long a31 = __shfl_sync(0xffffffffu, a, 31);
            const unsigned long long na = lane < pos ? a : (lane == pos ? key : a_up);
            const int lb = lane + 32;
            const unsigned long long nb = lb < pos ? b : (lb == pos ? key : (lane == 0 ? a31 : b_up));
            L[lane] = na; L[lb] = nb;
            const int kth = p.K - 1;          // slot of the K-th best: its score is the admission threshold
            const unsigned long long kv = __shfl_sync(0xffffffffu, kth < 32 ? na : nb, kth & 31);
            if (lane == 0) *reinterpret_cast<volatile float*>(&ms.thr[u]) = kv ? pxr_unord((uint32_t)(kv >> 32)) : -INFINITY;
          }
          __syncwarp();
          continue;
        }
        if (finished) {
          uint32_t tail = 0;
          if (lane == 0) tail = *reinterpret_cast<volatile uint32_t*>(&ms.q_tail[qh]);
          tail = __shfl_sync(0xffffffffu, tail, 0);
          if (head == tail) break;
          continue;
        }
        uint32_t dn = 0;
        if (lane == 0) dn = ptx::mbar_test_wait(BAR(BAR_UNIT_DONE), done_ph) ? 1u : 0u;
        dn = __shfl_sync(0xffffffffu, dn, 0);
        if (dn) { finished = true; done_ph ^= 1; }
        else __nanosleep(PXR_TOPK_IDLE_NS);    // idle: do not burn issue slots / power while the queue is empty
      }
      // write the K best of every user of this unit, then reset for the next unit
      for (int u = u_lo; u < u_hi; ++u) {
        const int64_t ord = ubase + u;
        for (int i = lane; i < KCAP; i += 32) {
          const unsigned long long kv = ms.list[u][i];
          if (ord < p.n_users && i < p.K) {
            const int64_t o = ((int64_t)un.s * p.n_users + ord) * p.K + i;
            p.out_scores[o] = kv ? pxr_unord((uint32_t)(kv >> 32)) : -INFINITY;
            p.out_idx[o] = kv ? (int32_t)(IDX_MASK - (uint32_t)(kv & IDX_MASK)) : -1;
          }
          ms.list[u][i] = 0ull;
        }
        if (lane == 0) ms.thr[u] = -INFINITY;
      }
      if (PAGED && w + n_pairs < p.n_units && lane >= u_lo && lane < u_hi) {          // the next unit's page bounds
        const int64_t ord = ((int64_t)((w + n_pairs) / p.S) * 2 + rank) * TU + lane;
        ms.upper[lane] = ord < p.n_users ? p.upper[ord] : ~0ull;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_local(BAR(BAR_UNIT_RESET));
    }
   }
  } else {
    // =============================================================== epilogue groups (warps 8-15)
    if (ATT) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    const int grp = (warp - 8) >> 2;                 // 0: even layer-1 chunks, first half of layer 2, layer 3
    const int q = warp & 3;                          // TMEM lane quarter this warp may touch
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const int r = q * 32 + lane;                     // row of the tile this thread owns
    const int ru = ATT ? (r & 15) : (r >> 4), rj = ATT ? (r >> 4) : (r & 15);   // user slot / item slot of the row
    uint32_t d1ph = 0, reset_ph = 0;                 // d1ph: phase bits of chunk buffers grp (bit 0) and grp + 2 (bit 1)
    uint32_t h1use0 = 0, h1use1 = 0;                 // concat: uses so far of chunk buffers grp and grp + 2
    int T = 0;
    // what the deferred layer-3 epilogue needs to know about the previous tile, kept small (the loop below is short
    // on registers): this thread's item row (-1: padding row or no such user) and two flags
    int64_t prev_row = -1; int prev_flags = 0; bool have_prev = false;      // flags: 1 = first tile of its unit, 2 = last
    // attention (88-register epilogue): ptxas spills less when the previous unit itself is kept and the row is derived
    // where it is used (measured on config C: 1.46 vs 1.36 G pairs/s); gated / concat keep the compact form above
    Unit prev; prev.ntiles = 0; prev.row_lo = prev.row_hi = 0; prev.g = prev.s = 0;
    int prev_t = 0; int64_t prev_ubase = 0;

    auto do_e2 = [&](int Tp) {                       // H2 half `grp` of tile Tp
      ptx::mbar_wait(BAR(BAR_D2_FULL), Tp & 1);
      ptx::tc_fence_after();
      const uint32_t c0 = MP::TM_D2 + grp * 128;
      epi_pack64<ACT, FMT>(tl + c0, tl + c0, PXR_CB2 + grp * 128);
      epi_pack64<ACT, FMT>(tl + c0 + 64, tl + c0 + 32, PXR_CB2 + grp * 128 + 64);
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(BAR(BAR_H2_FULL0 + grp), 0);
    };
    auto do_e3 = [&](int Tp, auto&& row_of, bool last_of_unit) {       // row_of(): this thread's item row or -1, evaluated late
      ptx::mbar_wait(BAR(BAR_D3_FULL), Tp & 1);
      ptx::tc_fence_after();
      float z = ms.b4;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld32(tl + MP::TM_D3 + h * 64, v0);
        ptx::tmem_ld32(tl + MP::TM_D3 + h * 64 + 32, v1);
        ptx::tc_wait_ld();
        const float4* b3v = reinterpret_cast<const float4*>(PXR_CB3 + h * 64);
        const float4* w4v = reinterpret_cast<const float4*>(PXR_CW4 + h * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b = b3v[i], wv = w4v[i];
          const uint32_t* v = i < 8 ? v0 + 4 * i : v1 + 4 * (i - 8);
          z = fmaf(act_fast<ACT>(__uint_as_float(v[0]) + b.x), wv.x, z);
          z = fmaf(act_fast<ACT>(__uint_as_float(v[1]) + b.y), wv.y, z);
          z = fmaf(act_fast<ACT>(__uint_as_float(v[2]) + b.z), wv.z, z);
          z = fmaf(act_fast<ACT>(__uint_as_float(v[3]) + b.w), wv.w, z);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(BAR(BAR_D3_EMPTY), 0);
      float y = pxr_apply_final(z, p.final_act);
      const int64_t row = row_of();
      if (p.item_missing && row >= 0 && p.item_missing[row]) y = 0.f;
      const bool ok = row >= 0 && !((ms.seen_mask[Tp & 3][ru] >> rj) & 1u);
      if (ok && y >= *reinterpret_cast<volatile float*>(&ms.thr[ru])) {
        const uint32_t gidx = (uint32_t)(p.item_base + row);
        const unsigned long long e = ((unsigned long long)pxr_ord(y) << 32) | ((unsigned long long)ru << 28) |
                                     (unsigned long long)(IDX_MASK - gidx);
        const int qh = TK2 ? (ru >= TU / 2 ? 1 : 0) : 0;
        constexpr uint32_t qcap = TK2 ? QC / 2 : QC;
        const uint32_t slot = atomicAdd(&ms.q_tail[qh], 1u);
        while (slot - *reinterpret_cast<volatile uint32_t*>(&ms.q_head[qh]) >= qcap) __nanosleep(64);
        *reinterpret_cast<volatile unsigned long long*>(&ms.queue[qh * (QC / 2) + (slot & (qcap - 1))]) = e;
      }
      if (last_of_unit) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_local(BAR(BAR_UNIT_DONE));
      }
    };
    const uint16_t* pi_g = nullptr;                  // concat, small batches: this row's item partial of the current tile (global)
    // wide gated: the item part of this row's gate logits (loaded at the top of the tile, used once the tile's user
    // constants are known to be in place), its gate weights
    float4 wl0 = make_float4(0.f, 0.f, 0.f, 0.f), wl1 = wl0;
    float wg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    __half2 wgh[5];                                  // PXR_GW_HALF2: the item-side gate weights as packed fp16 pairs
#pragma unroll
    for (int m = 0; m < 5; ++m) wgh[m] = __float2half2_rn(0.f);
    auto prev_e3 = [&]() {
      if (ATT) {
        if (prev_t == 0 && (T - 1) > 0) { ptx::mbar_wait(BAR(BAR_UNIT_RESET), reset_ph); reset_ph ^= 1; }
        do_e3(T - 1, [&]() -> int64_t {
          const int64_t row = prev.row_lo + (int64_t)prev_t * TI + rj;
          return (row < prev.row_hi && (prev_ubase + ru) < p.n_users) ? row : -1;
        }, prev_t == prev.ntiles - 1);
      } else {
        if ((prev_flags & 1) && (T - 1) > 0) { ptx::mbar_wait(BAR(BAR_UNIT_RESET), reset_ph); reset_ph ^= 1; }
        do_e3(T - 1, [&]() -> int64_t { return prev_row; }, (prev_flags & 2) != 0);
      }
    };
    // layer-1 chunk ci (0..3) of this group for the current tile: chunk c = 2 ci + grp in buffer c % 4
    auto do_l1 = [&](int ci) {
      const int c = 2 * ci + grp;
      const int b = c & 3;                           // grp (ci even) or grp + 2 (ci odd)
      if (GATED) {
        ptx::mbar_wait(BAR(BAR_D1_FULL0 + b), (d1ph >> (ci & 1)) & 1u); d1ph ^= 1u << (ci & 1);
        ptx::tc_fence_after();
        epi_pack64<ACT, FMT>(tl + MP::h1buf(b), tl + MP::h1buf(b), PXR_CB1 + c * 64);
      } else {
        const int buf = T & 1;
        if (ci == 0) ptx::mbar_wait(BAR(BAR_PI_FULL0 + buf), (T >> 1) & 1);          // this tile's item partials landed
        if (WIDE && ci == 0) {
          // gate of this row's pair: softmax over the M modality logits (layers.py:207-211); the unit's user parts
          // (ms.lu, Pu) are in place once the tile has been announced
          const float li[8] = {wl0.x, wl0.y, wl0.z, wl0.w, wl1.x, wl1.y, wl1.z, wl1.w};
          float mx = -INFINITY;
#pragma unroll
          for (int m = 0; m < 6; ++m) { wg[m] = m < p.M ? li[m] + ms.lu[ru][m] : -INFINITY; mx = fmaxf(mx, wg[m]); }
          float sum = 0.f;
#pragma unroll
          for (int m = 0; m < 6; ++m) { wg[m] = m < p.M ? expf(wg[m] - mx) : 0.f; sum += wg[m]; }
          const float inv = 1.f / sum;
#pragma unroll
          for (int m = 0; m < 6; ++m) wg[m] *= inv;
          if (GW_HALF2) {
#pragma unroll
            for (int m = 0; m < 5; ++m) wgh[m] = __float2half2_rn(wg[m + 1]);
          }
        }
        const uint32_t n = (ci & 1) ? h1use1++ : h1use0++;
        if (n > 0) { ptx::mbar_wait(BAR(BAR_H1_EMPTY0 + b), (n - 1) & 1); ptx::tc_fence_after(); }   // layer 2 consumed the buffer
        if (WIDE) {
          const uint32_t G = (uint32_t)T * 8u + c, slot = G % MP::Q_STAGES;
          ptx::mbar_wait(BAR(BAR_Q_FULL0 + slot), (G / MP::Q_STAGES) & 1);        // this chunk of the tile's item partials landed
          if (GW_HALF2)
            gatedw_h1_chunk_h2<ACT, FMT>(sm + MP::OFF_PI + slot * MP::Q_STAGE_MAX + rj * q_item_bytes(p.M - 1), p.M - 1, wg[0], wgh,
                                         sm + MP::OFF_PU + ru * MP::PU_STRIDE + c * 256, tl + MP::h1buf(b));
          else
          gatedw_h1_chunk<ACT, FMT>(sm + MP::OFF_PI + slot * MP::Q_STAGE_MAX + rj * q_item_bytes(p.M - 1), p.M - 1, wg,
                                    sm + MP::OFF_PU + ru * MP::PU_STRIDE + c * 256, tl + MP::h1buf(b));
        } else
        concat_h1_chunk<ACT, FMT>(SPREAD ? reinterpret_cast<const uint8_t*>(pi_g) + c * 128
                                         : sm + MP::OFF_PI + buf * MP::PI_BUF + rj * MP::PI_STRIDE + c * 128,
                                  sm + MP::OFF_PU + ru * MP::PU_STRIDE + c * 256, tl + MP::h1buf(b));
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive_cluster(BAR(BAR_H1_FULL0 + b), 0);
        if (WIDE) ptx::mbar_arrive_local(BAR(BAR_Q_EMPTY0 + ((uint32_t)T * 8u + c) % MP::Q_STAGES));   // this warp is done with the staged chunk
        if (!GATED && ci == 3) ptx::mbar_arrive_local(BAR(BAR_PI_EMPTY0 + (T & 1)));   // done with this tile's Pi (and Pu)
      }
    };

    for (int w = pair; w < p.n_units; w += n_pairs) {
      const Unit un = decode_unit<TI, SPREAD>(p, w);
      const int64_t ubase = ((int64_t)un.g * 2 + rank) * TU;
      for (int t = 0; t < un.ntiles; ++t, ++T) {
        // One rolled loop over this group's four layer-1 chunks (the body must stay resident in the instruction
        // cache: fully unrolling it costs ~6 % throughput).  The deferred layer-2 / layer-3 epilogues of the
        // PREVIOUS tile are slotted in where their inputs become available:
        //   gated : E2(T-1), L1(0), E3(T-1), L1(1), L1(2), L1(3)
        //   concat: L1(0), E2(T-1), L1(1), E3(T-1), L1(2), L1(3)   (L1(0) refills a free buffer while layer 2 drains)
        if (!GATED && SPREAD) {
          const int64_t lo = un.row_lo + ((ubase + ru) / p.n_real) * p.sub_rows;
          const int64_t row = lo + (int64_t)t * TI + rj;
          const bool okr = row < min(lo + (int64_t)p.sub_rows, un.row_hi);
          pi_g = p.item_pi + (okr ? row : un.row_lo) * H1;          // rows beyond the sub-range compute on a valid item and are dropped
          if (t + 1 < un.ntiles && row + TI < un.row_hi) {
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) asm volatile("prefetch.global.L2 [%0];" ::"l"(pi_g + (size_t)TI * H1 + (2 * ci + grp) * 64));
          }
        }
        if (WIDE) {
          const int64_t row = un.row_lo + (int64_t)t * TI + rj;
          const int64_t rr = row < un.row_hi ? row : un.row_lo;        // padding rows compute on a valid item's logits and are dropped
          wl0 = __ldg(reinterpret_cast<const float4*>(p.item_logit + rr * 8));
          wl1 = __ldg(reinterpret_cast<const float4*>(p.item_logit + rr * 8 + 4));
        }
        if (GATED && have_prev) do_e2(T - 1);
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          do_l1(ci);
          if (have_prev) {
            if (!GATED && ci == 0) do_e2(T - 1);
            if (grp == 0 && ci == (GATED ? 0 : 1)) prev_e3();
          }
        }
        if (ATT) {
          prev = un; prev_t = t; prev_ubase = ubase; have_prev = true;
        } else if (SPREAD) {
          // this thread's slot sweeps the sub-range (ubase + ru) / n_real of the shard (recomputed per tile: the division is
          // off the critical path and the epilogue keeps no per-unit state for it)
          const int64_t v = ubase + ru;
          const int64_t lo = un.row_lo + (v / p.n_real) * p.sub_rows;
          const int64_t row = lo + (int64_t)t * TI + rj;
          prev_row = (row < min(lo + (int64_t)p.sub_rows, un.row_hi) && v < p.n_users) ? row : -1;
          prev_flags = (t == 0 ? 1 : 0) | (t == un.ntiles - 1 ? 2 : 0);
          have_prev = true;
        } else {
          const int64_t row = un.row_lo + (int64_t)t * TI + rj;
          prev_row = (row < un.row_hi && (ubase + ru) < p.n_users) ? row : -1;
          prev_flags = (t == 0 ? 1 : 0) | (t == un.ntiles - 1 ? 2 : 0);
          have_prev = true;
        }
      }
    }
    if (have_prev) {
      do_e2(T - 1);
      if (grp == 0) prev_e3();
    }
  }

  // ------------------------------------------------------------------ teardown
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 4) ptx::tmem_dealloc_2cta(tmem, 512);
}

#if PXR_TC_TU == 0
// ---------------------------------------------------------------------------------------------
// one-off preparation kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  if (fmt == FMT_BF16) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  return *reinterpret_cast<uint16_t*>(&h);
}

// Builds the two per-CTA-rank operand images: 16-bit, K-major, 128-byte swizzle (16-byte chunk index XOR
// row-in-group), laid out exactly as the kernel's shared memory.  w1 / k1: layer-1 weight (row stride k1) or NULL.
// (n1, n2, n3) = the model's hidden dims: rows / columns beyond them are zero, i.e. a smaller MLP runs zero-padded to the
// kernel's [512, 256, 128] (a padded unit has bias 0 and weight 0: relu(0) = 0 feeds nothing forward -- exact).
__global__ void build_wimg_kernel(const float* __restrict__ w1, int k1, const float* __restrict__ k1_scale,
                                  const float* __restrict__ w2, const float* __restrict__ w3, uint8_t* __restrict__ img,
                                  int gated, int fmt, int n1, int n2, int n3) {
  const uint32_t off_w2 = gated ? 32768u : 0u, off_w3 = off_w2 + 131072u, wimg = off_w3 + 32768u;
  const int total = (int)wimg;                                       // 2 ranks x wimg/2 elements
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int rank = e / (int)(wimg / 2);
    const uint32_t off = (uint32_t)(e % (int)(wimg / 2)) * 2;        // byte offset inside the image
    uint32_t rel, rows_per_blk;
    int which;
    if (off < off_w2) { which = 1; rel = off; rows_per_blk = 32; }
    else if (off < off_w3) { which = 2; rel = off - off_w2; rows_per_blk = 128; }
    else { which = 3; rel = off - off_w3; rows_per_blk = 64; }
    const uint32_t blk_bytes = rows_per_blk * 128;
    const uint32_t blk = rel / blk_bytes, inb = rel % blk_bytes;
    const uint32_t nl = inb / 128, inrow = inb % 128;
    const uint32_t chunk = (inrow >> 4) ^ (nl & 7);                  // un-swizzle: stored chunk -> logical chunk
    const uint32_t kk = chunk * 8 + ((inrow & 15) >> 1);             // k inside the 64-wide block
    float v = 0.f;
    if (which == 1) {                                                // N-chunk blk, rows [32 rank, +32)
      const int n = blk * 64 + rank * 32 + nl;
      if (n < n1) v = w1[(size_t)n * k1 + kk] * (k1_scale ? k1_scale[kk] : 1.f);
    } else if (which == 2) {                                         // K-block blk, rows [128 rank, +128)
      const int n = rank * 128 + nl, k = blk * 64 + kk;
      if (n < n2 && k < n1) v = w2[(size_t)n * n1 + k];
    } else {
      const int n = rank * 64 + nl, k = blk * 64 + kk;
      if (n < n3 && k < n2) v = w3[(size_t)n * n2 + k];
    }
    reinterpret_cast<uint16_t*>(img + (size_t)rank * wimg)[off / 2] = to16(v, fmt);
  }
}

// concat / wide gated: W1^T [k][n1] -> [k][512], columns past n1 zero
__global__ void pad_w1t_kernel(const float* __restrict__ wt, int k1, int n1, float* __restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)k1 * H1) return;
  const int n = (int)(e % H1); const size_t k = e / H1;
  out[e] = n < n1 ? wt[k * n1 + n] : 0.f;
}

// gated: item part of the gate logits: Wg[:, D:] . concat(item-side vectors) + bg   (layers.py:207 split)
__global__ void item_logit_kernel(const float* __restrict__ feats, const float* __restrict__ gate_w,
                                  const float* __restrict__ gate_b, int M, int Dm, int64_t n_rows, float* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int FD = (M - 1) * Dm;
  const float* x = feats + row * FD;
  for (int m = 0; m < 8; ++m) {
    float acc = 0.f;
    if (m < M) {
      const float* wr = gate_w + (size_t)m * M * Dm + Dm;
      for (int k = lane; k < FD; k += 32) acc += wr[k] * x[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      acc += gate_b[m];
    }
    if (lane == 0) out[row * 8 + m] = acc;
  }
}

// concat: item partial of layer 1, Pi[row] = W1[:, D:] . concat(item-side vectors) + b1   -> 16 bit.
// 32 rows per block; W1^T (k-major, [M*D][512]) rows D.. are the item part.
// (wide gated: the same kernel with FD = D over the n_rows * (M - 1) modality vectors and all of W1^T: Q_m = W1 f_m + b1)
__global__ void __launch_bounds__(PXR_SIMT_THREADS) item_pi_kernel(const float* __restrict__ feats, const float* __restrict__ w1t,
                                                                    const float* __restrict__ b1, int FD, int64_t n_rows,
                                                                    uint16_t* __restrict__ out, int fmt, int qt_nm) {
  extern __shared__ __align__(16) float smem_pi[];
  float* in = smem_pi;                       // [32][FD]
  float* res = smem_pi + 32 * FD;            // [32][512]
  const int64_t row0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * FD; i += PXR_SIMT_THREADS) {
    const int64_t row = row0 + i / FD;
    in[i] = row < n_rows ? feats[row * FD + i % FD] : 0.f;
  }
  __syncthreads();
  linear_rows<32>(in, FD, FD, w1t, b1, H1, res, H1, -1);
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * H1; i += PXR_SIMT_THREADS) {
    const int64_t row = row0 + i / H1;
    if (row >= n_rows) continue;
    const int n = i % H1;
    if (qt_nm > 0) {            // wide gated: vector `row` = (item, modality), stored chunk-major (Params::item_q)
      const int64_t item = row / qt_nm; const int m = (int)(row % qt_nm);
      const size_t off = ((size_t)(item >> 4) * 8 + (n >> 6)) * q_stage_bytes(qt_nm) + (size_t)(item & 15) * q_item_bytes(qt_nm) + m * 128 + (n & 63) * 2;
      *reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(out) + off) = to16(res[i], fmt);
    } else {
      out[row * H1 + n] = to16(res[i], fmt);
    }
  }
}


// attention: per-item record (once per catalogue shard): everything of the attention layer that involves item tokens
// only, stored as the 16-bit B fragments of the front end's register MMAs in lane order (layout: REC_* above,
// definitions above attn_item_step).  One item per 64-thread block, thread = d.
//   wc : [64][64] fp32 centred out_proj weight, Wc[d][j] = W_o[d][j] - mean_d' W_o[d'][j]
__device__ __forceinline__ uint32_t pack16(float lo, float hi, int fmt) {
  return (uint32_t)to16(lo, fmt) | ((uint32_t)to16(hi, fmt) << 16);
}
__device__ __forceinline__ float from16(uint16_t v, int fmt) {
  if (fmt == FMT_BF16) return __uint_as_float((uint32_t)v << 16);
  return __half2float(*reinterpret_cast<const __half*>(&v));
}
__device__ __forceinline__ float lo_part(float x, int fmt) { return x - from16(to16(x, fmt), fmt); }   // x = hi + lo, both 16-bit

__global__ void __launch_bounds__(64) item_attn_kernel(const float* __restrict__ feats, const float* __restrict__ in_wt,
                                                       const float* __restrict__ in_b, const float* __restrict__ wc,
                                                       const float* __restrict__ out_b, int M, int64_t n_rows, int fmt,
                                                       uint4* __restrict__ rec) {
  __shared__ float x[ATT_TOKENS][D], q[ATT_TOKENS][D], k[ATT_TOKENS][D], v[ATT_TOKENS][D];
  __shared__ float Ssc[ATT_TOKENS][NH][ATT_TOKENS], P[ATT_TOKENS][NH][ATT_TOKENS], Lse[ATT_TOKENS][NH];
  __shared__ float vbar[ATT_TOKENS][D];                       // [a][16 h + e] = sum_b softmax_b(s_ab,h) v_b,h[e]
  __shared__ float Nc[ATT_TOKENS][NH][D], Uc[ATT_TOKENS][NH][D], xc[ATT_TOKENS][D];
  __shared__ float red[2];
  const int d = threadIdx.x, nt = M - 1, lane = d & 31;
  const int64_t row = blockIdx.x;
  if (row >= n_rows) return;
  for (int b = 0; b < ATT_TOKENS; ++b) x[b][d] = b < nt ? feats[(row * nt + b) * D + d] : 0.f;
  __syncthreads();
  {
    float aq[ATT_TOKENS], ak[ATT_TOKENS], av[ATT_TOKENS];
    for (int b = 0; b < ATT_TOKENS; ++b) { aq[b] = in_b[d]; ak[b] = in_b[D + d]; av[b] = in_b[2 * D + d]; }
    for (int kk = 0; kk < D; ++kk) {
      const float wq = in_wt[kk * 3 * D + d], wk = in_wt[kk * 3 * D + D + d], wv = in_wt[kk * 3 * D + 2 * D + d];
#pragma unroll
      for (int b = 0; b < ATT_TOKENS; ++b) {
        const float xv = x[b][kk];
        aq[b] = fmaf(wq, xv, aq[b]); ak[b] = fmaf(wk, xv, ak[b]); av[b] = fmaf(wv, xv, av[b]);
      }
    }
    for (int b = 0; b < ATT_TOKENS; ++b) { q[b][d] = aq[b] * 0.25f; k[b][d] = ak[b]; v[b][d] = av[b]; }   // q / sqrt(dh)
  }
  __syncthreads();
  for (int i = d; i < ATT_TOKENS * ATT_TOKENS * NH; i += 64) {     // item-item scores (already scaled through q)
    const int a = i / (ATT_TOKENS * NH), b = (i / NH) % ATT_TOKENS, h = i % NH;
    float acc = 0.f;
    for (int e = 0; e < DH; ++e) acc = fmaf(q[a][h * DH + e], k[b][h * DH + e], acc);
    Ssc[a][h][b] = acc;
  }
  __syncthreads();
  if (d < ATT_TOKENS * NH) {
    const int a = d >> 2, h = d & 3;
    float m = -INFINITY;
    for (int b = 0; b < nt; ++b) m = fmaxf(m, Ssc[a][h][b]);
    float sum = 0.f;
    for (int b = 0; b < nt; ++b) { const float e = expf(Ssc[a][h][b] - m); P[a][h][b] = e; sum += e; }
    for (int b = 0; b < ATT_TOKENS; ++b) P[a][h][b] = b < nt ? P[a][h][b] / sum : 0.f;
    Lse[a][h] = (nt > 0 && a < nt) ? m + logf(sum) : 0.f;
  }
  __syncthreads();
  {
    const int h = d >> 4;
    for (int a = 0; a < ATT_TOKENS; ++a) {
      float acc = 0.f;
      for (int b = 0; b < nt; ++b) acc = fmaf(P[a][h][b], v[b][d], acc);
      vbar[a][d] = acc;
    }
  }
  // xc_a = (x_a + b_o) centred over d
  for (int a = 0; a < ATT_TOKENS; ++a) {
    const float val = x[a][d] + out_b[d];
    float part = val;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __syncthreads();
    if (lane == 0) red[d >> 5] = part;
    __syncthreads();
    xc[a][d] = a < nt ? val - (red[0] + red[1]) * (1.f / D) : 0.f;
  }
  __syncthreads();
  for (int a = 0; a < ATT_TOKENS; ++a)
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float n_ = 0.f, u_ = 0.f;
      for (int e = 0; e < DH; ++e) {
        const float w = wc[d * D + h * DH + e];
        n_ = fmaf(w, vbar[a][h * DH + e], n_);
        u_ = fmaf(w, v[a][h * DH + e], u_);
      }
      Nc[a][h][d] = a < nt ? n_ : 0.f;
      Uc[a][h][d] = a < nt ? u_ : 0.f;
    }
  __syncthreads();
  uint4* out = rec + row * ATT_REC_U4;
  for (int o = d; o < ATT_REC_U4; o += 64) {
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (o < REC_L) {                                              // scores: [h][lane] = (q b0, q b1, k b0, k b1), n = token g
      const int h = o >> 5, ln = o & 31, g = ln >> 2, t = ln & 3, c = h * DH + 2 * t;
      if (g < nt) {
        w.x = pack16(q[g][c], q[g][c + 1], fmt); w.y = pack16(q[g][c + 8], q[g][c + 9], fmt);
        w.z = pack16(k[g][c], k[g][c + 1], fmt); w.w = pack16(k[g][c + 8], k[g][c + 9], fmt);
      }
    } else if (o < REC_T) {                                       // L: [t][2 tokens] x 4 heads, fp32
      const int i = o - REC_L, a = i;                             // token a = 2 t + tok = i
      if (a < ATT_TOKENS) w = make_uint4(__float_as_uint(Lse[a][0]), __float_as_uint(Lse[a][1]), __float_as_uint(Lse[a][2]), __float_as_uint(Lse[a][3]));
    } else if (o < REC_U) {                                       // token tails: [a][np][lane] = (b0, b1 of n-tile 2 np; of 2 np + 1)
      const int i = o - REC_T, a = i >> 7, np = (i >> 5) & 3, ln = i & 31, g = ln >> 2, t = ln & 3;
      uint32_t r[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 8 * (2 * np + e) + g;
        if (t < 2) {                                              // K rows 2t, 2t+1: Nc_h hi; rows 8+2t, 9+2t: Nc_h lo
          r[2 * e] = pack16(Nc[a][2 * t][n], Nc[a][2 * t + 1][n], fmt);
          r[2 * e + 1] = pack16(lo_part(Nc[a][2 * t][n], fmt), lo_part(Nc[a][2 * t + 1][n], fmt), fmt);
        } else if (t == 2) {                                      // K rows 4, 5: xc hi, xc lo
          r[2 * e] = pack16(xc[a][n], lo_part(xc[a][n], fmt), fmt);
        }
      }
      w = make_uint4(r[0], r[1], r[2], r[3]);
    } else {                                                      // user-row item values: [kk][np][lane], K row 8 (h & 1) + token
      const int i = o - REC_U, kk = i >> 7, np = (i >> 5) & 3, ln = i & 31, g = ln >> 2, t = ln & 3;
      uint32_t r[4] = {0u, 0u, 0u, 0u};
      const int b0 = 2 * t, b1 = 2 * t + 1;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = 8 * (2 * np + e) + g;
        const float u00 = b0 < ATT_TOKENS ? Uc[b0][2 * kk][n] : 0.f, u01 = b1 < ATT_TOKENS ? Uc[b1][2 * kk][n] : 0.f;
        const float u10 = b0 < ATT_TOKENS ? Uc[b0][2 * kk + 1][n] : 0.f, u11 = b1 < ATT_TOKENS ? Uc[b1][2 * kk + 1][n] : 0.f;
        r[2 * e] = pack16(u00, u01, fmt);                          // head 2 kk:     K rows 2t, 2t + 1
        r[2 * e + 1] = pack16(u10, u11, fmt);                      // head 2 kk + 1: K rows 8 + 2t, 9 + 2t
      }
      w = make_uint4(r[0], r[1], r[2], r[3]);
    }
    out[o] = w;
  }
}

// attention: centred out_proj weight Wc = (I - 11^T / D) W_o as fp32 [d][j], and as the B fragments of Y = A Wc^T
// (B[k = j][n = d]) in lane order: [4 k-steps][4 n-tile pairs][32 lanes] x 16 B
__global__ void attn_wc_kernel(const float* __restrict__ out_w, float* __restrict__ wc, uint4* __restrict__ frag, int fmt) {
  __shared__ float w[D][D + 1];
  const int tid = threadIdx.x;                                   // 256 threads
  for (int i = tid; i < D * D; i += blockDim.x) w[i / D][i % D] = out_w[i];
  __syncthreads();
  if (tid < D) {                                                  // column j = tid: subtract its mean over d
    float m = 0.f;
    for (int dd = 0; dd < D; ++dd) m += w[dd][tid];
    m *= (1.f / D);
    for (int dd = 0; dd < D; ++dd) w[dd][tid] -= m;
  }
  __syncthreads();
  for (int i = tid; i < D * D; i += blockDim.x) wc[i] = w[i / D][i % D];
  for (int o = tid; o < 4 * 4 * 32; o += blockDim.x) {
    const int kk = o >> 7, np = (o >> 5) & 3, ln = o & 31, g = ln >> 2, t = ln & 3, j = 16 * kk + 2 * t;
    uint32_t r[4];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = 8 * (2 * np + e) + g;
      r[2 * e] = pack16(w[n][j], w[n][j + 1], fmt);
      r[2 * e + 1] = pack16(w[n][j + 8], w[n][j + 9], fmt);
    }
    frag[o] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

// attention: b1' = b1 + W1 ln_b (LayerNorm bias folded into layer 1); s1[k] = ln_w[k] / M for the W1 image
__global__ void attn_fold_ln_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ ln_w,
                                    const float* __restrict__ ln_b, int M, int n1, float* __restrict__ b1_out, float* __restrict__ s1_out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < H1) {
    float acc = 0.f;
    if (n < n1) {
      acc = b1[n];
      for (int k = 0; k < D; ++k) acc = fmaf(w1[(size_t)n * D + k], ln_b[k], acc);
    }
    b1_out[n] = acc;
  }
  if (n < D) s1_out[n] = ln_w[n] / (float)M;
}

// top_k > 64: the merged 64-slot list of page `page` goes into the user's candidate row (64 * n_pages slots), and the key
// of its last entry becomes the bound of the next page (0 = the list was not full: nothing is left for the next page)
__global__ void page_commit_kernel(const float* __restrict__ page_s, const int32_t* __restrict__ page_i, int64_t n_users, int page,
                                   int n_pages, float* __restrict__ cand_s, int32_t* __restrict__ cand_i,
                                   unsigned long long* __restrict__ upper) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_users * KCAP) return;
  const int64_t u = i / KCAP; const int j = (int)(i % KCAP);
  const float sc = page_s[i]; const int32_t gi = page_i[i];
  const int64_t o = (u * n_pages + page) * KCAP + j;
  cand_s[o] = sc; cand_i[o] = gi;
  if (j == KCAP - 1) upper[u] = gi >= 0 ? (((unsigned long long)pxr_ord(sc) << 32) | (unsigned long long)(IDX_MASK - (uint32_t)gi)) : 0ull;
}

struct FastWeights {       // lives in h->fast_w; attention: followed by one xc_0 scratch (ATT_XC0_U4 x 16 B) per SM
  uint8_t wimg[2 * Map<F_GATED>::WIMG];
  float bias[H1 + H2 + H3 + H3 + 4];
  float s1[D];             // attention: ln_w / M, the per-input scale folded into the layer-1 image
  float wc[D * D];         // attention: centred out_proj weight (fp32), for the per-item records
  uint4 wo_frag[4 * 4 * 32];   // attention: its 16-bit B fragments for the front end's register MMAs
};

#endif  // PXR_TC_TU == 0

template <int FUS, int FMT, bool TK2, int ACT, int MODE>
static int launch_fused_tk(pxr_handle* h, const Params& p, int n_pairs, cudaStream_t st) {
  auto kern = score_fused_kernel<FUS, FMT, TK2, ACT, MODE>;
  static_assert(!(MODE == M_PAGED && TK2), "paged passes use one top-K warp");
  const int slot = ((ACT * 4 + FUS) * 2 + FMT) * 4 + (MODE == M_PLAIN ? (TK2 ? 1 : 0) : 1 + MODE);   // < 160: three words
  static_assert(sizeof(h->tc_attr_fused) >= 3 * sizeof(uint64_t), "one bit per instantiation");
  if (!(h->tc_attr_fused[slot >> 6] & (1ull << (slot & 63)))) {
    PXR_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Map<FUS>::SMEM));
    h->tc_attr_fused[slot >> 6] |= (1ull << (slot & 63));
  }
  pxr_prof_begin(h, st);
  kern<<<2 * n_pairs, n_threads<FUS>(), Map<FUS>::SMEM, st>>>(p);
  pxr_prof_end(h, st);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

// The number of list updates per user grows like K (1 + ln(n / K)) with the n item rows of a unit, most of them at its
// start: below ~8 K rows per unit one inserting warp is the bottleneck (measured: 10 K items x 1 024 users 5.1 -> 3.5 ms with
// two); for long units the second warp only costs the front end issue slots (-2.8 % on the gated headline config).
template <int FUS, int FMT, int ACT>
static int launch_fused(pxr_handle* h, const Params& p, int n_pairs, cudaStream_t st) {
  static int thr = -1;                    // PXR_TK2_ROWS: rows per unit below which the second top-K warp is used (experiments)
  if (thr < 0) { const char* e = getenv("PXR_TK2_ROWS"); thr = e ? atoi(e) : 8192; }
  if (p.upper) return launch_fused_tk<FUS, FMT, false, ACT, M_PAGED>(h, p, n_pairs, st);
  if (p.sub_rows > 0) {                   // small batches: short sub-ranges per slot, two top-K warps
    if constexpr (FUS == F_GATED || FUS == F_CONCAT) return launch_fused_tk<FUS, FMT, true, ACT, M_SPREAD>(h, p, n_pairs, st);
    else PXR_FAIL(h, PXR_ERR_INVALID, "small-batch mode exists for gated (embedding_dim 64) and concat fusion only");
  }
  return p.rows_per_split < thr ? launch_fused_tk<FUS, FMT, true, ACT, M_PLAIN>(h, p, n_pairs, st)
                                : launch_fused_tk<FUS, FMT, false, ACT, M_PLAIN>(h, p, n_pairs, st);
}

// every (front end, operand format) of one activation; fp16 operands (an opt-in of the ReLU kernels) are not built for the
// other activations (build time: 20 instantiations per activation object otherwise)
template <int ACT>
static int launch_fused_act(pxr_handle* h, const Params& p, int fusion, int fmt, int n_pairs, cudaStream_t st) {
  if constexpr (ACT != PXR_ACT_RELU || PXR_TC_TU == 0) {     // TU 0: the bf16 half of the ReLU kernels (fp16: TU 5)
    if (fmt != FMT_BF16) PXR_FAIL(h, PXR_ERR_INVALID, "fp16 operands are built for relu models only");
    if (fusion == PXR_FUSION_GATED && p.item_q) return launch_fused<F_GATEDW, FMT_BF16, ACT>(h, p, n_pairs, st);   // embedding_dim != 64
    if (fusion == PXR_FUSION_GATED) return launch_fused<F_GATED, FMT_BF16, ACT>(h, p, n_pairs, st);
    if (fusion == PXR_FUSION_ATTENTION) return launch_fused<F_ATTN, FMT_BF16, ACT>(h, p, n_pairs, st);
    return launch_fused<F_CONCAT, FMT_BF16, ACT>(h, p, n_pairs, st);
  }
  else {                                                      // TU 5: ReLU, fp16 operands
    if (fusion == PXR_FUSION_GATED && p.item_q) return launch_fused<F_GATEDW, FMT_FP16, ACT>(h, p, n_pairs, st);
    if (fusion == PXR_FUSION_GATED) return launch_fused<F_GATED, FMT_FP16, ACT>(h, p, n_pairs, st);
    if (fusion == PXR_FUSION_ATTENTION) return launch_fused<F_ATTN, FMT_FP16, ACT>(h, p, n_pairs, st);
    return launch_fused<F_CONCAT, FMT_FP16, ACT>(h, p, n_pairs, st);
  }
}

}  // namespace tc

// launchers of the other activations' objects (same signature in every TU)
#define PXR_TC_CAT2(a, b) a##b
#define PXR_TC_CAT(a, b) PXR_TC_CAT2(a, b)
#if PXR_TC_TU == 5
int pxr_tc_launch_relu_fp16(pxr_handle* h, const tc::Params& p, int fusion, int fmt, int n_pairs, cudaStream_t st) {
  return tc::launch_fused_act<PXR_ACT_RELU>(h, p, fusion, fmt, n_pairs, st);
}
#elif PXR_TC_TU != 0
int PXR_TC_CAT(pxr_tc_launch_act, PXR_TC_TU)(pxr_handle* h, const tc::Params& p, int fusion, int fmt, int n_pairs, cudaStream_t st) {
  return tc::launch_fused_act<PXR_TC_TU>(h, p, fusion, fmt, n_pairs, st);
}
#else
int pxr_tc_launch_relu_fp16(pxr_handle* h, const tc::Params& p, int fusion, int fmt, int n_pairs, cudaStream_t st);
int pxr_tc_launch_act1(pxr_handle* h, const tc::Params& p, int fusion, int fmt, int n_pairs, cudaStream_t st);   // gelu
int pxr_tc_launch_act2(pxr_handle* h, const tc::Params& p, int fusion, int fmt, int n_pairs, cudaStream_t st);   // tanh
int pxr_tc_launch_act3(pxr_handle* h, const tc::Params& p, int fusion, int fmt, int n_pairs, cudaStream_t st);   // leaky_relu
int pxr_tc_launch_act4(pxr_handle* h, const tc::Params& p, int fusion, int fmt, int n_pairs, cudaStream_t st);   // silu
static_assert(PXR_ACT_GELU == 1 && PXR_ACT_TANH == 2 && PXR_ACT_LEAKY_RELU == 3 && PXR_ACT_SILU == 4, "object <-> activation mapping");

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// NULL when the fused tcgen05 kernel covers this model configuration, else why it does not
const char* pxr_tc_unsupported_reason(const pxr_handle* h) {
  const pxr_config& c = h->cfg;
  if (h->n_sm < 2) return "the device has fewer than 2 SMs (the kernel runs on CTA pairs)";
  if (c.n_hidden != 3) return "the prediction MLP does not have three hidden layers";
  if (c.hidden[0] > tc::H1 || c.hidden[1] > tc::H2 || c.hidden[2] > tc::H3)
    return "fusion_hidden_dims exceeds [512, 256, 128] (the three weight matrices are resident in the CTA pair's shared memory)";
  // smaller hidden layers run zero-padded to [512, 256, 128] (concat / wide gated: the layer-1 partials are padded to 512 columns)
  if (c.activation != PXR_ACT_RELU && c.precision == PXR_PRECISION_FP16) return "fp16 operands with a fusion_activation other than relu (only the relu kernels are built for fp16)";
  if (h->M < 4 || h->M > 6) return "fewer than 4 modalities";
  if (c.fusion == PXR_FUSION_ATTENTION && c.num_heads != tc::NH) return "attention fusion with num_attention_heads != 4";
  if (c.fusion == PXR_FUSION_ATTENTION && c.embedding_dim != tc::D) return "attention fusion with embedding_dim != 64 (its front end is built for 4 heads of 16)";
  // concat, and gated at embedding_dim != 64 (F_GATEDW): layer 1 is applied as per-user / per-item partials, so the fused
  // kernel does not depend on embedding_dim (item side: 3xTF32 GEMMs of items_tc.cu, which need single-layer projections
  // and 16-byte aligned rows)
  if (c.fusion != PXR_FUSION_ATTENTION && c.embedding_dim != tc::D &&
      !(c.embedding_dim % 16 == 0 && c.embedding_dim <= 512 && c.projection_hidden == 0 && c.vision_dim % 4 == 0 &&
        c.language_dim % 4 == 0 && c.num_numerical <= 32))
    return "concat / gated fusion with embedding_dim != 64 needs embedding_dim % 16 == 0, single-layer projections and feature dims % 4 == 0";
  return nullptr;
}

// gated fusion at embedding_dim != 64: the F_GATEDW front end (gate-weighted layer-1 partials on the concat pipeline)
bool pxr_tc_gated_wide(const pxr_handle* h) { return h->cfg.fusion == PXR_FUSION_GATED && h->cfg.embedding_dim != tc::D; }

// concat / wide gated: layer 1 is applied as partials of exactly 512 columns, so a smaller first hidden layer runs on a
// zero-padded copy of W1^T ([k][512], the rows of h->mlp[0].wt padded with zeros) and the zero-padded b1 of the bias block
static bool tc_l1_partials(const pxr_handle* h) { return h->cfg.fusion == PXR_FUSION_CONCAT || pxr_tc_gated_wide(h); }
static size_t tc_w1t_pad_bytes(const pxr_handle* h) {
  if (!tc_l1_partials(h)) return 0;
  const int k1 = (h->cfg.fusion == PXR_FUSION_CONCAT ? h->M : 1) * h->cfg.embedding_dim;
  return pxr_align_up((size_t)k1 * tc::H1 * sizeof(float), 256);
}

bool pxr_tc_supported(const pxr_handle* h) { return pxr_tc_unsupported_reason(h) == nullptr; }

bool pxr_tc_can_run(const pxr_handle* h, int32_t k) {
  return h->fast_ok && k <= PXR_TC_MAX_K && h->n_rows > 0 && h->item_base + h->n_rows < (int64_t)tc::IDX_MASK;
}

size_t pxr_tc_weight_bytes(const pxr_handle* h) {
  return pxr_align_up(sizeof(tc::FastWeights), 256) + tc_w1t_pad_bytes(h) +
         (h->cfg.fusion == PXR_FUSION_ATTENTION ? (size_t)h->n_sm * tc::ATT_XC0_U4 * sizeof(uint4) : 0);
}

// [k][512] zero-padded W1^T (concat / wide gated; lives behind FastWeights in h->fast_w) and the zero-padded b1
const float* pxr_tc_w1t_padded(const pxr_handle* h) {
  return reinterpret_cast<const float*>((const char*)h->fast_w + pxr_align_up(sizeof(tc::FastWeights), 256));
}
const float* pxr_tc_b1_padded(const pxr_handle* h) { return reinterpret_cast<const tc::FastWeights*>(h->fast_w)->bias; }

static int tc_fmt(const pxr_handle* h) { return h->cfg.precision == PXR_PRECISION_FP16 ? tc::FMT_FP16 : tc::FMT_BF16; }

int pxr_tc_prepare_weights(pxr_handle* h, cudaStream_t st) {
  if (!h->fast_w) PXR_CUDA(h, cudaMalloc(&h->fast_w, pxr_tc_weight_bytes(h)));
  tc::FastWeights* fw = reinterpret_cast<tc::FastWeights*>(h->fast_w);
  const bool gated = h->cfg.fusion != PXR_FUSION_CONCAT && !pxr_tc_gated_wide(h);    // layer 1 on the tensor pipe
  const bool attn = h->cfg.fusion == PXR_FUSION_ATTENTION;
  float* b = fw->bias;
  const int n1 = h->cfg.hidden[0], n2 = h->cfg.hidden[1], n3 = h->cfg.hidden[2];     // <= 512 / 256 / 128: zero-padded to the kernel's shape
  PXR_CUDA(h, cudaMemsetAsync(b, 0, sizeof(fw->bias), st));
  if (attn) {     // LayerNorm affine folded into layer 1 (see attn_item_step); centred out_proj weight and its MMA fragments
    tc::attn_fold_ln_kernel<<<(tc::H1 + 127) / 128, 128, 0, st>>>(h->mlp[0].w, h->mlp[0].b, h->ln_w, h->ln_b, h->M, n1, b, fw->s1);
    tc::attn_wc_kernel<<<1, 256, 0, st>>>(h->attn_out.w, fw->wc, fw->wo_frag, tc_fmt(h));
    h->launches += 2;
  } else {
    PXR_CUDA(h, cudaMemcpyAsync(b, h->mlp[0].b, sizeof(float) * n1, cudaMemcpyDeviceToDevice, st));
  }
  if (tc_l1_partials(h)) {
    const int k1 = h->mlp[0].k;
    float* wp = const_cast<float*>(pxr_tc_w1t_padded(h));
    tc::pad_w1t_kernel<<<(unsigned)(((size_t)k1 * tc::H1 + 255) / 256), 256, 0, st>>>(h->mlp[0].wt, k1, n1, wp);
    h->launches++;
  }
  tc::build_wimg_kernel<<<296, 256, 0, st>>>(gated ? h->mlp[0].w : nullptr, h->mlp[0].k, attn ? fw->s1 : nullptr, h->mlp[1].w,
                                             h->mlp[2].w, fw->wimg, gated ? 1 : 0, tc_fmt(h), n1, n2, n3);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1, h->mlp[1].b, sizeof(float) * n2, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2, h->mlp[2].b, sizeof(float) * n3, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2 + tc::H3, h->out.w, sizeof(float) * n3, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2 + 2 * tc::H3, h->out.b, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (attn) {     // the attention kernel reads these through the kernel-parameter constant bank: keep a host copy
    PXR_CUDA(h, cudaMemcpyAsync(h->tc_bias_host, b, sizeof(float) * (tc::H1 + tc::H2 + 2 * tc::H3 + 1), cudaMemcpyDeviceToHost, st));
    PXR_CUDA(h, cudaStreamSynchronize(st));
  }
  return PXR_OK;
}

size_t pxr_tc_item_bytes(const pxr_handle* h, int64_t n_rows) {
  const size_t rows = (size_t)((n_rows + 31) / 32 * 32);
  if (pxr_tc_gated_wide(h))      // gate logits + the M - 1 layer-1 partials per item (16 bit, 512 columns each, chunk-major with padding)
    return pxr_align_up(rows * 8 * sizeof(float), 256) + pxr_align_up(rows / 16 * 8 * (size_t)tc::q_stage_bytes(h->M - 1), 256);
  if (h->cfg.fusion == PXR_FUSION_GATED) return pxr_align_up(rows * 8 * sizeof(float), 256);
  if (h->cfg.fusion == PXR_FUSION_ATTENTION) return pxr_align_up(rows * tc::ATT_REC_U4 * sizeof(uint4), 256);
  return pxr_align_up(rows * tc::H1 * sizeof(uint16_t), 256);
}

int pxr_tc_prepare_items(pxr_handle* h, int64_t n_rows, void* ws, cudaStream_t st) {
  if (n_rows == 0) return PXR_OK;
  if (pxr_tc_gated_wide(h)) {
    // gate logits (as for embedding_dim 64), then the M - 1 layer-1 partials per item: one GEMM over the record viewed as
    // n_rows * (M - 1) modality vectors of embedding_dim, [.. x D] . W1^T + b1 -> 16 bit
    const int Dm = h->cfg.embedding_dim;
    const int64_t rows = (n_rows + 31) / 32 * 32;
    uint16_t* q = (uint16_t*)((char*)ws + pxr_align_up((size_t)rows * 8 * sizeof(float), 256));
    const bool on_tc = h->path == PXR_PATH_TCGEN05 && h->tc_items_img[3] && h->tc_items_img[2];
    // rows of the last tile that do not exist are staged like the others (their scores are dropped): keep them finite
    PXR_CUDA(h, cudaMemsetAsync(q, 0, (size_t)rows / 16 * 8 * tc::q_stage_bytes(h->M - 1), st));
    if (on_tc) {
      int rc = pxr_launch_item_logit_tc(h, n_rows, (float*)ws, st);
      if (rc) return rc;
      return pxr_launch_item_q_tc(h, n_rows, q, tc::GW_HALF2 ? tc::FMT_FP16 : tc_fmt(h), st);
    }
    const int wpb = 8;
    tc::item_logit_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(h->item_feats, h->gate.w, h->gate.b, h->M, Dm,
                                                                                     n_rows, (float*)ws);
    h->launches++;
    const size_t smem = (size_t)32 * (Dm + tc::H1) * sizeof(float);
    if (!(h->tc_attr_set & (1ull << 63))) {
      PXR_CUDA(h, cudaFuncSetAttribute(tc::item_pi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      h->tc_attr_set |= (1ull << 63);
    }
    const int64_t nv = n_rows * (h->M - 1);
    tc::item_pi_kernel<<<(unsigned)((nv + 31) / 32), PXR_SIMT_THREADS, smem, st>>>(h->item_feats, pxr_tc_w1t_padded(h), pxr_tc_b1_padded(h), Dm, nv, q, tc::GW_HALF2 ? tc::FMT_FP16 : tc_fmt(h), h->M - 1);
  } else if (h->cfg.fusion == PXR_FUSION_GATED && h->tc_items_img[3] && h->path == PXR_PATH_TCGEN05) {
    return pxr_launch_item_logit_tc(h, n_rows, (float*)ws, st);                  // 3xTF32 GEMM on the tensor pipe (N = 6 padded to 16)
  } else if (h->cfg.fusion == PXR_FUSION_GATED) {
    const int wpb = 8;
    tc::item_logit_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(h->item_feats, h->gate.w, h->gate.b, h->M,
                                                                                     h->cfg.embedding_dim, n_rows, (float*)ws);
  } else if (h->cfg.fusion == PXR_FUSION_ATTENTION) {
    tc::item_attn_kernel<<<(unsigned)n_rows, 64, 0, st>>>(h->item_feats, h->attn_in.wt, h->attn_in.b,
                                                          reinterpret_cast<tc::FastWeights*>(h->fast_w)->wc, h->attn_out.b, h->M,
                                                          n_rows, tc_fmt(h), (uint4*)ws);
  } else if (h->tc_items_img[2] && h->path == PXR_PATH_TCGEN05) {
    return pxr_launch_item_pi_tc(h, n_rows, (uint16_t*)ws, tc_fmt(h), st);      // 3xTF32 GEMM on the tensor pipe
  } else {
    const int FD = (h->M - 1) * h->cfg.embedding_dim;
    const size_t smem = (size_t)32 * (FD + tc::H1) * sizeof(float);
    if (smem > (size_t)h->max_smem_optin) PXR_FAIL(h, PXR_ERR_INVALID, "concat item partial: embedding_dim %d needs the tensor-pipe item path", h->cfg.embedding_dim);
    if (!(h->tc_attr_set & (1ull << 63))) {
      PXR_CUDA(h, cudaFuncSetAttribute(tc::item_pi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      h->tc_attr_set |= (1ull << 63);
    }
    tc::item_pi_kernel<<<(unsigned)((n_rows + 31) / 32), PXR_SIMT_THREADS, smem, st>>>(h->item_feats, pxr_tc_w1t_padded(h) + (size_t)h->cfg.embedding_dim * tc::H1,
                                                                                       pxr_tc_b1_padded(h), FD, n_rows, (uint16_t*)ws, tc_fmt(h), 0);
  }
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

struct TcPlan { int n_groups, S, rows_per_split, n_units, n_pairs; };

static TcPlan tc_plan(const pxr_handle* h, int64_t n_users) {
  TcPlan pl;
  const int max_pairs = h->n_sm / 2;
  const bool att = h->cfg.fusion == PXR_FUSION_ATTENTION;
  const int TU = att ? tc::tile_users<tc::F_ATTN>() : tc::tile_users<tc::F_GATED>();
  const int TI = att ? tc::tile_items<tc::F_ATTN>() : tc::tile_items<tc::F_GATED>();
  pl.n_groups = (int)((n_users + 2 * TU - 1) / (2 * TU));
  const int64_t max_tiles = (h->n_rows + TI - 1) / TI;
  // split the item range so that the (equal-cost) units fill the CTA pairs evenly: pick the smallest S whose
  // last scheduling round is at least 97 % full (or the best one seen), capped by K4's merge width
  int best_s = 1; double best_eff = 0.0;
  const int s_cap = (int)std::min<int64_t>(std::min<int64_t>(max_tiles, 64), 4096 / 64);
  for (int s = 1; s <= s_cap; ++s) {
    const int64_t units = (int64_t)pl.n_groups * s;
    const int64_t rounds = (units + max_pairs - 1) / max_pairs;
    const double eff = (double)units / (double)(rounds * max_pairs);
    if (eff > best_eff + 1e-9) { best_eff = eff; best_s = s; }
    if (eff >= 0.97) break;
  }
  int64_t rps = (h->n_rows + best_s - 1) / best_s;
  rps = (rps + 15) / 16 * 16;
  pl.rows_per_split = (int)rps;
  pl.S = (int)((h->n_rows + rps - 1) / rps);
  pl.n_units = pl.n_groups * pl.S;
  pl.n_pairs = std::min(max_pairs, pl.n_units);
  return pl;
}

static int tc_pages(int32_t k) { return (k + tc::KCAP - 1) / tc::KCAP; }

// Small batches (<= 8 users, gated / concat fusion, K <= 64): a unit's 16 user slots would be mostly empty, so each slot becomes a
// (user, item sub-range) pair instead -- `lists` sub-ranges of `sub_rows` rows per user, sized so that the slots of all CTA
// pairs are busy -- and the per-slot lists are merged afterwards in one or two K4 passes (G groups of J <= 64 lists).
struct SpreadPlan { bool on; int sub_rows, lists, G, J; int64_t n_virtual; int n_groups, n_pairs; };

static SpreadPlan tc_spread_plan(const pxr_handle* h, int64_t n_users, int32_t k) {
  SpreadPlan sp;
  memset(&sp, 0, sizeof(sp));
  // h->small_batch (pxr_set_small_batch): 0 = never, 1 = whenever the shape allows it (tests), -1 = when the cost model below says so
  if (h->small_batch == 0 || h->cfg.fusion == PXR_FUSION_ATTENTION || pxr_tc_gated_wide(h) || n_users > 8 || n_users <= 0 || tc_pages(k) > 1 || h->n_rows < 512) return sp;
  const int max_pairs = h->n_sm / 2;
  const int64_t slots = (int64_t)max_pairs * 16;
  int64_t sub = (h->n_rows * n_users + slots - 1) / slots;
  sub = std::max<int64_t>(128, (sub + 15) / 16 * 16);                      // >= 8 tiles per slot: list warm-up and unit setup amortised
  sub = std::min<int64_t>(sub, (h->n_rows + 15) / 16 * 16);
  sp.sub_rows = (int)sub;
  sp.lists = (int)((h->n_rows + sub - 1) / sub);
  if (sp.lists < 2 || sp.lists > 4096) return sp;
  sp.G = (sp.lists + 63) / 64;
  sp.J = (sp.lists + sp.G - 1) / sp.G;
  sp.n_virtual = n_users * sp.lists;
  sp.n_groups = (int)((sp.n_virtual + 15) / 16);
  sp.n_pairs = std::min(max_pairs, sp.n_groups);
  // worth it?  Measured model (profiles/r02_sweep_small_batch.jsonl): a spread call costs ~110 us (every slot warms up its
  // own list, 8x the record fetches per tile) + 7.4 us per tile of a slot; the plain shape ~20 us + 5.7 us per tile of a unit
  const TcPlan pl = tc_plan(h, n_users);
  const double t_plain = 20.0 + 5.7 * (double)((pl.rows_per_split + 15) / 16) * (double)((pl.n_units + max_pairs - 1) / max_pairs);
  const double t_spread = 110.0 + 7.4 * (double)(sp.sub_rows / 16) * (double)((sp.n_groups + max_pairs - 1) / max_pairs);
  sp.on = h->small_batch == 1 || t_spread < t_plain;
  return sp;
}

// workspace: [S > 1: per-split partial lists][exact mode: the merged 64-slot lists + the re-score pair arrays]
// top_k > 64 (pages): [partial lists][one page list][candidate rows of 64 * pages slots][page bounds][re-score arrays]
size_t pxr_tc_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k) {
  if (n_users <= 0 || h->n_rows <= 0) return 256;
  const TcPlan pl = tc_plan(h, n_users);
  const int P = tc_pages(k);
  size_t b = 256;
  if (P > 1) {
    const size_t page = (size_t)n_users * tc::KCAP * 8;
    if (pl.S > 1) b += pxr_align_up((size_t)pl.S * page, 256);
    b += pxr_align_up(page, 256) + pxr_align_up(page * P, 256) + pxr_align_up((size_t)n_users * 8, 256);
    if (h->rescore) b += pxr_rescore_list_bytes(n_users, P * tc::KCAP);
    return b;
  }
  const int kk = h->rescore ? tc::KCAP : k;
  const SpreadPlan sp = tc_spread_plan(h, n_users, k);
  if (sp.on) b += pxr_align_up((size_t)sp.J * sp.G * n_users * kk * 8, 256) + pxr_align_up((size_t)sp.G * n_users * kk * 8, 256);
  else if (pl.S > 1) b += pxr_align_up((size_t)pl.S * n_users * kk * 8, 256);
  if (h->rescore) b += pxr_align_up((size_t)n_users * kk * 8, 256) + pxr_rescore_list_bytes(n_users, tc::KCAP);
  return b;
}

static int tc_launch(pxr_handle* h, const tc::Params& p, int n_pairs, cudaStream_t st) {
  const int fmt = tc_fmt(h);
  switch (h->cfg.activation) {
    case PXR_ACT_RELU: return fmt == tc::FMT_BF16 ? tc::launch_fused_act<PXR_ACT_RELU>(h, p, h->cfg.fusion, fmt, n_pairs, st)
                                                  : pxr_tc_launch_relu_fp16(h, p, h->cfg.fusion, fmt, n_pairs, st);
    case PXR_ACT_GELU: return pxr_tc_launch_act1(h, p, h->cfg.fusion, fmt, n_pairs, st);
    case PXR_ACT_TANH: return pxr_tc_launch_act2(h, p, h->cfg.fusion, fmt, n_pairs, st);
    case PXR_ACT_LEAKY_RELU: return pxr_tc_launch_act3(h, p, h->cfg.fusion, fmt, n_pairs, st);
    case PXR_ACT_SILU: return pxr_tc_launch_act4(h, p, h->cfg.fusion, fmt, n_pairs, st);
    default: PXR_FAIL(h, PXR_ERR_INVALID, "unknown fusion_activation %d", h->cfg.activation);
  }
}

int pxr_tc_score_topk(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                      const int64_t* seen_indptr, const int32_t* seen_idx, int32_t k, float* out_scores, int32_t* out_idx,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!pxr_tc_can_run(h, k)) PXR_FAIL(h, PXR_ERR_INVALID, "tcgen05 path cannot run this call (top_k <= %d, 28-bit item index)", PXR_TC_MAX_K);
  const TcPlan pl = tc_plan(h, n_users);
  tc::FastWeights* fw = reinterpret_cast<tc::FastWeights*>(h->fast_w);
  const bool gated = h->cfg.fusion == PXR_FUSION_GATED;
  const bool attn = h->cfg.fusion == PXR_FUSION_ATTENTION;
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.wimg = fw->wimg; p.bias = fw->bias; p.gate_w = h->gate.w;
  p.item_feats = h->item_feats;
  p.item_logit = gated ? (const float*)h->item_fast : nullptr;
  p.item_pi = (gated || attn) ? nullptr : (const uint16_t*)h->item_fast;
  if (pxr_tc_gated_wide(h))      // [gate logits][item partials], see pxr_tc_item_bytes
    p.item_q = (const uint16_t*)((const char*)h->item_fast + pxr_align_up((size_t)((h->n_rows + 31) / 32 * 32) * 8 * sizeof(float), 256));
  if (attn) {
    p.attn_rec = (const uint4*)h->item_fast;
    p.wo_frag = fw->wo_frag;
    p.xc0_scratch = reinterpret_cast<uint4*>((char*)h->fast_w + pxr_align_up(sizeof(tc::FastWeights), 256));
    p.attn_in_wt = h->attn_in.wt; p.attn_in_b = h->attn_in.b; p.attn_out_wt = h->attn_out.wt; p.attn_out_b = h->attn_out.b;
    p.ln_w = h->ln_w; p.ln_b = h->ln_b;
    memcpy(p.bias_c, h->tc_bias_host, sizeof(float) * (tc::H1 + tc::H2 + 2 * tc::H3 + 1));
  }
  p.w1u_t = tc_l1_partials(h) ? pxr_tc_w1t_padded(h) : h->mlp[0].wt;     // [k][512] (zero-padded columns): rows 0..D-1 are the user columns of W1
  p.user_emb = user_embedding; p.user_idx = user_idx; p.seen_indptr = seen_indptr; p.seen_idx = seen_idx;
  p.item_missing = h->item_missing;
  p.n_users = n_users; p.n_rows = h->n_rows; p.item_base = h->item_base;
  p.Dm = h->cfg.embedding_dim;
  // exact mode: the kernel keeps its full 64-slot list per user (admission threshold = 64th best); the lists are
  // re-scored in fp32 and re-ranked afterwards (pxr_launch_rescore)
  const bool exact = h->rescore;
  const int P = tc_pages(k);
  const int32_t kk = (exact || P > 1) ? tc::KCAP : k;
  p.M = h->M; p.K = kk; p.S = pl.S; p.rows_per_split = pl.rows_per_split; p.n_units = pl.n_units;
  p.final_act = h->cfg.final_activation;
  if (ws_bytes < pxr_tc_topk_bytes(h, n_users, k)) PXR_FAIL(h, PXR_ERR_WORKSPACE, "tcgen05 top-K workspace too small");
  char* wp = (char*)ws;
  if (P > 1) {
    // ---- top_k > 64: one pass of the fused kernel per 64-slot page, each bounded by the previous page's last key
    const size_t page = (size_t)n_users * tc::KCAP;
    float* part_s = nullptr; int32_t* part_i = nullptr;
    if (pl.S > 1) { part_s = (float*)wp; part_i = (int32_t*)(wp + (size_t)pl.S * page * 4); wp += pxr_align_up((size_t)pl.S * page * 8, 256); }
    float* page_s = (float*)wp; int32_t* page_i = (int32_t*)(wp + page * 4);
    wp += pxr_align_up(page * 8, 256);
    float* cand_s = (float*)wp; int32_t* cand_i = (int32_t*)(wp + page * P * 4);
    wp += pxr_align_up(page * P * 8, 256);
    unsigned long long* upper = (unsigned long long*)wp;
    wp += pxr_align_up((size_t)n_users * 8, 256);
    for (int pg = 0; pg < P; ++pg) {
      p.upper = pg ? upper : nullptr;
      p.out_scores = pl.S > 1 ? part_s : page_s; p.out_idx = pl.S > 1 ? part_i : page_i;
      int rc = tc_launch(h, p, pl.n_pairs, st);
      if (rc) return rc;
      if (pl.S > 1) {
        rc = pxr_launch_merge(part_s, part_i, pl.S, n_users, tc::KCAP, page_s, page_i, st);
        h->launches++;
        if (rc) PXR_FAIL(h, rc, "top-K merge of %d item splits failed", pl.S);
      }
      tc::page_commit_kernel<<<(unsigned)((page + 255) / 256), 256, 0, st>>>(page_s, page_i, n_users, pg, P, cand_s, cand_i, upper);
      h->launches++;
      PXR_CUDA(h, cudaGetLastError());
    }
    const int L = P * tc::KCAP;
    if (exact) return pxr_launch_rescore(h, user_embedding, user_idx, n_users, cand_i, L, k, out_scores, out_idx, wp, st);
    // raw mode: the pages are already in descending key order, the first k slots of a row are the list
    PXR_CUDA(h, cudaMemcpy2DAsync(out_scores, (size_t)k * 4, cand_s, (size_t)L * 4, (size_t)k * 4, (size_t)n_users, cudaMemcpyDeviceToDevice, st));
    PXR_CUDA(h, cudaMemcpy2DAsync(out_idx, (size_t)k * 4, cand_i, (size_t)L * 4, (size_t)k * 4, (size_t)n_users, cudaMemcpyDeviceToDevice, st));
    return PXR_OK;
  }
  const SpreadPlan sp = tc_spread_plan(h, n_users, k);
  if (sp.on) {
    // ---- small batches: (user, item sub-range) slots, then one or two merge passes over the per-slot lists
    const size_t lst = (size_t)n_users * kk;                   // entries of one list set
    const size_t n_all = (size_t)sp.J * sp.G * lst;
    float* part_s = (float*)wp; int32_t* part_i = (int32_t*)(wp + n_all * 4);
    wp += pxr_align_up(n_all * 8, 256);
    float* tmp_s = (float*)wp; int32_t* tmp_i = (int32_t*)(wp + (size_t)sp.G * lst * 4);
    wp += pxr_align_up((size_t)sp.G * lst * 8, 256);
    float* list_s = out_scores; int32_t* list_i = out_idx;
    if (exact) { list_s = (float*)wp; list_i = (int32_t*)(wp + lst * 4); wp += pxr_align_up(lst * 8, 256); }
    if ((size_t)sp.J * sp.G > (size_t)sp.lists)                // lists that do not exist: empty (index -1)
      PXR_CUDA(h, cudaMemsetAsync(part_i + (size_t)sp.lists * lst, 0xFF, ((size_t)sp.J * sp.G - sp.lists) * lst * 4, st));
    p.n_real = n_users; p.n_users = sp.n_virtual; p.sub_rows = sp.sub_rows;
    p.S = 1; p.rows_per_split = (int)((h->n_rows + 15) / 16 * 16); p.n_units = sp.n_groups;
    p.out_scores = part_s; p.out_idx = part_i;
    int rc = tc_launch(h, p, sp.n_pairs, st);
    if (rc) return rc;
    if (sp.G == 1) {
      rc = pxr_launch_merge(part_s, part_i, sp.lists, n_users, kk, list_s, list_i, st);
    } else {       // list l = j * G + g: one pass merges, for every (g, user), its J lists; the second pass the G results
      rc = pxr_launch_merge(part_s, part_i, sp.J, (int64_t)sp.G * n_users, kk, tmp_s, tmp_i, st);
      h->launches++;
      if (!rc) rc = pxr_launch_merge(tmp_s, tmp_i, sp.G, n_users, kk, list_s, list_i, st);
    }
    h->launches++;
    if (rc) PXR_FAIL(h, rc, "top-K merge of %d sub-range lists failed", sp.lists);
    if (exact) return pxr_launch_rescore(h, user_embedding, user_idx, n_users, list_i, tc::KCAP, k, out_scores, out_idx, wp, st);
    return PXR_OK;
  }
  float* part_s = out_scores; int32_t* part_i = out_idx;       // what the kernel writes
  float* list_s = out_scores; int32_t* list_i = out_idx;       // the merged lists
  if (pl.S > 1) {
    const size_t need = (size_t)pl.S * n_users * kk;
    part_s = (float*)wp; part_i = (int32_t*)(wp + need * 4);
    wp += pxr_align_up(need * 8, 256);
  }
  if (exact) {
    const size_t need = (size_t)n_users * kk;
    list_s = (float*)wp; list_i = (int32_t*)(wp + need * 4);
    wp += pxr_align_up(need * 8, 256);
    if (pl.S == 1) { part_s = list_s; part_i = list_i; }
  }
  p.out_scores = part_s; p.out_idx = part_i;
  int rc = tc_launch(h, p, pl.n_pairs, st);
  if (rc) return rc;
  if (pl.S > 1) {
    rc = pxr_launch_merge(part_s, part_i, pl.S, n_users, kk, list_s, list_i, st);
    h->launches++;
    if (rc) PXR_FAIL(h, rc, "top-K merge of %d item splits failed", pl.S);
  }
  if (exact) return pxr_launch_rescore(h, user_embedding, user_idx, n_users, list_i, tc::KCAP, k, out_scores, out_idx, wp, st);
  return PXR_OK;
}
#endif  // PXR_TC_TU == 0
