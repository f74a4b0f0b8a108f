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

#include <cuda_fp16.h>

#include "pxr_common.cuh"
#include "tc_ptx.cuh"

namespace tc {

constexpr int D = 64, H1 = 512, H2 = 256, H3 = 128;
constexpr int TU = 8, TI = 16;            // users x items per CTA tile (128 rows)
constexpr int KCAP = 64;                  // slots of the per-user sorted list (K <= KCAP)
constexpr int QCAP = 512;                 // candidate queue entries
constexpr int THREADS = 512;                 // gated / concat: 16 warps
constexpr int THREADS_ATT = 640;             // attention: a second front-end warpgroup (warps 16-19)
constexpr uint32_t IDX_MASK = 0x0FFFFFFFu;   // 28-bit item index inside a queue / list key
constexpr int FMT_BF16 = 0, FMT_FP16 = 1;
constexpr int F_CONCAT = 0, F_GATED = 1, F_ATTN = 2;   // front ends of the kernel template
template <int FUS> __host__ __device__ constexpr int n_threads() { return FUS == F_ATTN ? THREADS_ATT : THREADS; }
constexpr int NH = 4, DH = D / NH;                      // attention: heads x head dim (fast path: 4 x 16)
// attention: per-item record, ATT_TOKENS token blocks of ATT_TOK floats (see item_attn_kernel)
constexpr int ATT_TOKENS = 5, ATT_TOK = 720, ATT_ITEM = ATT_TOKENS * ATT_TOK;
constexpr int ATT_C = 0, ATT_NB = 64, ATT_U = 320, ATT_Q = 576, ATT_K = 640, ATT_L = 704;
#ifndef PXR_TOPK_IDLE_NS
#define PXR_TOPK_IDLE_NS 200
#endif

enum {
  BAR_W = 0, BAR_A_FULL, BAR_A_EMPTY, BAR_D1_FULL0, BAR_D1_FULL1, BAR_D1_FULL2, BAR_D1_FULL3, BAR_H1_FULL0, BAR_H1_FULL1,
  BAR_H1_FULL2, BAR_H1_FULL3, BAR_H1_EMPTY0, BAR_H1_EMPTY1, BAR_H1_EMPTY2, BAR_H1_EMPTY3, BAR_PI_FULL0, BAR_PI_FULL1,
  BAR_PI_EMPTY0, BAR_PI_EMPTY1, BAR_D2_FULL, BAR_H2_FULL0, BAR_H2_FULL1, BAR_D3_FULL, BAR_D3_EMPTY, BAR_UNIT_DONE,
  BAR_UNIT_RESET, N_BARS
};

// Per-CTA scratch in shared memory.  The attention front end needs 12 KB of per-user constants next to the weights,
// so that variant keeps the epilogue biases in the kernel-parameter constant bank, a shorter candidate queue, and
// stages E_u in the (idle) A1 tile during the per-unit setup.
template <int FUS>
struct MiscT {
  static constexpr bool ATT = (FUS == F_ATTN);
  static constexpr int QC = ATT ? 128 : QCAP;          // candidate queue entries
  float b1[ATT ? 4 : H1]; float b2[ATT ? 4 : H2]; float b3[ATT ? 4 : H3]; float w4[ATT ? 4 : H3];   // contiguous
  float eu[ATT ? 1 : TU][D];
  float lu[ATT ? 1 : TU][8];
  float qu[ATT ? TU : 1][D], ku[ATT ? TU : 1][D];       // attention: in_proj q / k of the user token (unscaled)
  float U0c[ATT ? TU : 1][NH][D];                       // attention: per-head out_proj of v_u, centred over d
  float S00[ATT ? TU : 1][NH];                          // attention: q_u,h . k_u,h / sqrt(dh)
  unsigned long long list[TU][KCAP];
  unsigned long long queue[QC];
  float thr[TU];
  uint32_t seen_mask[4][TU];
  uint32_t q_tail[2], q_head[2];                     // two candidate queues: users 0-3 -> top-K warp 5, users 4-7 -> warp 6
  uint32_t tmem_base;
  float b4;
  unsigned long long bars[N_BARS];
};

// shared / tensor memory maps per front end
template <int FUS>
struct Map {
  static constexpr bool GATED = (FUS != F_CONCAT);               // layer 1 on the tensor pipe from an A1 tile
  // shared memory (bytes from a 1024-aligned base); the weight image is the first WIMG bytes
  static constexpr uint32_t OFF_W1 = 0;                          // gated: 8 N-chunks x (32 rows x 128 B) = 32 KB
  static constexpr uint32_t OFF_W2 = GATED ? 32768u : 0u;        // 8 K-blocks x (128 rows x 128 B) = 128 KB
  static constexpr uint32_t OFF_W3 = OFF_W2 + 131072u;           // 4 K-blocks x (64 rows x 128 B)  = 32 KB
  static constexpr uint32_t WIMG = OFF_W3 + 32768u;
  static constexpr uint32_t OFF_A1 = WIMG;                       // gated: 128 rows x 128 B, SWIZZLE_128B
  static constexpr uint32_t PI_STRIDE = 1040, PI_BUF = TI * PI_STRIDE;   // concat: 16 item partials (512 x 16 bit)
  static constexpr uint32_t OFF_PI = WIMG;                       // padded by 16 B per row: conflict-free reads
  static constexpr uint32_t PU_STRIDE = 2064;                    // concat: 8 user partials (512 fp32), padded
  static constexpr uint32_t OFF_PU = OFF_PI + 2 * PI_BUF;
  static constexpr uint32_t OFF_MISC = GATED ? OFF_A1 + 16384u : OFF_PU + TU * PU_STRIDE;
  static constexpr uint32_t SMEM = OFF_MISC + (uint32_t)sizeof(MiscT<FUS>) + 1024u;   // + alignment slack
  // tensor memory (columns)
  static constexpr uint32_t TM_D3 = 128, TM_D2 = 256;
  // H1 chunk buffer b.  gated: 64-col fp32 accumulators packed in place, buffers 2,3 share the D3 columns
  // (D3 is only live at the tile boundary).  concat: four 32-col packed buffers.
  __host__ __device__ static constexpr uint32_t h1buf(int b) {
    return GATED ? (b < 2 ? 64u * b : TM_D3 + 64u * (b - 2)) : 32u * b;
  }
};
static_assert(Map<F_GATED>::SMEM <= 232448 && Map<F_CONCAT>::SMEM <= 232448 && Map<F_ATTN>::SMEM <= 232448, "shared memory budget");
static_assert(Map<F_CONCAT>::OFF_MISC % 16 == 0 && Map<F_GATED>::OFF_MISC % 16 == 0, "alignment");
static_assert(offsetof(MiscT<F_ATTN>, qu) % 16 == 0 && offsetof(MiscT<F_ATTN>, U0c) % 16 == 0 && offsetof(MiscT<F_ATTN>, list) % 8 == 0, "alignment");
static_assert(offsetof(MiscT<F_GATED>, list) % 8 == 0 && offsetof(MiscT<F_GATED>, bars) % 8 == 0 && offsetof(MiscT<F_ATTN>, bars) % 8 == 0, "alignment");

// attention: E_u + out_proj bias, centred over d, of one CTA's 8 users (rebuilt per unit by the front-end warps).
// Read once per half tile, so it lives in global memory (one slot per CTA); the hot per-user constants are in smem.
struct UserAttn {
  float C0c[TU][D];
};

struct Params {
  const uint8_t* wimg;          // [2][WIMG] pre-swizzled 16-bit operand images (rank 0, rank 1)
  const float* bias;            // b1[512] b2[256] b3[128] w4[128] b4   (b1 unused by concat: folded into Pi)
  const float* gate_w;          // gated: (M, M*D) fp32 row-major; the user part is the first D of each row
  const float* item_feats;      // gated: [rows][M-1][D] fp32 projected item-side modality vectors
  const float* item_logit;      // gated: [rows][8] fp32 item part of the gate logits (+ gate bias)
  const uint16_t* item_pi;      // concat: [rows][512] 16-bit item partial of layer 1 (+ b1)
  const float* w1u_t;           // concat: [64][512] fp32, user columns of W1 transposed
  const float* attn_rec;        // attention: [rows][ATT_ITEM] fp32 per-item records
  const float* attn_in_wt;      // attention: in_proj weight transposed [64][192], bias [192]
  const float* attn_in_b;
  const float* attn_out_wt;     // attention: out_proj weight transposed [64][64], bias [64]
  const float* attn_out_b;
  const float* ln_w;            // attention: LayerNorm weight / bias [64]
  const float* ln_b;
  UserAttn* user_scratch;       // attention: one UserAttn per CTA
  float bias_c[H1 + H2 + H3 + H3 + 4];   // attention: b1' b2 b3 w4 b4 read through the constant bank (no room in smem)
  const float* user_emb;        // (n_users_total, D) fp32 table
  const int64_t* user_idx;      // (n_users,)
  const int64_t* seen_indptr;   // (n_users + 1,) or NULL
  const int32_t* seen_idx;      // global item indices, ascending per user
  const uint8_t* item_missing;  // per item row: 1 = features missing, the score is 0.0 (recommender.py:229-230); NULL = none
  float* out_scores;            // [S][n_users][K]
  int32_t* out_idx;
  int64_t n_users, n_rows, item_base;
  int M, K, S, rows_per_split, n_units, final_act;
  int Dm;                       // embedding dim of the model (concat: any multiple of 4 up to 512; gated / attention: 64)
};

struct Unit { int g, s; int64_t row_lo, row_hi; int ntiles; };

__device__ __forceinline__ Unit decode_unit(const Params& p, int w) {
  Unit u;
  u.g = w / p.S; u.s = w % p.S;
  u.row_lo = (int64_t)u.s * p.rows_per_split;
  u.row_hi = min(p.n_rows, u.row_lo + (int64_t)p.rows_per_split);
  u.ntiles = (int)((u.row_hi - u.row_lo + TI - 1) / TI);
  return u;
}

// ---------------------------------------------------------------------------------------------
// 16-bit operand format helpers
// ---------------------------------------------------------------------------------------------
template <int FMT> __device__ __forceinline__ uint32_t relu_pack(float lo, float hi) {
  uint32_t d;
  if (FMT == FMT_BF16) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));   // saturate: no inf from fp16 range
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
template <int FMT>
__device__ __forceinline__ void bias_relu_pack32(const uint32_t* v, const float* bias, uint32_t* o) {
  const float4* bb = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = bb[q];
    o[2 * q] = relu_pack<FMT>(__uint_as_float(v[4 * q]) + b.x, __uint_as_float(v[4 * q + 1]) + b.y);
    o[2 * q + 1] = relu_pack<FMT>(__uint_as_float(v[4 * q + 2]) + b.z, __uint_as_float(v[4 * q + 3]) + b.w);
  }
}

// 64 fp32 accumulator columns -> 32 packed 16-bit columns written over the start of the same region
template <int FMT>
__device__ __forceinline__ void epi_pack64(uint32_t t_src, uint32_t t_dst, const float* bias) {
  uint32_t v0[32], v1[32], o[32];
  ptx::tmem_ld32(t_src, v0);
  ptx::tmem_ld32(t_src + 32, v1);
  ptx::tc_wait_ld();
  bias_relu_pack32<FMT>(v0, bias, o);
  bias_relu_pack32<FMT>(v1, bias + 32, o + 16);
  ptx::tmem_st32(t_dst, o);
}

// concat layer 1 for one row and one 64-wide chunk: relu(Pu[user] + Pi[item]) -> 32 packed columns
template <int FMT>
__device__ __forceinline__ void concat_h1_chunk(const uint8_t* pi_row, const uint8_t* pu_row, uint32_t t_dst) {
  uint32_t o[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 pv = *reinterpret_cast<const uint4*>(pi_row + 16 * q);
    const float4 a = *reinterpret_cast<const float4*>(pu_row + 32 * q);
    const float4 b = *reinterpret_cast<const float4*>(pu_row + 32 * q + 16);
    const float2 p0 = unpack2<FMT>(pv.x), p1 = unpack2<FMT>(pv.y), p2 = unpack2<FMT>(pv.z), p3 = unpack2<FMT>(pv.w);
    o[4 * q + 0] = relu_pack<FMT>(a.x + p0.x, a.y + p0.y);
    o[4 * q + 1] = relu_pack<FMT>(a.z + p1.x, a.w + p1.y);
    o[4 * q + 2] = relu_pack<FMT>(b.x + p2.x, b.y + p2.y);
    o[4 * q + 3] = relu_pack<FMT>(b.z + p3.x, b.w + p3.y);
  }
  ptx::tmem_st32(t_dst, o);
}


// ---------------------------------------------------------------------------------------------
// attention fusion front end (src/models/layers.py:135-164 as documented; SURVEY.md A5 split)
//
// Tokens of pair (u, i): x_0 = E_u, x_1.. = the item-side modality vectors.  Everything that involves only
// item tokens is folded into the per-item record once per catalogue (item_attn_kernel):
//   row a >= 1:  softmax over [s_a0 | item part] => with L_ah = logsumexp_b>=1 s_ab,h the user column gets weight
//                w_ah = sigmoid(q_a,h . k_u,h / sqrt(dh) - L_ah) and the item columns share 1 - w_ah, so
//                y_a = x_a + attn_a = C_a + sum_h w_ah (U0_h[u] - Nbar_ah),   C_a = x_a + b_o + sum_h Nbar_ah,
//                Nbar_ah = sum_b>=1 softmax_b(s_ab,h) U_bh,   U_bh = W_o[:, head h] v_b,h   (out_proj folded in)
//   row 0:       softmax over [q_u.k_u, q_u.k_b ...] per head, y_0 = E_u + b_o + sum_h (p_h0 U0_h + sum_b p_hb U_bh)
// LayerNorm only needs y - mean(y): every stored vector (C, Nbar, U, U0, E_u + b_o) is CENTRED over d in advance,
// and as the weights of each row sum to one the combination is centred too -- no per-pair mean.  Then
//   fused = ln_b + ln_w / M * sum_a (y_a - mean) * rstd_a;
// the affine part is folded into layer 1 (W1' = W1 diag(ln_w / M), b1' = b1 + W1 ln_b), so the A1 operand is
// the 16-bit rounding of   acc = sum_a (y_a - mean) * rstd_a.
// Thread = (item j of a half tile of 8, 4-wide slice s of D; head of the slice = s / 4), all 8 users of the CTA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float* p) {      // item records: read once per tile, keep them out of L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
__device__ __forceinline__ void fma4(float* acc, float w, const float4& v) {
  acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]); acc[2] = fmaf(w, v.z, acc[2]); acc[3] = fmaf(w, v.w, acc[3]);
}

// per-unit constants of this CTA's 8 users (128 threads, named barrier 1).  `stage` = 4 KB of scratch shared memory
// (the idle A1 tile): E_u at [0, 2 KB), v_u at [2 KB, 4 KB).
template <class MiscA>
__device__ __forceinline__ void attn_user_setup(const Params& p, MiscA& ms, float* stage, UserAttn* us, int tid) {
  float (*eu)[D] = reinterpret_cast<float (*)[D]>(stage);
  float (*vu)[D] = reinterpret_cast<float (*)[D]>(stage + TU * D);
  for (int n = tid; n < 3 * D; n += 128) {                 // in_proj of the user token: q | k | v
    float acc[TU];
    const float b = p.attn_in_b[n];
#pragma unroll
    for (int u = 0; u < TU; ++u) acc[u] = b;
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
      const float w = p.attn_in_wt[k * 3 * D + n];
#pragma unroll
      for (int u = 0; u < TU; ++u) acc[u] = fmaf(w, eu[u][k], acc[u]);
    }
    float* dst = n < D ? &ms.qu[0][n] : (n < 2 * D ? &ms.ku[0][n - D] : &vu[0][n - 2 * D]);
#pragma unroll
    for (int u = 0; u < TU; ++u) dst[u * D] = acc[u];
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  {
    const int d = tid & 63, hp = tid >> 6;                 // out_proj per head: U0[u][h][d] = sum_e W_o[d][16h+e] v_u[16h+e]
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int h = 2 * hp + hh;
      float acc[TU];
#pragma unroll
      for (int u = 0; u < TU; ++u) acc[u] = 0.f;
#pragma unroll 4
      for (int e = 0; e < DH; ++e) {
        const float w = p.attn_out_wt[(h * DH + e) * D + d];
#pragma unroll
        for (int u = 0; u < TU; ++u) acc[u] = fmaf(w, vu[u][h * DH + e], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < TU; ++u) ms.U0c[u][h][d] = acc[u];
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  // centre over d: one warp-level pass per 64-vector (8 users x (4 heads + C0) = 40 vectors, 10 per warp)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int v = warp; v < TU * (NH + 1); v += 4) {
      const int u = v / (NH + 1), h = v % (NH + 1);
      float a, b;
      if (h < NH) { a = ms.U0c[u][h][lane]; b = ms.U0c[u][h][lane + 32]; }
      else { a = eu[u][lane] + p.attn_out_b[lane]; b = eu[u][lane + 32] + p.attn_out_b[lane + 32]; }
      float sm_ = a + b;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sm_ += __shfl_xor_sync(0xffffffffu, sm_, o);
      const float mean = sm_ * (1.f / D);
      if (h < NH) { ms.U0c[u][h][lane] = a - mean; ms.U0c[u][h][lane + 32] = b - mean; }
      else { us->C0c[u][lane] = a - mean; us->C0c[u][lane + 32] = b - mean; }
    }
  }
  if (tid < TU * NH) {
    const int u = tid >> 2, h = tid & 3;
    float dot = 0.f;
    for (int e = 0; e < DH; ++e) dot = fmaf(ms.qu[u][h * DH + e], ms.ku[u][h * DH + e], dot);
    ms.S00[u][h] = dot * 0.25f;                            // 1 / sqrt(dh), dh = 16
  }
}

// v[u4]: partial sums of 4 users held by every lane of a 4-lane head quad -> reduce-scatter: lane ql of the quad ends
// up with the total of user u4 = ql (3 shuffles instead of 8, and the value is then transformed once, not four times)
__device__ __forceinline__ float quad_scatter4(const float* v, int ql) {
  const bool b1 = ql & 2, b0 = ql & 1;
  const float k0 = (b1 ? v[2] : v[0]) + __shfl_xor_sync(0xffffffffu, b1 ? v[0] : v[2], 2);
  const float k1 = (b1 ? v[3] : v[1]) + __shfl_xor_sync(0xffffffffu, b1 ? v[1] : v[3], 2);
  return (b0 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, b0 ? k0 : k1, 1);
}

// y[u4][4]: 4 users x this lane's 4 dims, already centred over d.  r[u4] = LayerNorm rstd of each user: the sums of
// squares are reduce-scattered over the 16 lanes of the item (5 shuffles), one rsqrt per lane, 4 indexed gathers.
__device__ __forceinline__ void ln_rstd4(const float (*y)[4], int lane, float* r) {
  float ss[4];
#pragma unroll
  for (int u4 = 0; u4 < 4; ++u4) ss[u4] = fmaf(y[u4][3], y[u4][3], fmaf(y[u4][2], y[u4][2], fmaf(y[u4][1], y[u4][1], y[u4][0] * y[u4][0])));
  const bool b3 = lane & 8, b2 = lane & 4;
  const float k0 = (b3 ? ss[2] : ss[0]) + __shfl_xor_sync(0xffffffffu, b3 ? ss[0] : ss[2], 8);
  const float k1 = (b3 ? ss[3] : ss[1]) + __shfl_xor_sync(0xffffffffu, b3 ? ss[1] : ss[3], 8);
  float t = (b2 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, b2 ? k0 : k1, 4);       // user u4 = 2 b3 + b2, summed over 4 lanes
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  const float rr = rsqrtf(fmaf(t, 1.f / D, 1e-5f));
#pragma unroll
  for (int u4 = 0; u4 < 4; ++u4) r[u4] = __shfl_sync(0xffffffffu, rr, (lane & 19) | (u4 << 2));
}

// acc = sum over tokens of the normalised rows, for the 8 users x 8 items of one half tile -> 16-bit A1 rows.
// `wait_a1` is called once, right before the first store into A1.
// The front end is latency-bound per warp (two warps per scheduler), so latency is hidden by instruction-level
// parallelism: every step is a loop over 4 users (independent chains) with its shuffles issued back to back, and
// cross-lane sums are reduce-scatters + indexed gathers (28 shuffles per (token, 4 users) instead of 40; one
// softmax / sigmoid / rsqrt per value instead of one per lane).
template <int FMT, class MiscA, typename WaitA1>
__device__ __forceinline__ void attn_half_tile(const MiscA& ms, const UserAttn* us, const float* rec, int nt, uint8_t* a1,
                                               int half, int tid, int lane, WaitA1 wait_a1) {
  const int j8 = tid >> 4, s = tid & 15, hd = s >> 2, ql = s & 3;
  const int gb = lane & 16;                                // first lane of this item's 16-lane group
  float acc[TU][4];
  // token data of the first item token: fetched now, consumed after the user-token row
  float4 c, q, nb[NH]; float Lh;
  auto load_tok = [&](int a, float4& c_, float4& q_, float4* nb_, float& L_) {
    const float* tr = rec + a * ATT_TOK;
    c_ = ldg_stream(tr + ATT_C + 4 * s);
    q_ = ldg_stream(tr + ATT_Q + 4 * s);
#pragma unroll
    for (int h = 0; h < NH; ++h) nb_[h] = ldg_stream(tr + ATT_NB + h * D + 4 * s);
    L_ = __ldg(tr + ATT_L + hd);
  };
  {
    // ---- token 0 (the user token): scores against every token, softmax per head, value mix, normalise
    float pw[2][ATT_TOKENS + 1];                           // softmax weights of users ql and 4 + ql, head hd: [0] user token, [1 + b] item token b
    {
      float4 kb[ATT_TOKENS];
#pragma unroll
      for (int b = 0; b < ATT_TOKENS; ++b) kb[b] = ldg_stream(rec + min(b, nt - 1) * ATT_TOK + ATT_K + 4 * s);
#pragma unroll
      for (int u = 0; u < TU; ++u) {
        const float4 c0 = __ldcg(reinterpret_cast<const float4*>(&us->C0c[u][4 * s]));
        acc[u][0] = c0.x; acc[u][1] = c0.y; acc[u][2] = c0.z; acc[u][3] = c0.w;
      }
#pragma unroll
      for (int ug = 0; ug < 2; ++ug) {
        float S[ATT_TOKENS][4];
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const float4 qv = *reinterpret_cast<const float4*>(&ms.qu[4 * ug + u4][4 * s]);
#pragma unroll
          for (int b = 0; b < ATT_TOKENS; ++b) S[b][u4] = dot4(qv, kb[b]);
        }
        float sv[ATT_TOKENS];
#pragma unroll
        for (int b = 0; b < ATT_TOKENS; ++b) sv[b] = quad_scatter4(S[b], ql);
        const float s00 = ms.S00[4 * ug + ql][hd];
        float m = s00;
#pragma unroll
        for (int b = 0; b < ATT_TOKENS; ++b) { if (b >= nt) sv[b] = -INFINITY; m = fmaxf(m, sv[b]); }
        const float e0 = __expf(s00 - m);
        float sum = e0;
#pragma unroll
        for (int b = 0; b < ATT_TOKENS; ++b) { sv[b] = __expf(sv[b] - m); sum += sv[b]; }
        const float inv = __fdividef(1.f, sum);
        pw[ug][0] = e0 * inv;
#pragma unroll
        for (int b = 0; b < ATT_TOKENS; ++b) pw[ug][1 + b] = sv[b] * inv;
      }
    }
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float ph[TU];
#pragma unroll
      for (int u = 0; u < TU; ++u) ph[u] = __shfl_sync(0xffffffffu, pw[u >> 2][0], gb | (4 * h) | (u & 3));
#pragma unroll
      for (int u = 0; u < TU; ++u) fma4(acc[u], ph[u], *reinterpret_cast<const float4*>(&ms.U0c[u][h][4 * s]));
    }
#pragma unroll
    for (int b = 0; b < ATT_TOKENS; ++b) {
      float4 ub[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) ub[h] = ldg_stream(rec + min(b, nt - 1) * ATT_TOK + ATT_U + h * D + 4 * s);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float ph[TU];
#pragma unroll
        for (int u = 0; u < TU; ++u) ph[u] = __shfl_sync(0xffffffffu, pw[u >> 2][1 + b], gb | (4 * h) | (u & 3));   // weight 0 for b >= nt
#pragma unroll
        for (int u = 0; u < TU; ++u) fma4(acc[u], ph[u], ub[h]);
      }
    }
    load_tok(0, c, q, nb, Lh);
#pragma unroll
    for (int ug = 0; ug < 2; ++ug) {
      float r[4];
      ln_rstd4(&acc[4 * ug], lane, r);
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[4 * ug + u4][i] *= r[u4];
    }
  }
  // ---- item tokens a >= 1 (token data of a + 1 is fetched while a is being combined)
#pragma unroll 1
  for (int a = 0; a < nt; ++a) {
    float4 c2, q2, nb2[NH]; float L2;
    load_tok(min(a + 1, nt - 1), c2, q2, nb2, L2);
#pragma unroll
    for (int ug = 0; ug < 2; ++ug) {
      float d[4], y[4][4];
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) d[u4] = dot4(q, *reinterpret_cast<const float4*>(&ms.ku[4 * ug + u4][4 * s]));
      const float w = __fdividef(1.f, 1.f + __expf(Lh - quad_scatter4(d, ql)));      // sigmoid(s_a0 - L_ah) of user 4 ug + ql
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) { y[u4][0] = c.x; y[u4][1] = c.y; y[u4][2] = c.z; y[u4][3] = c.w; }
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float wh[4];
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) wh[u4] = __shfl_sync(0xffffffffu, w, gb | (4 * h) | u4);
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const float4 uv = *reinterpret_cast<const float4*>(&ms.U0c[4 * ug + u4][h][4 * s]);
          y[u4][0] = fmaf(wh[u4], uv.x - nb[h].x, y[u4][0]); y[u4][1] = fmaf(wh[u4], uv.y - nb[h].y, y[u4][1]);
          y[u4][2] = fmaf(wh[u4], uv.z - nb[h].z, y[u4][2]); y[u4][3] = fmaf(wh[u4], uv.w - nb[h].w, y[u4][3]);
        }
      }
      float r[4];
      ln_rstd4(y, lane, r);
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[4 * ug + u4][i] = fmaf(r[u4], y[u4][i], acc[4 * ug + u4][i]);
    }
    c = c2; q = q2; Lh = L2;
#pragma unroll
    for (int h = 0; h < NH; ++h) nb[h] = nb2[h];
  }
  // ---- 16-bit pack into the swizzled A1 rows (row = user * 16 + item; 8 bytes per thread)
  wait_a1();
  const int j = 8 * half + j8;
#pragma unroll
  for (int u = 0; u < TU; ++u) {
    uint2 pk;
    pk.x = pack2<FMT>(acc[u][0], acc[u][1]); pk.y = pack2<FMT>(acc[u][2], acc[u][3]);
    const int r = u * TI + j;
    *reinterpret_cast<uint2*>(a1 + (r >> 3) * 1024 + (r & 7) * 128 + (((s >> 1) ^ (r & 7)) << 4) + (s & 1) * 8) = pk;
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// FUS selects the front end.  GATED below means "layer 1 runs on the tensor pipe from an A1 tile in shared memory"
// (gated and attention fusion: the fused vector depends on the pair); concat feeds layer-1 partial sums instead.
// TK2: two top-K warps (short units, where list updates are a visible share of the work) instead of one.
template <int FUS, int FMT, bool TK2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(n_threads<FUS>(), 1)
score_fused_kernel(const __grid_constant__ Params p) {
  constexpr int NT = n_threads<FUS>();
  constexpr bool GATED = (FUS != F_CONCAT);
  constexpr bool ATT = (FUS == F_ATTN);
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
  if (threadIdx.x < TU) ms.thr[threadIdx.x] = -INFINITY;
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

  // attention: register budget per warpgroup.  640 threads are launched with 96 registers each and setmaxnreg can
  // only move registers inside that pool (61 440): the MMA / top-K warpgroup drops to 40, the two front-end
  // warpgroups take 128, the epilogue warpgroups drop to 88  (2*128*128 + 128*40 + 2*128*88 = 60 416).
  // (setmaxnreg sits at the top of each warpgroup's branch so that ptxas allocates per role).
  if (warp < 4 || (ATT && warp >= 16)) {
    if (ATT) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    // =============================================================== front end
    // attention: warpgroup A (warps 0-3) builds the per-unit user constants and the first half of every tile,
    // warpgroup B (warps 16-19) the second half; named barriers 2 / 3 fence the user constants between units.
    const bool feB = ATT && warp >= 16;
    const int tid = threadIdx.x & 127;      // 0..127 inside the front-end warpgroup
    const int Mm = p.M;
    int T = 0;
    for (int w = pair; w < p.n_units; w += n_pairs) {
      const Unit un = decode_unit(p, w);
      const int64_t ubase = ((int64_t)un.g * 2 + rank) * TU;     // first user ordinal of this CTA's group
      if (!GATED && T > 0) {
        // Pu is read by the layer-1 producers of the previous unit's last tile: wait until they are done with it
        ptx::mbar_wait(BAR(BAR_PI_EMPTY0 + ((T - 1) & 1)), ((T - 1) >> 1) & 1);
      }
      if (ATT) asm volatile("bar.sync 2, 256;" ::: "memory");      // both warpgroups are done with the previous unit
      if (!feB) {
      asm volatile("bar.sync 1, 128;" ::: "memory");              // previous unit's readers of eu/lu are done
      {
        // attention stages E_u (and v_u) in the A1 tile: wait until the layer-1 MMAs of the previous tile have read it
        if (ATT && T > 0) ptx::mbar_wait(BAR(BAR_A_EMPTY), (T - 1) & 1);
        float (*eu_dst)[D] = ATT ? reinterpret_cast<float (*)[D]>(sm + MP::OFF_A1) : ms.eu;
        const int u = tid >> 4, d4 = (tid & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ubase + u < p.n_users && d4 < p.Dm) v = *reinterpret_cast<const float4*>(p.user_emb + p.user_idx[ubase + u] * p.Dm + d4);
        *reinterpret_cast<float4*>(&eu_dst[u][d4]) = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (FUS == F_ATTN) {
        attn_user_setup(p, ms, reinterpret_cast<float*>(sm + MP::OFF_A1), p.user_scratch + blockIdx.x, tid);
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
        for (int k0 = 0;; k0 += D) {
          const int kn = min(D, p.Dm - k0);
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
            if (ubase + u < p.n_users && d4 < p.Dm) v = *reinterpret_cast<const float4*>(p.user_emb + p.user_idx[ubase + u] * p.Dm + d4);
            *reinterpret_cast<float4*>(&ms.eu[u][(tid & 15) * 4]) = v;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll
        for (int u = 0; u < TU; ++u)
          *reinterpret_cast<float4*>(sm + MP::OFF_PU + u * MP::PU_STRIDE + 16 * tid) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (ATT) asm volatile("bar.sync 3, 256;" ::: "memory");      // user constants of this unit are in place
      // seen-item cursors: lanes 0..7 of warp 0 walk user u's ascending history with the item sweep
      int64_t cur = 0, cend = 0; int32_t nextv = 0x7fffffff;
      if (warp == 0 && lane < TU && p.seen_indptr && ubase + lane < p.n_users) {
        cur = p.seen_indptr[ubase + lane]; cend = p.seen_indptr[ubase + lane + 1];
        const int32_t first = (int32_t)(p.item_base + un.row_lo);
        int64_t lo = cur, hi = cend;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (p.seen_idx[mid] < first) lo = mid + 1; else hi = mid; }
        cur = lo;
        nextv = cur < cend ? p.seen_idx[cur] : 0x7fffffff;
      }
      if (!GATED && warp != 0) { T += un.ntiles; continue; }     // concat: warp 0 alone stages the tiles
      for (int t = 0; t < un.ntiles; ++t, ++T) {
        const int64_t row0 = un.row_lo + (int64_t)t * TI;
        auto write_seen_mask = [&]() {         // lanes 0..7 of warp 0: the 16-bit seen mask of this tile per user
          if (warp == 0 && lane < TU) {
            uint32_t mask = 0;
            const int32_t i0 = (int32_t)(p.item_base + row0);
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
          {
            const int half = feB ? 1 : 0;
            const int64_t row = row0 + 8 * half + (tid >> 4);
            const int64_t rr = row < un.row_hi ? row : un.row_lo;    // padding rows recompute a valid item (discarded later)
            attn_half_tile<FMT>(ms, p.user_scratch + blockIdx.x, p.attn_rec + rr * ATT_ITEM, Mm - 1, sm + MP::OFF_A1, half, tid,
                                lane, [&]() { if (T > 0) ptx::mbar_wait(BAR(BAR_A_EMPTY), (T - 1) & 1); });
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
          if (tid == 0) {
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
      for (int w = pair; w < p.n_units; w += n_pairs) NT += decode_unit(p, w).ntiles;
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
    const int u_lo = TK2 ? 4 * qh : 0, u_hi = TK2 ? 4 * qh + 4 : TU;
    unsigned long long* const queue = ms.queue + qh * (QC / 2);
    uint32_t head = 0, done_ph = 0;
    for (int w = pair; w < p.n_units; w += n_pairs) {
      const Unit un = decode_unit(p, w);
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
          const int u = (int)((e >> 28) & 7ull);
          const unsigned long long key = e & 0xFFFFFFFF0FFFFFFFull;
          // sorted (descending) insertion into list[u]: lanes hold slots lane and lane + 32
          unsigned long long* L = ms.list[u];
          const unsigned long long a = L[lane], b = L[lane + 32];
          const unsigned ba = __ballot_sync(0xffffffffu, key > a), bb = __ballot_sync(0xffffffffu, key > b);
          const int pos = ba ? 32 - __popc(ba) : 64 - __popc(bb);     // entries >= key come first
          if (pos < p.K) {
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
    const int ru = r >> 4, rj = r & 15;              // user slot / item slot of the row
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
      epi_pack64<FMT>(tl + c0, tl + c0, PXR_CB2 + grp * 128);
      epi_pack64<FMT>(tl + c0 + 64, tl + c0 + 32, PXR_CB2 + grp * 128 + 64);
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
          z = fmaf(fmaxf(__uint_as_float(v[0]) + b.x, 0.f), wv.x, z);
          z = fmaf(fmaxf(__uint_as_float(v[1]) + b.y, 0.f), wv.y, z);
          z = fmaf(fmaxf(__uint_as_float(v[2]) + b.z, 0.f), wv.z, z);
          z = fmaf(fmaxf(__uint_as_float(v[3]) + b.w, 0.f), wv.w, z);
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
        const int qh = TK2 ? (ru >> 2) : 0;
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
        epi_pack64<FMT>(tl + MP::h1buf(b), tl + MP::h1buf(b), PXR_CB1 + c * 64);
      } else {
        const int buf = T & 1;
        if (ci == 0) ptx::mbar_wait(BAR(BAR_PI_FULL0 + buf), (T >> 1) & 1);          // this tile's item partials landed
        const uint32_t n = (ci & 1) ? h1use1++ : h1use0++;
        if (n > 0) { ptx::mbar_wait(BAR(BAR_H1_EMPTY0 + b), (n - 1) & 1); ptx::tc_fence_after(); }   // layer 2 consumed the buffer
        concat_h1_chunk<FMT>(sm + MP::OFF_PI + buf * MP::PI_BUF + rj * MP::PI_STRIDE + c * 128,
                             sm + MP::OFF_PU + ru * MP::PU_STRIDE + c * 256, tl + MP::h1buf(b));
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive_cluster(BAR(BAR_H1_FULL0 + b), 0);
        if (!GATED && ci == 3) ptx::mbar_arrive_local(BAR(BAR_PI_EMPTY0 + (T & 1)));   // done with this tile's Pi (and Pu)
      }
    };

    for (int w = pair; w < p.n_units; w += n_pairs) {
      const Unit un = decode_unit(p, w);
      const int64_t ubase = ((int64_t)un.g * 2 + rank) * TU;
      for (int t = 0; t < un.ntiles; ++t, ++T) {
        // One rolled loop over this group's four layer-1 chunks (the body must stay resident in the instruction
        // cache: fully unrolling it costs ~6 % throughput).  The deferred layer-2 / layer-3 epilogues of the
        // PREVIOUS tile are slotted in where their inputs become available:
        //   gated : E2(T-1), L1(0), E3(T-1), L1(1), L1(2), L1(3)
        //   concat: L1(0), E2(T-1), L1(1), E3(T-1), L1(2), L1(3)   (L1(0) refills a free buffer while layer 2 drains)
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
__global__ void build_wimg_kernel(const float* __restrict__ w1, int k1, const float* __restrict__ k1_scale,
                                  const float* __restrict__ w2, const float* __restrict__ w3, uint8_t* __restrict__ img,
                                  int gated, int fmt) {
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
    float v;
    if (which == 1) v = w1[(size_t)(blk * 64 + rank * 32 + nl) * k1 + kk] * (k1_scale ? k1_scale[kk] : 1.f);   // N-chunk blk, rows [32 rank, +32)
    else if (which == 2) v = w2[(size_t)(rank * 128 + nl) * H1 + blk * 64 + kk];     // K-block blk, rows [128 rank, +128)
    else v = w3[(size_t)(rank * 64 + nl) * H2 + blk * 64 + kk];
    reinterpret_cast<uint16_t*>(img + (size_t)rank * wimg)[off / 2] = to16(v, fmt);
  }
}

// gated: item part of the gate logits: Wg[:, D:] . concat(item-side vectors) + bg   (layers.py:207 split)
__global__ void item_logit_kernel(const float* __restrict__ feats, const float* __restrict__ gate_w,
                                  const float* __restrict__ gate_b, int M, int64_t n_rows, float* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int FD = (M - 1) * D;
  const float* x = feats + row * FD;
  for (int m = 0; m < 8; ++m) {
    float acc = 0.f;
    if (m < M) {
      const float* wr = gate_w + (size_t)m * M * D + D;
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
__global__ void __launch_bounds__(PXR_SIMT_THREADS) item_pi_kernel(const float* __restrict__ feats, const float* __restrict__ w1t,
                                                                    const float* __restrict__ b1, int M, int Dm, int64_t n_rows,
                                                                    uint16_t* __restrict__ out, int fmt) {
  extern __shared__ __align__(16) float smem_pi[];
  const int FD = (M - 1) * Dm;
  float* in = smem_pi;                       // [32][FD]
  float* res = smem_pi + 32 * FD;            // [32][512]
  const int64_t row0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * FD; i += PXR_SIMT_THREADS) {
    const int64_t row = row0 + i / FD;
    in[i] = row < n_rows ? feats[row * FD + i % FD] : 0.f;
  }
  __syncthreads();
  linear_rows<32>(in, FD, FD, w1t + (size_t)Dm * H1, b1, H1, res, H1, -1);
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * H1; i += PXR_SIMT_THREADS) {
    const int64_t row = row0 + i / H1;
    if (row < n_rows) out[row * H1 + i % H1] = to16(res[i], fmt);
  }
}


// attention: per-item record (once per catalogue shard).  4 items per block, thread = (item, d).  Record layout per
// item token a (ATT_TOK floats): C[64] | Nbar[4][64] | U[4][64] | q[64] / sqrt(dh) | k[64] / sqrt(dh) | L[4] | pad,
// with C, Nbar_h and U_h each centred over d (definitions above attn_half_tile).
__global__ void __launch_bounds__(256) item_attn_kernel(const float* __restrict__ feats, const float* __restrict__ in_wt,
                                                        const float* __restrict__ in_b, const float* __restrict__ out_wt,
                                                        const float* __restrict__ out_b, int M, int64_t n_rows,
                                                        float* __restrict__ rec) {
  __shared__ float x[4][ATT_TOKENS][D], q[4][ATT_TOKENS][D], k[4][ATT_TOKENS][D], v[4][ATT_TOKENS][D];
  __shared__ float Ssc[4][ATT_TOKENS][NH][ATT_TOKENS], P[4][ATT_TOKENS][NH][ATT_TOKENS], Lse[4][ATT_TOKENS][NH];
  __shared__ float red[4][2][2 * NH + 1];
  const int it = threadIdx.x >> 6, d = threadIdx.x & 63, nt = M - 1;
  const int64_t row = (int64_t)blockIdx.x * 4 + it;
  const bool valid = row < n_rows;
  const float scale = rsqrtf((float)DH);
  for (int b = 0; b < ATT_TOKENS; ++b) x[it][b][d] = (valid && b < nt) ? feats[(row * nt + b) * D + d] : 0.f;
  __syncthreads();
  {
    float aq[ATT_TOKENS], ak[ATT_TOKENS], av[ATT_TOKENS];
    for (int b = 0; b < ATT_TOKENS; ++b) { aq[b] = in_b[d]; ak[b] = in_b[D + d]; av[b] = in_b[2 * D + d]; }
    for (int kk = 0; kk < D; ++kk) {
      const float wq = in_wt[kk * 3 * D + d], wk = in_wt[kk * 3 * D + D + d], wv = in_wt[kk * 3 * D + 2 * D + d];
#pragma unroll
      for (int b = 0; b < ATT_TOKENS; ++b) {
        const float xv = x[it][b][kk];
        aq[b] = fmaf(wq, xv, aq[b]); ak[b] = fmaf(wk, xv, ak[b]); av[b] = fmaf(wv, xv, av[b]);
      }
    }
    for (int b = 0; b < ATT_TOKENS; ++b) { q[it][b][d] = aq[b] * scale; k[it][b][d] = ak[b]; v[it][b][d] = av[b]; }
  }
  __syncthreads();
  float U[ATT_TOKENS][NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
#pragma unroll
    for (int b = 0; b < ATT_TOKENS; ++b) U[b][h] = 0.f;
    for (int e = 0; e < DH; ++e) {
      const float w = out_wt[(h * DH + e) * D + d];
#pragma unroll
      for (int b = 0; b < ATT_TOKENS; ++b) U[b][h] = fmaf(w, v[it][b][h * DH + e], U[b][h]);
    }
  }
  for (int i = d; i < ATT_TOKENS * ATT_TOKENS * NH; i += 64) {     // item-item scores (already scaled through q)
    const int a = i / (ATT_TOKENS * NH), b = (i / NH) % ATT_TOKENS, h = i % NH;
    float acc = 0.f;
    for (int e = 0; e < DH; ++e) acc = fmaf(q[it][a][h * DH + e], k[it][b][h * DH + e], acc);
    Ssc[it][a][h][b] = acc;
  }
  __syncthreads();
  if (d < ATT_TOKENS * NH) {
    const int a = d >> 2, h = d & 3;
    float m = -INFINITY;
    for (int b = 0; b < nt; ++b) m = fmaxf(m, Ssc[it][a][h][b]);
    float sum = 0.f;
    for (int b = 0; b < nt; ++b) { const float e = expf(Ssc[it][a][h][b] - m); P[it][a][h][b] = e; sum += e; }
    for (int b = 0; b < ATT_TOKENS; ++b) P[it][a][h][b] = b < nt ? P[it][a][h][b] / sum : 0.f;
    Lse[it][a][h] = nt > 0 ? m + logf(sum) : 0.f;
  }
  __syncthreads();
  float* out = rec + row * ATT_ITEM;
  const int warp2 = d >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int a = 0; a < nt; ++a) {
    // vals: Nbar_h (4), U_h (4), C -> centre each over the item's 64 threads (two warps)
    float vals[2 * NH + 1];
    float c = x[it][a][d] + out_b[d];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float nb = 0.f;
#pragma unroll
      for (int b = 0; b < ATT_TOKENS; ++b) nb = fmaf(P[it][a][h][b], U[b][h], nb);
      vals[h] = nb; c += nb;
    }
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float uah = 0.f;
#pragma unroll
      for (int b = 0; b < ATT_TOKENS; ++b) uah = (b == a) ? U[b][h] : uah;
      vals[NH + h] = uah;
    }
    vals[2 * NH] = c;
#pragma unroll
    for (int i = 0; i < 2 * NH + 1; ++i) {
      float part = vals[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) red[it][warp2][i] = part;
    }
    __syncthreads();
    if (valid) {
      float* tr = out + a * ATT_TOK;
#pragma unroll
      for (int i = 0; i < 2 * NH + 1; ++i) vals[i] -= (red[it][0][i] + red[it][1][i]) * (1.f / D);
      tr[ATT_C + d] = vals[2 * NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) { tr[ATT_NB + h * D + d] = vals[h]; tr[ATT_U + h * D + d] = vals[NH + h]; }
      tr[ATT_Q + d] = q[it][a][d];
      tr[ATT_K + d] = k[it][a][d] * scale;
      if (d < NH) tr[ATT_L + d] = Lse[it][a][d];
    }
    __syncthreads();
  }
}

// attention: b1' = b1 + W1 ln_b (LayerNorm bias folded into layer 1); s1[k] = ln_w[k] / M for the W1 image
__global__ void attn_fold_ln_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ ln_w,
                                    const float* __restrict__ ln_b, int M, float* __restrict__ b1_out, float* __restrict__ s1_out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < H1) {
    float acc = b1[n];
    for (int k = 0; k < D; ++k) acc = fmaf(w1[(size_t)n * D + k], ln_b[k], acc);
    b1_out[n] = acc;
  }
  if (n < D) s1_out[n] = ln_w[n] / (float)M;
}

struct FastWeights {       // lives in h->fast_w; attention: followed by one UserAttn per SM
  uint8_t wimg[2 * Map<F_GATED>::WIMG];
  float bias[H1 + H2 + H3 + H3 + 4];
  float s1[D];             // attention: ln_w / M, the per-input scale folded into the layer-1 image
};

template <int FUS, int FMT, bool TK2>
static int launch_fused_tk(pxr_handle* h, const Params& p, int n_pairs, cudaStream_t st) {
  auto kern = score_fused_kernel<FUS, FMT, TK2>;
  const int slot = 2 * FUS + FMT + (TK2 ? 16 : 0);
  if (!(h->tc_attr_set & (1u << slot))) {
    PXR_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Map<FUS>::SMEM));
    h->tc_attr_set |= (1u << slot);
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
template <int FUS, int FMT>
static int launch_fused(pxr_handle* h, const Params& p, int n_pairs, cudaStream_t st) {
  return p.rows_per_split < 8192 ? launch_fused_tk<FUS, FMT, true>(h, p, n_pairs, st) : launch_fused_tk<FUS, FMT, false>(h, p, n_pairs, st);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool pxr_tc_supported(const pxr_handle* h) {
  const pxr_config& c = h->cfg;
  const bool fusion_ok = c.fusion == PXR_FUSION_GATED || c.fusion == PXR_FUSION_CONCAT ||
                         (c.fusion == PXR_FUSION_ATTENTION && c.num_heads == tc::NH);
  // concat: layer 1 is applied as per-user / per-item partials, so the fused kernel does not depend on embedding_dim
  // (item side: 3xTF32 GEMMs of items_tc.cu, which need single-layer projections and 16-byte aligned rows)
  const bool dim_ok = c.embedding_dim == tc::D ||
                      (c.fusion == PXR_FUSION_CONCAT && c.embedding_dim % 16 == 0 && c.embedding_dim <= 512 && c.projection_hidden == 0 &&
                       c.vision_dim % 4 == 0 && c.language_dim % 4 == 0 && c.num_numerical <= 32);
  return fusion_ok && dim_ok && c.n_hidden == 3 &&
         c.hidden[0] == tc::H1 && c.hidden[1] == tc::H2 && c.hidden[2] == tc::H3 && c.activation == PXR_ACT_RELU &&
         h->M >= 4 && h->M <= 6 && h->n_sm >= 2;
}

bool pxr_tc_can_run(const pxr_handle* h, int32_t k) {
  return h->fast_ok && k <= tc::KCAP && h->n_rows > 0 && h->item_base + h->n_rows < (int64_t)tc::IDX_MASK;
}

size_t pxr_tc_weight_bytes(const pxr_handle* h) {
  return pxr_align_up(sizeof(tc::FastWeights), 256) +
         (h->cfg.fusion == PXR_FUSION_ATTENTION ? (size_t)h->n_sm * sizeof(tc::UserAttn) : 0);
}

static int tc_fmt(const pxr_handle* h) { return h->cfg.precision == PXR_PRECISION_FP16 ? tc::FMT_FP16 : tc::FMT_BF16; }

int pxr_tc_prepare_weights(pxr_handle* h, cudaStream_t st) {
  if (!h->fast_w) PXR_CUDA(h, cudaMalloc(&h->fast_w, pxr_tc_weight_bytes(h)));
  tc::FastWeights* fw = reinterpret_cast<tc::FastWeights*>(h->fast_w);
  const bool gated = h->cfg.fusion != PXR_FUSION_CONCAT;    // layer 1 on the tensor pipe
  const bool attn = h->cfg.fusion == PXR_FUSION_ATTENTION;
  float* b = fw->bias;
  if (attn) {     // LayerNorm affine folded into layer 1 (see attn_half_tile)
    tc::attn_fold_ln_kernel<<<(tc::H1 + 127) / 128, 128, 0, st>>>(h->mlp[0].w, h->mlp[0].b, h->ln_w, h->ln_b, h->M, b, fw->s1);
    h->launches++;
  } else {
    PXR_CUDA(h, cudaMemcpyAsync(b, h->mlp[0].b, sizeof(float) * tc::H1, cudaMemcpyDeviceToDevice, st));
  }
  tc::build_wimg_kernel<<<296, 256, 0, st>>>(gated ? h->mlp[0].w : nullptr, h->mlp[0].k, attn ? fw->s1 : nullptr, h->mlp[1].w,
                                             h->mlp[2].w, fw->wimg, gated ? 1 : 0, tc_fmt(h));
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1, h->mlp[1].b, sizeof(float) * tc::H2, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2, h->mlp[2].b, sizeof(float) * tc::H3, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2 + tc::H3, h->out.w, sizeof(float) * tc::H3, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2 + 2 * tc::H3, h->out.b, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (attn) {     // the attention kernel reads these through the kernel-parameter constant bank: keep a host copy
    PXR_CUDA(h, cudaMemcpyAsync(h->tc_bias_host, b, sizeof(float) * (tc::H1 + tc::H2 + 2 * tc::H3 + 1), cudaMemcpyDeviceToHost, st));
    PXR_CUDA(h, cudaStreamSynchronize(st));
  }
  return PXR_OK;
}

size_t pxr_tc_item_bytes(const pxr_handle* h, int64_t n_rows) {
  const size_t rows = (size_t)((n_rows + 31) / 32 * 32);
  if (h->cfg.fusion == PXR_FUSION_GATED) return pxr_align_up(rows * 8 * sizeof(float), 256);
  if (h->cfg.fusion == PXR_FUSION_ATTENTION) return pxr_align_up(rows * tc::ATT_ITEM * sizeof(float), 256);
  return pxr_align_up(rows * tc::H1 * sizeof(uint16_t), 256);
}

int pxr_tc_prepare_items(pxr_handle* h, int64_t n_rows, void* ws, cudaStream_t st) {
  if (n_rows == 0) return PXR_OK;
  if (h->cfg.fusion == PXR_FUSION_GATED && h->tc_items_img[3] && h->path == PXR_PATH_TCGEN05) {
    return pxr_launch_item_logit_tc(h, n_rows, (float*)ws, st);                  // 3xTF32 GEMM on the tensor pipe (N = 6 padded to 16)
  } else if (h->cfg.fusion == PXR_FUSION_GATED) {
    const int wpb = 8;
    tc::item_logit_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(h->item_feats, h->gate.w, h->gate.b, h->M,
                                                                                     n_rows, (float*)ws);
  } else if (h->cfg.fusion == PXR_FUSION_ATTENTION) {
    tc::item_attn_kernel<<<(unsigned)((n_rows + 3) / 4), 256, 0, st>>>(h->item_feats, h->attn_in.wt, h->attn_in.b, h->attn_out.wt,
                                                                       h->attn_out.b, h->M, n_rows, (float*)ws);
  } else if (h->tc_items_img[2] && h->path == PXR_PATH_TCGEN05) {
    return pxr_launch_item_pi_tc(h, n_rows, (uint16_t*)ws, tc_fmt(h), st);      // 3xTF32 GEMM on the tensor pipe
  } else {
    const int FD = (h->M - 1) * h->cfg.embedding_dim;
    const size_t smem = (size_t)32 * (FD + tc::H1) * sizeof(float);
    if (smem > (size_t)h->max_smem_optin) PXR_FAIL(h, PXR_ERR_INVALID, "concat item partial: embedding_dim %d needs the tensor-pipe item path", h->cfg.embedding_dim);
    if (!(h->tc_attr_set & 256u)) {
      PXR_CUDA(h, cudaFuncSetAttribute(tc::item_pi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      h->tc_attr_set |= 256u;
    }
    tc::item_pi_kernel<<<(unsigned)((n_rows + 31) / 32), PXR_SIMT_THREADS, smem, st>>>(h->item_feats, h->mlp[0].wt, h->mlp[0].b,
                                                                                       h->M, h->cfg.embedding_dim, n_rows, (uint16_t*)ws, tc_fmt(h));
  }
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

struct TcPlan { int n_groups, S, rows_per_split, n_units, n_pairs; };

static TcPlan tc_plan(const pxr_handle* h, int64_t n_users) {
  TcPlan pl;
  const int max_pairs = h->n_sm / 2;
  pl.n_groups = (int)((n_users + 2 * tc::TU - 1) / (2 * tc::TU));
  const int64_t max_tiles = (h->n_rows + tc::TI - 1) / tc::TI;
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
  rps = (rps + tc::TI - 1) / tc::TI * tc::TI;
  pl.rows_per_split = (int)rps;
  pl.S = (int)((h->n_rows + rps - 1) / rps);
  pl.n_units = pl.n_groups * pl.S;
  pl.n_pairs = std::min(max_pairs, pl.n_units);
  return pl;
}

// workspace: [S > 1: per-split partial lists][exact mode: the merged 64-slot lists + the re-score pair arrays]
size_t pxr_tc_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k) {
  if (n_users <= 0 || h->n_rows <= 0) return 256;
  const TcPlan pl = tc_plan(h, n_users);
  const int kk = h->rescore ? tc::KCAP : k;
  size_t b = 256;
  if (pl.S > 1) b += pxr_align_up((size_t)pl.S * n_users * kk * 8, 256);
  if (h->rescore) b += pxr_align_up((size_t)n_users * kk * 8, 256) + pxr_rescore_bytes(n_users);
  return b;
}

int pxr_tc_score_topk(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                      const int64_t* seen_indptr, const int32_t* seen_idx, int32_t k, float* out_scores, int32_t* out_idx,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!pxr_tc_can_run(h, k)) PXR_FAIL(h, PXR_ERR_INVALID, "tcgen05 path cannot run this call (top_k <= %d, 28-bit item index)", tc::KCAP);
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
  if (attn) {
    p.attn_rec = (const float*)h->item_fast;
    p.attn_in_wt = h->attn_in.wt; p.attn_in_b = h->attn_in.b; p.attn_out_wt = h->attn_out.wt; p.attn_out_b = h->attn_out.b;
    p.ln_w = h->ln_w; p.ln_b = h->ln_b;
    p.user_scratch = reinterpret_cast<tc::UserAttn*>((char*)h->fast_w + pxr_align_up(sizeof(tc::FastWeights), 256));
    memcpy(p.bias_c, h->tc_bias_host, sizeof(float) * (tc::H1 + tc::H2 + 2 * tc::H3 + 1));
  }
  p.w1u_t = h->mlp[0].wt;                   // [k][512]: rows 0..63 are the user columns of W1
  p.user_emb = user_embedding; p.user_idx = user_idx; p.seen_indptr = seen_indptr; p.seen_idx = seen_idx;
  p.item_missing = h->item_missing;
  p.n_users = n_users; p.n_rows = h->n_rows; p.item_base = h->item_base;
  p.Dm = h->cfg.embedding_dim;
  // exact mode: the kernel keeps its full 64-slot list per user (admission threshold = 64th best); the lists are
  // re-scored in fp32 and re-ranked afterwards (pxr_launch_rescore)
  const bool exact = h->rescore;
  const int32_t kk = exact ? tc::KCAP : k;
  p.M = h->M; p.K = kk; p.S = pl.S; p.rows_per_split = pl.rows_per_split; p.n_units = pl.n_units;
  p.final_act = h->cfg.final_activation;
  if (ws_bytes < pxr_tc_topk_bytes(h, n_users, k)) PXR_FAIL(h, PXR_ERR_WORKSPACE, "tcgen05 top-K workspace too small");
  char* wp = (char*)ws;
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
  int rc;
  const int fmt = tc_fmt(h);
  if (gated) rc = fmt == tc::FMT_BF16 ? tc::launch_fused<tc::F_GATED, tc::FMT_BF16>(h, p, pl.n_pairs, st)
                                      : tc::launch_fused<tc::F_GATED, tc::FMT_FP16>(h, p, pl.n_pairs, st);
  else if (attn) rc = fmt == tc::FMT_BF16 ? tc::launch_fused<tc::F_ATTN, tc::FMT_BF16>(h, p, pl.n_pairs, st)
                                          : tc::launch_fused<tc::F_ATTN, tc::FMT_FP16>(h, p, pl.n_pairs, st);
  else rc = fmt == tc::FMT_BF16 ? tc::launch_fused<tc::F_CONCAT, tc::FMT_BF16>(h, p, pl.n_pairs, st)
                                : tc::launch_fused<tc::F_CONCAT, tc::FMT_FP16>(h, p, pl.n_pairs, st);
  if (rc) return rc;
  if (pl.S > 1) {
    rc = pxr_launch_merge(part_s, part_i, pl.S, n_users, kk, list_s, list_i, st);
    h->launches++;
    if (rc) PXR_FAIL(h, rc, "top-K merge of %d item splits failed", pl.S);
  }
  if (exact) return pxr_launch_rescore(h, user_embedding, user_idx, n_users, list_i, k, out_scores, out_idx, wp, st);
  return PXR_OK;
}
