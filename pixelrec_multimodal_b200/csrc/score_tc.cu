// Fused pair-scoring + top-K kernels for sm_100a (tcgen05 / TMEM / TMA engine).
//
// Replace, for a block of users against a shard of the catalogue, the loop of reference
// src/inference/recommender.py:97-106 around MultimodalRecommender.forward
// (src/models/multimodal.py:528-610) for fusion_type 'gated' (src/models/layers.py:195-225) and
// 'concatenate' (multimodal.py:583-584) with the default prediction MLP [512, 256, 128] -> 1
// (multimodal.py:366-386).  One kernel template, two front ends.
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
//     of the leader CTA) and owns TMEM/TMA setup, 5 top-K, 8-11 / 12-15 two epilogue groups.
//     All hand-offs are mbarriers; tcgen05.commit multicasts completion to both CTAs.
#include <algorithm>

#include <cuda_fp16.h>

#include "pxr_common.cuh"
#include "tc_ptx.cuh"

namespace tc {

constexpr int D = 64, H1 = 512, H2 = 256, H3 = 128;
constexpr int TU = 8, TI = 16;            // users x items per CTA tile (128 rows)
constexpr int KCAP = 64;                  // slots of the per-user sorted list (K <= KCAP)
constexpr int QCAP = 512;                 // candidate queue entries
constexpr int THREADS = 512;
constexpr uint32_t IDX_MASK = 0x0FFFFFFFu;   // 28-bit item index inside a queue / list key
constexpr int FMT_BF16 = 0, FMT_FP16 = 1;
#ifndef PXR_TOPK_IDLE_NS
#define PXR_TOPK_IDLE_NS 200
#endif

enum {
  BAR_W = 0, BAR_A_FULL, BAR_A_EMPTY, BAR_D1_FULL0, BAR_D1_FULL1, BAR_D1_FULL2, BAR_D1_FULL3, BAR_H1_FULL0, BAR_H1_FULL1,
  BAR_H1_FULL2, BAR_H1_FULL3, BAR_H1_EMPTY0, BAR_H1_EMPTY1, BAR_H1_EMPTY2, BAR_H1_EMPTY3, BAR_PI_FULL0, BAR_PI_FULL1,
  BAR_PI_EMPTY0, BAR_PI_EMPTY1, BAR_D2_FULL, BAR_H2_FULL0, BAR_H2_FULL1, BAR_D3_FULL, BAR_D3_EMPTY, BAR_UNIT_DONE,
  BAR_UNIT_RESET, N_BARS
};

struct Misc {
  float b1[H1]; float b2[H2]; float b3[H3]; float w4[H3];
  float eu[TU][D];
  float lu[TU][8];
  unsigned long long list[TU][KCAP];
  unsigned long long queue[QCAP];
  float thr[TU];
  uint32_t seen_mask[4][TU];
  uint32_t q_tail, q_head;
  uint32_t tmem_base;
  float b4;
  unsigned long long bars[N_BARS];
};

// shared / tensor memory maps per fusion type
template <bool GATED>
struct Map {
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
  static constexpr uint32_t SMEM = OFF_MISC + (uint32_t)sizeof(Misc) + 1024u;   // + alignment slack
  // tensor memory (columns)
  static constexpr uint32_t TM_D3 = 128, TM_D2 = 256;
  // H1 chunk buffer b.  gated: 64-col fp32 accumulators packed in place, buffers 2,3 share the D3 columns
  // (D3 is only live at the tile boundary).  concat: four 32-col packed buffers.
  __host__ __device__ static constexpr uint32_t h1buf(int b) {
    return GATED ? (b < 2 ? 64u * b : TM_D3 + 64u * (b - 2)) : 32u * b;
  }
};
static_assert(Map<true>::SMEM <= 232448 && Map<false>::SMEM <= 232448, "shared memory budget");
static_assert(Map<false>::OFF_MISC % 16 == 0 && Map<true>::OFF_MISC % 16 == 0, "alignment");

struct Params {
  const uint8_t* wimg;          // [2][WIMG] pre-swizzled 16-bit operand images (rank 0, rank 1)
  const float* bias;            // b1[512] b2[256] b3[128] w4[128] b4   (b1 unused by concat: folded into Pi)
  const float* gate_w;          // gated: (M, M*D) fp32 row-major; the user part is the first D of each row
  const float* item_feats;      // gated: [rows][M-1][D] fp32 projected item-side modality vectors
  const float* item_logit;      // gated: [rows][8] fp32 item part of the gate logits (+ gate bias)
  const uint16_t* item_pi;      // concat: [rows][512] 16-bit item partial of layer 1 (+ b1)
  const float* w1u_t;           // concat: [64][512] fp32, user columns of W1 transposed
  const float* user_emb;        // (n_users_total, D) fp32 table
  const int64_t* user_idx;      // (n_users,)
  const int64_t* seen_indptr;   // (n_users + 1,) or NULL
  const int32_t* seen_idx;      // global item indices, ascending per user
  float* out_scores;            // [S][n_users][K]
  int32_t* out_idx;
  int64_t n_users, n_rows, item_base;
  int M, K, S, rows_per_split, n_units, final_act;
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
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool GATED, int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
score_fused_kernel(const __grid_constant__ Params p) {
  using MP = Map<GATED>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);
  Misc& ms = *reinterpret_cast<Misc*>(sm + MP::OFF_MISC);
  const uint32_t bar0 = ptx::smem_u32(&ms.bars[0]);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  // ------------------------------------------------------------------ setup
  for (int i = threadIdx.x; i < H1 + H2 + H3 + H3; i += THREADS) ms.b1[i] = p.bias[i];   // b1,b2,b3,w4 are contiguous
  if (threadIdx.x == 0) {
    ms.b4 = p.bias[H1 + H2 + H3 + H3];
    ms.q_tail = 0; ms.q_head = 0;
    ptx::mbar_init(BAR(BAR_W), 1);
    ptx::mbar_init(BAR(BAR_A_FULL), 8);
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
    ptx::mbar_init(BAR(BAR_UNIT_RESET), 1);
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < TU * KCAP; i += THREADS) (&ms.list[0][0])[i] = 0ull;
  for (int i = threadIdx.x; i < QCAP; i += THREADS) ms.queue[i] = 0ull;
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

  if (warp < 4) {
    // =============================================================== front end
    const int tid = threadIdx.x;            // 0..127
    const int Mm = p.M;
    int T = 0;
    for (int w = pair; w < p.n_units; w += n_pairs) {
      const Unit un = decode_unit(p, w);
      const int64_t ubase = ((int64_t)un.g * 2 + rank) * TU;     // first user ordinal of this CTA's group
      if (!GATED && T > 0) {
        // Pu is read by the layer-1 producers of the previous unit's last tile: wait until they are done with it
        ptx::mbar_wait(BAR(BAR_PI_EMPTY0 + ((T - 1) & 1)), ((T - 1) >> 1) & 1);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");              // previous unit's readers of eu/lu are done
      {
        const int u = tid >> 4, d4 = (tid & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ubase + u < p.n_users) v = *reinterpret_cast<const float4*>(p.user_emb + p.user_idx[ubase + u] * D + d4);
        *reinterpret_cast<float4*>(&ms.eu[u][d4]) = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (GATED) {
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
        // per-user partial of layer 1: Pu[u][n] = sum_k W1[n][k < 64] Eu[u][k]   (SURVEY.md A3), fp32
        float acc[TU][4];
#pragma unroll
        for (int u = 0; u < TU; ++u) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f; }
        const float* wcol = p.w1u_t + 4 * tid;
#pragma unroll 4
        for (int k = 0; k < D; ++k) {
          const float4 wv = *reinterpret_cast<const float4*>(wcol + (size_t)k * H1);
#pragma unroll
          for (int u = 0; u < TU; ++u) {
            const float e = ms.eu[u][k];
            acc[u][0] = fmaf(e, wv.x, acc[u][0]); acc[u][1] = fmaf(e, wv.y, acc[u][1]);
            acc[u][2] = fmaf(e, wv.z, acc[u][2]); acc[u][3] = fmaf(e, wv.w, acc[u][3]);
          }
        }
#pragma unroll
        for (int u = 0; u < TU; ++u)
          *reinterpret_cast<float4*>(sm + MP::OFF_PU + u * MP::PU_STRIDE + 16 * tid) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
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
        if (GATED) {
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
  } else if (warp == 4) {
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
  } else if (warp == 5) {
    // =============================================================== top-K warp
    uint32_t head = 0, done_ph = 0;
    for (int w = pair; w < p.n_units; w += n_pairs) {
      const Unit un = decode_unit(p, w);
      if (un.ntiles == 0) continue;
      const int64_t ubase = ((int64_t)un.g * 2 + rank) * TU;
      bool finished = false;
      while (true) {
        unsigned long long e = 0ull;
        if (lane == 0) e = *reinterpret_cast<volatile unsigned long long*>(&ms.queue[head % QCAP]);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e != 0ull) {
          __syncwarp();
          if (lane == 0) {
            *reinterpret_cast<volatile unsigned long long*>(&ms.queue[head % QCAP]) = 0ull;
            *reinterpret_cast<volatile uint32_t*>(&ms.q_head) = head + 1;
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
          if (lane == 0) tail = *reinterpret_cast<volatile uint32_t*>(&ms.q_tail);
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
      for (int u = 0; u < TU; ++u) {
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
  } else if (warp >= 8) {
    // =============================================================== epilogue groups
    const int grp = (warp - 8) >> 2;                 // 0: even layer-1 chunks, first half of layer 2, layer 3
    const int q = warp & 3;                          // TMEM lane quarter this warp may touch
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const int r = q * 32 + lane;                     // row of the tile this thread owns
    const int ru = r >> 4, rj = r & 15;              // user slot / item slot of the row
    uint32_t d1ph = 0, reset_ph = 0;                 // d1ph: phase bits of chunk buffers grp (bit 0) and grp + 2 (bit 1)
    uint32_t h1use0 = 0, h1use1 = 0;                 // concat: uses so far of chunk buffers grp and grp + 2
    int T = 0;
    Unit prev; prev.ntiles = 0; prev.row_lo = prev.row_hi = 0; prev.g = prev.s = 0;
    int prev_t = 0; int64_t prev_ubase = 0; bool have_prev = false;

    auto do_e2 = [&](int Tp) {                       // H2 half `grp` of tile Tp
      ptx::mbar_wait(BAR(BAR_D2_FULL), Tp & 1);
      ptx::tc_fence_after();
      const uint32_t c0 = MP::TM_D2 + grp * 128;
      epi_pack64<FMT>(tl + c0, tl + c0, ms.b2 + grp * 128);
      epi_pack64<FMT>(tl + c0 + 64, tl + c0 + 32, ms.b2 + grp * 128 + 64);
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(BAR(BAR_H2_FULL0 + grp), 0);
    };
    auto do_e3 = [&](int Tp, const Unit& un, int t, int64_t ubase, bool last_of_unit) {
      ptx::mbar_wait(BAR(BAR_D3_FULL), Tp & 1);
      ptx::tc_fence_after();
      float z = ms.b4;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld32(tl + MP::TM_D3 + h * 64, v0);
        ptx::tmem_ld32(tl + MP::TM_D3 + h * 64 + 32, v1);
        ptx::tc_wait_ld();
        const float4* b3v = reinterpret_cast<const float4*>(ms.b3 + h * 64);
        const float4* w4v = reinterpret_cast<const float4*>(ms.w4 + h * 64);
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
      const float y = pxr_apply_final(z, p.final_act);
      const int64_t row = un.row_lo + (int64_t)t * TI + rj;
      const bool ok = row < un.row_hi && (ubase + ru) < p.n_users && !((ms.seen_mask[Tp & 3][ru] >> rj) & 1u);
      if (ok && y >= *reinterpret_cast<volatile float*>(&ms.thr[ru])) {
        const uint32_t gidx = (uint32_t)(p.item_base + row);
        const unsigned long long e = ((unsigned long long)pxr_ord(y) << 32) | ((unsigned long long)ru << 28) |
                                     (unsigned long long)(IDX_MASK - gidx);
        const uint32_t slot = atomicAdd(&ms.q_tail, 1u);
        while (slot - *reinterpret_cast<volatile uint32_t*>(&ms.q_head) >= (uint32_t)QCAP) __nanosleep(64);
        *reinterpret_cast<volatile unsigned long long*>(&ms.queue[slot % QCAP]) = e;
      }
      if (last_of_unit) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_local(BAR(BAR_UNIT_DONE));
      }
    };
    auto prev_e3 = [&]() {
      const bool last = (prev_t == prev.ntiles - 1);
      if (prev_t == 0 && (T - 1) > 0) { ptx::mbar_wait(BAR(BAR_UNIT_RESET), reset_ph); reset_ph ^= 1; }
      do_e3(T - 1, prev, prev_t, prev_ubase, last);
    };
    // layer-1 chunk ci (0..3) of this group for the current tile: chunk c = 2 ci + grp in buffer c % 4
    auto do_l1 = [&](int ci) {
      const int c = 2 * ci + grp;
      const int b = c & 3;                           // grp (ci even) or grp + 2 (ci odd)
      if (GATED) {
        ptx::mbar_wait(BAR(BAR_D1_FULL0 + b), (d1ph >> (ci & 1)) & 1u); d1ph ^= 1u << (ci & 1);
        ptx::tc_fence_after();
        epi_pack64<FMT>(tl + MP::h1buf(b), tl + MP::h1buf(b), ms.b1 + c * 64);
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
        prev = un; prev_t = t; prev_ubase = ubase; have_prev = true;
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
__global__ void build_wimg_kernel(const float* __restrict__ w1, int k1, const float* __restrict__ w2,
                                  const float* __restrict__ w3, uint8_t* __restrict__ img, int gated, int fmt) {
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
    if (which == 1) v = w1[(size_t)(blk * 64 + rank * 32 + nl) * k1 + kk];          // N-chunk blk, rows [32 rank, +32)
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
                                                                    const float* __restrict__ b1, int M, int64_t n_rows,
                                                                    uint16_t* __restrict__ out, int fmt) {
  extern __shared__ __align__(16) float smem_pi[];
  const int FD = (M - 1) * D;
  float* in = smem_pi;                       // [32][FD]
  float* res = smem_pi + 32 * FD;            // [32][512]
  const int64_t row0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * FD; i += PXR_SIMT_THREADS) {
    const int64_t row = row0 + i / FD;
    in[i] = row < n_rows ? feats[row * FD + i % FD] : 0.f;
  }
  __syncthreads();
  linear_rows<32>(in, FD, FD, w1t + (size_t)D * H1, b1, H1, res, H1, -1);
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * H1; i += PXR_SIMT_THREADS) {
    const int64_t row = row0 + i / H1;
    if (row < n_rows) out[row * H1 + i % H1] = to16(res[i], fmt);
  }
}

struct FastWeights {       // lives in h->fast_w
  uint8_t wimg[2 * Map<true>::WIMG];
  float bias[H1 + H2 + H3 + H3 + 4];
};

template <bool GATED, int FMT>
static int launch_fused(pxr_handle* h, const Params& p, int n_pairs, cudaStream_t st) {
  auto kern = score_fused_kernel<GATED, FMT>;
  const int slot = (GATED ? 0 : 2) + FMT;
  if (!(h->tc_attr_set & (1u << slot))) {
    PXR_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Map<GATED>::SMEM));
    h->tc_attr_set |= (1u << slot);
  }
  pxr_prof_begin(h, st);
  kern<<<2 * n_pairs, THREADS, Map<GATED>::SMEM, st>>>(p);
  pxr_prof_end(h, st);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool pxr_tc_supported(const pxr_handle* h) {
  const pxr_config& c = h->cfg;
  return (c.fusion == PXR_FUSION_GATED || c.fusion == PXR_FUSION_CONCAT) && c.embedding_dim == tc::D && c.n_hidden == 3 &&
         c.hidden[0] == tc::H1 && c.hidden[1] == tc::H2 && c.hidden[2] == tc::H3 && c.activation == PXR_ACT_RELU &&
         h->M >= 4 && h->M <= 6 && h->n_sm >= 2;
}

bool pxr_tc_can_run(const pxr_handle* h, int32_t k) {
  return h->fast_ok && k <= tc::KCAP && h->n_rows > 0 && h->item_base + h->n_rows < (int64_t)tc::IDX_MASK;
}

size_t pxr_tc_weight_bytes(const pxr_handle* h) { (void)h; return sizeof(tc::FastWeights); }

static int tc_fmt(const pxr_handle* h) { return h->cfg.precision == PXR_PRECISION_FP16 ? tc::FMT_FP16 : tc::FMT_BF16; }

int pxr_tc_prepare_weights(pxr_handle* h, cudaStream_t st) {
  if (!h->fast_w) PXR_CUDA(h, cudaMalloc(&h->fast_w, sizeof(tc::FastWeights)));
  tc::FastWeights* fw = reinterpret_cast<tc::FastWeights*>(h->fast_w);
  const bool gated = h->cfg.fusion == PXR_FUSION_GATED;
  tc::build_wimg_kernel<<<296, 256, 0, st>>>(gated ? h->mlp[0].w : nullptr, h->mlp[0].k, h->mlp[1].w, h->mlp[2].w, fw->wimg,
                                             gated ? 1 : 0, tc_fmt(h));
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  float* b = fw->bias;
  PXR_CUDA(h, cudaMemcpyAsync(b, h->mlp[0].b, sizeof(float) * tc::H1, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1, h->mlp[1].b, sizeof(float) * tc::H2, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2, h->mlp[2].b, sizeof(float) * tc::H3, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2 + tc::H3, h->out.w, sizeof(float) * tc::H3, cudaMemcpyDeviceToDevice, st));
  PXR_CUDA(h, cudaMemcpyAsync(b + tc::H1 + tc::H2 + 2 * tc::H3, h->out.b, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return PXR_OK;
}

size_t pxr_tc_item_bytes(const pxr_handle* h, int64_t n_rows) {
  const size_t rows = (size_t)((n_rows + 31) / 32 * 32);
  if (h->cfg.fusion == PXR_FUSION_GATED) return pxr_align_up(rows * 8 * sizeof(float), 256);
  return pxr_align_up(rows * tc::H1 * sizeof(uint16_t), 256);
}

int pxr_tc_prepare_items(pxr_handle* h, int64_t n_rows, void* ws, cudaStream_t st) {
  if (n_rows == 0) return PXR_OK;
  if (h->cfg.fusion == PXR_FUSION_GATED) {
    const int wpb = 8;
    tc::item_logit_kernel<<<(unsigned)((n_rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(h->item_feats, h->gate.w, h->gate.b, h->M,
                                                                                     n_rows, (float*)ws);
  } else {
    const int FD = (h->M - 1) * tc::D;
    const size_t smem = (size_t)32 * (FD + tc::H1) * sizeof(float);
    if (!(h->tc_attr_set & 16u)) {
      PXR_CUDA(h, cudaFuncSetAttribute(tc::item_pi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      h->tc_attr_set |= 16u;
    }
    tc::item_pi_kernel<<<(unsigned)((n_rows + 31) / 32), PXR_SIMT_THREADS, smem, st>>>(h->item_feats, h->mlp[0].wt, h->mlp[0].b,
                                                                                       h->M, n_rows, (uint16_t*)ws, tc_fmt(h));
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

size_t pxr_tc_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k) {
  if (n_users <= 0 || h->n_rows <= 0) return 256;
  const TcPlan pl = tc_plan(h, n_users);
  if (pl.S == 1) return 256;
  return pxr_align_up((size_t)pl.S * n_users * k * 8, 256) + 256;
}

int pxr_tc_score_topk(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                      const int64_t* seen_indptr, const int32_t* seen_idx, int32_t k, float* out_scores, int32_t* out_idx,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!pxr_tc_can_run(h, k)) PXR_FAIL(h, PXR_ERR_INVALID, "tcgen05 path cannot run this call (top_k <= %d, 28-bit item index)", tc::KCAP);
  const TcPlan pl = tc_plan(h, n_users);
  tc::FastWeights* fw = reinterpret_cast<tc::FastWeights*>(h->fast_w);
  const bool gated = h->cfg.fusion == PXR_FUSION_GATED;
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.wimg = fw->wimg; p.bias = fw->bias; p.gate_w = h->gate.w;
  p.item_feats = h->item_feats;
  p.item_logit = gated ? (const float*)h->item_fast : nullptr;
  p.item_pi = gated ? nullptr : (const uint16_t*)h->item_fast;
  p.w1u_t = h->mlp[0].wt;                   // [k][512]: rows 0..63 are the user columns of W1
  p.user_emb = user_embedding; p.user_idx = user_idx; p.seen_indptr = seen_indptr; p.seen_idx = seen_idx;
  p.n_users = n_users; p.n_rows = h->n_rows; p.item_base = h->item_base;
  p.M = h->M; p.K = k; p.S = pl.S; p.rows_per_split = pl.rows_per_split; p.n_units = pl.n_units;
  p.final_act = h->cfg.final_activation;
  float* part_s = out_scores; int32_t* part_i = out_idx;
  if (pl.S > 1) {
    const size_t need = (size_t)pl.S * n_users * k;
    if (ws_bytes < need * 8) PXR_FAIL(h, PXR_ERR_WORKSPACE, "tcgen05 top-K workspace too small");
    part_s = (float*)ws; part_i = (int32_t*)((char*)ws + need * 4);
  }
  p.out_scores = part_s; p.out_idx = part_i;
  int rc;
  const int fmt = tc_fmt(h);
  if (gated) rc = fmt == tc::FMT_BF16 ? tc::launch_fused<true, tc::FMT_BF16>(h, p, pl.n_pairs, st)
                                      : tc::launch_fused<true, tc::FMT_FP16>(h, p, pl.n_pairs, st);
  else rc = fmt == tc::FMT_BF16 ? tc::launch_fused<false, tc::FMT_BF16>(h, p, pl.n_pairs, st)
                                : tc::launch_fused<false, tc::FMT_FP16>(h, p, pl.n_pairs, st);
  if (rc) return rc;
  if (pl.S > 1) {
    rc = pxr_launch_merge(part_s, part_i, pl.S, n_users, k, out_scores, out_idx, st);
    h->launches++;
    if (rc) PXR_FAIL(h, rc, "top-K merge of %d item splits failed", pl.S);
  }
  return PXR_OK;
}
