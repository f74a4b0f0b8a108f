// Shared declarations for libpxr.so (see include/pxr.h for the ABI).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "pxr.h"

#define PXR_MAX_MODALITIES 6

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
struct PxrLinear {            // one (BN-folded) Linear of the prediction MLP or a projection
  int n = 0, k = 0;           // out features, in features
  float* w = nullptr;         // [n][k] row-major fp32 (folded)
  float* wt = nullptr;        // [k][n] transposed fp32 (folded) for linear_rows
  float* b = nullptr;         // [n]
};

struct pxr_handle {
  pxr_config cfg;
  int device = 0;
  int n_sm = 0;
  int M = 0;                  // number of modalities (tokens), reference multimodal.py:336-342
  int max_smem_optin = 0;
  char err[512];
  int64_t launches = 0;
  int path = PXR_PATH_SIMT;   // resolved path
  bool weights_loaded = false;
  bool items_ready = false;   // pxr_precompute_items ran since the last pxr_load_weights

  // weight arena (device)
  void* arena = nullptr;
  size_t arena_bytes = 0, arena_used = 0;
  float* tag_emb = nullptr;   // copy of tag_embedding.weight
  PxrLinear proj[3][2];       // [vision, language, numerical][layer 0 / layer 1]
  bool has_mod[3] = {false, false, false};
  PxrLinear gate;             // (M, M*D)
  PxrLinear attn_in, attn_out;
  float* ln_w = nullptr; float* ln_b = nullptr;
  PxrLinear mlp[PXR_MAX_HIDDEN];
  PxrLinear out;              // (1, H_L)

  // fast (tcgen05) path images
  bool fast_ok = false;
  bool records_only = false;  // pxr_set_records_only: keep only the fp32 item records (a handle used for re-scoring / explicit pairs)
  int small_batch = -1;       // small-batch tile shape of the fused gated kernel: -1 auto (cost model), 0 off, 1 whenever possible (pxr_set_small_batch)
  bool rescore = true;        // exact mode of the fused path: fp32 re-score + re-rank of the 64-slot lists (pxr_set_rescore)
  uint64_t tc_attr_set = 0;   // bit per kernel whose max-dynamic-smem attribute has been set
  uint64_t tc_attr_fused[3] = {0, 0, 0};   // the same for the instantiations of the fused scoring kernel (score_tc.cu)
  void* fast_w = nullptr;     // bf16 swizzled operand images + fp32 vectors (see score_tc.cu)
  float tc_bias_host[1028];   // attention fast path: host copy of b1' b2 b3 w4 b4 (passed as kernel parameters)
  void* tc_items_w = nullptr;       // item precompute on the tensor pipe: hi / lo tf32 weight chunk images (items_tc.cu)
  uint8_t* tc_items_img[4] = {nullptr, nullptr, nullptr, nullptr};   // vision, language projections; concat layer-1 item columns; gate logits
  float* tc_gate_bias = nullptr;    // gate bias padded to 16

  // live timing of the dominant kernel (pxr_profile_*)
  bool profile = false;
  cudaEvent_t* prof_ev = nullptr;   // 2 * PXR_PROFILE_SLOTS events, created lazily
  int prof_n = 0;

  // precomputed item records (caller workspace)
  float* item_feats = nullptr;   // [n_rows][M-1][D] fp32
  void* item_fast = nullptr;     // fast-path per-item records
  int64_t n_rows = 0;
  int64_t item_base = 0;
  const uint8_t* item_missing = nullptr;   // caller-owned [n_rows] flags (pxr_set_missing_items)
};

#define PXR_FAIL(h, code, ...)                                   \
  do {                                                           \
    snprintf((h)->err, sizeof((h)->err), __VA_ARGS__);           \
    return (code);                                               \
  } while (0)

#define PXR_CUDA(h, expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      snprintf((h)->err, sizeof((h)->err), "%s failed: %s (%s:%d)", #expr,           \
               cudaGetErrorString(_e), __FILE__, __LINE__);                          \
      return PXR_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

// bracket the dominant kernel launch with events when profiling is on
static inline void pxr_prof_begin(pxr_handle* h, cudaStream_t st) {
  if (h->profile && h->prof_n < PXR_PROFILE_SLOTS) cudaEventRecord(h->prof_ev[2 * h->prof_n], st);
}
static inline void pxr_prof_end(pxr_handle* h, cudaStream_t st) {
  if (h->profile && h->prof_n < PXR_PROFILE_SLOTS) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], st); h->prof_n++; }
}

static inline size_t pxr_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float pxr_apply_act(float x, int act) {
  // reference src/models/multimodal.py:150-167
  switch (act) {
    case PXR_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    case PXR_ACT_TANH: return tanhf(x);
    case PXR_ACT_LEAKY_RELU: return x >= 0.f ? x : 0.01f * x;
    case PXR_ACT_SILU: return x / (1.0f + expf(-x));
    default: return fmaxf(x, 0.f);
  }
}

__device__ __forceinline__ float pxr_apply_final(float z, int fin) {
  // final activation (multimodal.py:381-384) + NaN/Inf guard (multimodal.py:596-597)
  float y = z;
  if (fin == PXR_FINAL_SIGMOID) y = 1.0f / (1.0f + expf(-z));
  else if (fin == PXR_FINAL_TANH) y = tanhf(z);
  if (isnan(y)) y = 0.f;
  else if (isinf(y)) y = y > 0 ? 10.f : -10.f;
  return y;
}

// Order-preserving float -> uint32 (larger float => larger key).
__device__ __forceinline__ uint32_t pxr_ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pxr_unord(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
// Composite ranking key: higher score first, then LOWER item index
// (stable sort over index-ordered candidates, recommender.py:76,105).
__device__ __forceinline__ unsigned long long pxr_key(float score, uint32_t idx) {
  return ((unsigned long long)pxr_ord(score) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float pxr_key_score(unsigned long long k) { return pxr_unord((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t pxr_key_idx(unsigned long long k) { return 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull); }

// ---------------------------------------------------------------------------
// SIMT fp32 row-block linear layers (generic path + item precompute).
// 256 threads; ROWS in {32,16,8,4}.
// ---------------------------------------------------------------------------
#define PXR_SIMT_THREADS 256

// out[r][n] = act(sum_k in[r][k] * Wt[k][n] + bias[n]), n < N, N % 4 == 0,
// in rows 16-byte aligned (ldin % 4 == 0).  act < 0: identity.
template <int ROWS>
__device__ __forceinline__ void linear_rows(const float* in, int ldin, int K, const float* __restrict__ Wt,
                                            const float* __restrict__ bias, int N, float* out, int ldout,
                                            int act) {
  constexpr int RG = ROWS / 4;
  constexpr int CT = PXR_SIMT_THREADS / RG;
  constexpr int CW = CT * 4;
  const int ct = threadIdx.x % CT, rg = threadIdx.x / CT;
  const int K4 = K & ~3;
  for (int nc = 0; nc < N; nc += CW) {
    const int n0 = nc + ct * 4;
    if (n0 >= N) continue;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    for (int k = 0; k < K4; k += 4) {
      float4 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(Wt + (size_t)(k + j) * N + n0);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(in + (size_t)(rg * 4 + r) * ldin + k);
        acc[r][0] += x.x * w[0].x + x.y * w[1].x + x.z * w[2].x + x.w * w[3].x;
        acc[r][1] += x.x * w[0].y + x.y * w[1].y + x.z * w[2].y + x.w * w[3].y;
        acc[r][2] += x.x * w[0].z + x.y * w[1].z + x.z * w[2].z + x.w * w[3].z;
        acc[r][3] += x.x * w[0].w + x.y * w[1].w + x.z * w[2].w + x.w * w[3].w;
      }
    }
    for (int k = K4; k < K; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(Wt + (size_t)k * N + n0);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float x = in[(size_t)(rg * 4 + r) * ldin + k];
        acc[r][0] += x * w.x; acc[r][1] += x * w.y; acc[r][2] += x * w.z; acc[r][3] += x * w.w;
      }
    }
    const float4 b = bias ? *reinterpret_cast<const float4*>(bias + n0) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4 o;
      o.x = acc[r][0] + b.x; o.y = acc[r][1] + b.y; o.z = acc[r][2] + b.z; o.w = acc[r][3] + b.w;
      if (act >= 0) { o.x = pxr_apply_act(o.x, act); o.y = pxr_apply_act(o.y, act); o.z = pxr_apply_act(o.z, act); o.w = pxr_apply_act(o.w, act); }
      *reinterpret_cast<float4*>(out + (size_t)(rg * 4 + r) * ldout + n0) = o;
    }
  }
}

// Small / odd N: one warp per row, lanes stride over k; W row-major [N][K].
template <int ROWS>
__device__ __forceinline__ void linear_small(const float* in, int ldin, int K, const float* __restrict__ W,
                                             const float* __restrict__ bias, int N, float* out, int ldout) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < ROWS; r += PXR_SIMT_THREADS / 32) {
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s += in[(size_t)r * ldin + k] * W[(size_t)n * K + k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) out[(size_t)r * ldout + n] = s + (bias ? bias[n] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------
// kernel launch entry points implemented in the other translation units
// ---------------------------------------------------------------------------
int pxr_simt_smem_rows(const pxr_handle* h, bool items_kernel);
int pxr_launch_items_simt(pxr_handle* h, const float* item_embedding, const int64_t* item_idx,
                          const int64_t* tag_idx, const float* vis, const float* txt, const float* num,
                          int64_t n_rows, int64_t item_base, float* feats_out, cudaStream_t st);
// explicit (user, item row) pairs -> scores (and logits)
int pxr_launch_score_simt(pxr_handle* h, const float* user_embedding, const int64_t* user_idx,
                          const int64_t* item_row, int64_t n_pairs, float* out, float* out_logit, cudaStream_t st);
// generic full-catalogue scoring with the running top-K kept in shared memory (no dense score matrix)
size_t pxr_simt_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k);
int pxr_launch_score_topk_simt(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                               const int64_t* seen_indptr, const int32_t* seen_idx, int32_t k, float* out_scores,
                               int32_t* out_idx, void* ws, size_t ws_bytes, cudaStream_t st);
int pxr_launch_topk_rows(pxr_handle* h, const float* scores, int64_t n_users, int64_t n_items, int64_t item_base,
                         int32_t k, float* out_scores, int32_t* out_idx, cudaStream_t st);
int pxr_launch_merge(const float* scores_in, const int32_t* idx_in, int32_t n_shards, int64_t n_users, int32_t k,
                     float* out_scores, int32_t* out_idx, cudaStream_t st);
// exact mode: fp32 re-score + re-rank of (n_users, list_len) candidate lists (global indices, -1 padded; list_len = 64 or a
// multiple of 64 up to 1 024) -> (n_users, k)
size_t pxr_rescore_list_bytes(int64_t n_users, int32_t list_len);
int pxr_launch_rescore(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                       const int32_t* list_idx, int32_t list_len, int32_t k, float* out_scores, int32_t* out_idx, void* ws,
                       cudaStream_t st);
int pxr_launch_metrics(const int32_t* topk_idx, int32_t k_stride, int64_t n_users, const int64_t* gt_indptr,
                       const int32_t* gt_idx, const int32_t* recall_den, const int32_t* ks, int32_t n_ks, const double* discount,
                       const double* ideal, double* out_sums, void* ws, cudaStream_t st);

// tcgen05 path (score_tc.cu)
bool pxr_tc_supported(const pxr_handle* h);
const char* pxr_tc_unsupported_reason(const pxr_handle* h);   // NULL = supported
#define PXR_TC_H1 512      // columns of the fused kernel's first hidden layer (layer-1 partials are padded to it)
const float* pxr_tc_w1t_padded(const pxr_handle* h);   // concat / wide gated: W1^T as [k][512], zero-padded columns
const float* pxr_tc_b1_padded(const pxr_handle* h);    // b1 zero-padded to 512
bool pxr_tc_gated_wide(const pxr_handle* h);   // gated fusion at embedding_dim != 64: gate-weighted layer-1 partials (F_GATEDW)
#define PXR_TC_MAX_K 1024   // 16 pages of the fused kernel's 64-slot lists
bool pxr_tc_can_run(const pxr_handle* h, int32_t k);   // this call (k, shard size) fits the kernel's limits
size_t pxr_tc_weight_bytes(const pxr_handle* h);
int pxr_tc_prepare_weights(pxr_handle* h, cudaStream_t st);
size_t pxr_tc_item_bytes(const pxr_handle* h, int64_t n_rows);
int pxr_tc_prepare_items(pxr_handle* h, int64_t n_rows, void* ws, cudaStream_t st);
size_t pxr_tc_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k);
// tensor-pipe item precompute (items_tc.cu)
bool pxr_items_tc_supported(const pxr_handle* h);
int pxr_items_tc_prepare_weights(pxr_handle* h, cudaStream_t st);
int pxr_launch_items_tc(pxr_handle* h, const float* item_embedding, const int64_t* item_idx, const int64_t* tag_idx,
                        const float* vis, const float* txt, const float* num, int64_t n_rows, int64_t item_base,
                        float* feats_out, cudaStream_t st);
int pxr_launch_item_pi_tc(pxr_handle* h, int64_t n_rows, uint16_t* out, int fmt16, cudaStream_t st);
int pxr_launch_item_logit_tc(pxr_handle* h, int64_t n_rows, float* out, cudaStream_t st);
int pxr_launch_item_q_tc(pxr_handle* h, int64_t n_rows, uint16_t* out, int fmt16, cudaStream_t st);
int pxr_tc_score_topk(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                      const int64_t* seen_indptr, const int32_t* seen_idx, int32_t k, float* out_scores,
                      int32_t* out_idx, void* ws, size_t ws_bytes, cudaStream_t st);
