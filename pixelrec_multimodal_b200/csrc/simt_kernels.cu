// Generic fp32 SIMT kernels of libpxr.so:
//   K1/K2  items_simt      gather + modality projections, once per item row
//   K3g    score_simt      literal per-pair forward (all fusion types / shapes)
//   K3t    topk_rows       exact row top-K over a dense score block (radix select)
//   K4     merge_topk      S-way merge of per-shard top-K lists
//   K5     metrics         Precision/Recall/F1/HitRate/NDCG/MRR@K sums
// The tcgen05 path (score_tc.cu) replaces K3g+K3t for the supported shapes; these
// kernels remain the path for every other configuration and for explicit pairs.
#include <algorithm>
#include <climits>

#include <stdlib.h>

#include "pxr_common.cuh"

// ===========================================================================
// K1/K2: item precompute (reference multimodal.py:554-570, once per item)
// ===========================================================================
struct ItemsParams {
  int D, M, act;
  const float* item_embedding; const int64_t* item_idx; const int64_t* tag_idx;
  const float* tag_emb;
  const float* in[3]; int in_dim[3]; int has[3]; int slot[3];
  const float* w0t[3]; const float* b0[3]; int n0[3];
  const float* w1t[3]; const float* b1[3];     // optional second layer (n0 -> D)
  int64_t n_rows, item_base;
  float* feats;                                // [n_rows_padded][M-1][D]
  int ld_in, ld_hid;
};

template <int ROWS>
__global__ void __launch_bounds__(PXR_SIMT_THREADS) items_simt_kernel(ItemsParams p) {
  extern __shared__ __align__(16) float smem[];
  float* inbuf = smem;                              // [ROWS][ld_in]
  float* hid = inbuf + (size_t)ROWS * p.ld_in;      // [ROWS][ld_hid]
  const int64_t row0 = (int64_t)blockIdx.x * ROWS;
  const int FD = (p.M - 1) * p.D;
  float* out_base = p.feats + row0 * FD;

  // slots 0 (item embedding) and 1 (tag embedding): plain gathers
  for (int idx = threadIdx.x; idx < ROWS * (p.D / 4) * 2; idx += PXR_SIMT_THREADS) {
    const int r = idx / (p.D / 4 * 2);
    const int rem = idx % (p.D / 4 * 2);
    const int which = rem / (p.D / 4);
    const int c = (rem % (p.D / 4)) * 4;
    const int64_t row = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < p.n_rows) {
      if (which == 0) {
        const int64_t it = p.item_idx ? p.item_idx[row] : (p.item_base + row);
        v = *reinterpret_cast<const float4*>(p.item_embedding + it * p.D + c);
      } else {
        v = *reinterpret_cast<const float4*>(p.tag_emb + p.tag_idx[row] * p.D + c);
      }
    }
    *reinterpret_cast<float4*>(out_base + (size_t)r * FD + which * p.D + c) = v;
  }

  for (int m = 0; m < 3; ++m) {
    if (!p.has[m]) continue;
    const int K = p.in_dim[m];
    __syncthreads();
    for (int idx = threadIdx.x; idx < ROWS * p.ld_in; idx += PXR_SIMT_THREADS) {
      const int r = idx / p.ld_in, k = idx % p.ld_in;
      const int64_t row = row0 + r;
      inbuf[idx] = (row < p.n_rows && k < K) ? p.in[m][row * K + k] : 0.f;
    }
    __syncthreads();
    float* dst = out_base + p.slot[m] * p.D;
    if (p.w1t[m]) {
      linear_rows<ROWS>(inbuf, p.ld_in, K, p.w0t[m], p.b0[m], p.n0[m], hid, p.ld_hid, p.act);
      __syncthreads();
      linear_rows<ROWS>(hid, p.ld_hid, p.n0[m], p.w1t[m], p.b1[m], p.D, dst, FD, p.act);
    } else {
      linear_rows<ROWS>(inbuf, p.ld_in, K, p.w0t[m], p.b0[m], p.D, dst, FD, p.act);
    }
  }
}

static int items_rows_for(const pxr_handle* h, int* ld_in, int* ld_hid, size_t* smem) {
  int kmax = 4;
  const int dims[3] = {h->cfg.vision_dim, h->cfg.language_dim, h->cfg.num_numerical};
  for (int m = 0; m < 3; ++m) if (dims[m] > kmax) kmax = dims[m];
  *ld_in = (kmax + 3) & ~3;
  *ld_hid = h->cfg.projection_hidden > 0 ? ((h->cfg.projection_hidden + 3) & ~3) : 4;
  const int cands[4] = {32, 16, 8, 4};
  for (int i = 0; i < 4; ++i) {
    size_t s = (size_t)cands[i] * (*ld_in + *ld_hid) * sizeof(float);
    if (s <= (size_t)h->max_smem_optin - 1024) { *smem = s; return cands[i]; }
  }
  return 0;
}

template <int ROWS>
static int launch_items(pxr_handle* h, const ItemsParams& p, size_t smem, cudaStream_t st) {
  PXR_CUDA(h, cudaFuncSetAttribute(items_simt_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (p.n_rows + ROWS - 1) / ROWS;
  items_simt_kernel<ROWS><<<(unsigned)blocks, PXR_SIMT_THREADS, smem, st>>>(p);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

int pxr_launch_items_simt(pxr_handle* h, const float* item_embedding, const int64_t* item_idx,
                          const int64_t* tag_idx, const float* vis, const float* txt, const float* num,
                          int64_t n_rows, int64_t item_base, float* feats_out, cudaStream_t st) {
  ItemsParams p;
  memset(&p, 0, sizeof(p));
  p.D = h->cfg.embedding_dim; p.M = h->M; p.act = h->cfg.activation;
  p.item_embedding = item_embedding; p.item_idx = item_idx; p.tag_idx = tag_idx; p.tag_emb = h->tag_emb;
  const float* ins[3] = {vis, txt, num};
  const int dims[3] = {h->cfg.vision_dim, h->cfg.language_dim, h->cfg.num_numerical};
  int slot = 2;
  for (int m = 0; m < 3; ++m) {
    p.has[m] = h->has_mod[m];
    if (!p.has[m]) continue;
    if (!ins[m]) PXR_FAIL(h, PXR_ERR_INVALID, "modality %d is configured but its feature pointer is NULL", m);
    p.in[m] = ins[m]; p.in_dim[m] = dims[m]; p.slot[m] = slot++;
    p.w0t[m] = h->proj[m][0].wt; p.b0[m] = h->proj[m][0].b; p.n0[m] = h->proj[m][0].n;
    p.w1t[m] = h->proj[m][1].wt; p.b1[m] = h->proj[m][1].b;
  }
  p.n_rows = n_rows; p.item_base = item_base; p.feats = feats_out;
  size_t smem = 0;
  const int rows = items_rows_for(h, &p.ld_in, &p.ld_hid, &smem);
  switch (rows) {
    case 32: return launch_items<32>(h, p, smem, st);
    case 16: return launch_items<16>(h, p, smem, st);
    case 8: return launch_items<8>(h, p, smem, st);
    case 4: return launch_items<4>(h, p, smem, st);
  }
  PXR_FAIL(h, PXR_ERR_INVALID, "feature dims too large for the item precompute kernel");
}

// ===========================================================================
// K3g: literal per-pair forward (reference multimodal.py:553-597)
// ===========================================================================
struct ScoreParams {
  int fusion, D, M, n_hidden, hidden[PXR_MAX_HIDDEN], heads, act, fin, maxdim;
  const float* gate_w; const float* gate_b;
  const float* attn_in_wt; const float* attn_in_b; const float* attn_out_wt; const float* attn_out_b;
  const float* ln_w; const float* ln_b;
  const float* mlp_wt[PXR_MAX_HIDDEN]; const float* mlp_b[PXR_MAX_HIDDEN];
  const float* out_w; const float* out_b;
  const float* user_emb; const int64_t* user_idx; const int64_t* item_row;
  const float* item_feats; int64_t n_items; int64_t n_pairs; int dense; int64_t item_base;
  const uint8_t* item_missing;                 // per item row: 1 = features missing, score is 0.0 (recommender.py:229-230)
  const int64_t* seen_indptr; const int32_t* seen_idx;
  float* out; float* out_logit;
};

#define MERGE_MAX_KEYS_SIMT 4096
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n_pow2);

// The forward of ROWS pairs whose (user row, item row) are in prow[2 r], prow[2 r + 1] (user < 0: padding row, computed
// on zeros): leaves the pre-activation logit of row r in G[8 r].  Called by every thread of the block.
template <int ROWS>
__device__ __forceinline__ void score_rows(const ScoreParams& p, float* smem) {
  const int D = p.D, M = p.M, MD = M * D, maxdim = p.maxdim;
  float* X = smem;                                   // [ROWS][M*D]: the token stack == the concat vector
  float* A = X + (size_t)ROWS * MD;                  // [ROWS][maxdim]
  float* B = A + (size_t)ROWS * maxdim;              // [ROWS][maxdim]
  float* G = B + (size_t)ROWS * maxdim;              // [ROWS][8]
  const long long* prow = reinterpret_cast<const long long*>(G + ROWS * 8);   // [ROWS][2] (user row, item row)
  for (int idx = threadIdx.x; idx < ROWS * (MD / 4); idx += PXR_SIMT_THREADS) {
    const int r = idx / (MD / 4), c = (idx % (MD / 4)) * 4;
    const long long u = prow[r * 2], ir = prow[r * 2 + 1];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u >= 0) {
      v = (c < D) ? *reinterpret_cast<const float4*>(p.user_emb + u * D + c)
                  : *reinterpret_cast<const float4*>(p.item_feats + ir * (MD - D) + (c - D));
    }
    *reinterpret_cast<float4*>(X + (size_t)r * MD + c) = v;
  }
  __syncthreads();

  const float* cur = X; int ldcur = MD, Kcur = MD;
  float* nxt = A;
  if (p.fusion == PXR_FUSION_GATED) {
    // reference layers.py:195-225
    linear_small<ROWS>(X, MD, MD, p.gate_w, p.gate_b, M, G, 8);
    __syncthreads();
    if (threadIdx.x < ROWS) {
      float* g = G + threadIdx.x * 8;
      float mx = g[0];
      for (int m = 1; m < M; ++m) mx = fmaxf(mx, g[m]);
      float s = 0.f;
      for (int m = 0; m < M; ++m) { g[m] = expf(g[m] - mx); s += g[m]; }
      for (int m = 0; m < M; ++m) g[m] /= s;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < ROWS * D; idx += PXR_SIMT_THREADS) {
      const int r = idx / D, d = idx % D;
      float s = 0.f;
      for (int m = 0; m < M; ++m) s += G[r * 8 + m] * X[(size_t)r * MD + m * D + d];
      A[(size_t)r * maxdim + d] = s;
    }
    __syncthreads();
    cur = A; ldcur = maxdim; Kcur = D; nxt = B;
  } else if (p.fusion == PXR_FUSION_ATTENTION) {
    // reference layers.py:135-164 (documented semantics), nn.MultiheadAttention batch_first=False
    const int D3 = 3 * D, dh = D / p.heads;
    for (int m = 0; m < M; ++m)
      linear_rows<ROWS>(X + m * D, MD, D, p.attn_in_wt, p.attn_in_b, D3, A + m * D3, maxdim, -1);
    __syncthreads();
    const float scale = rsqrtf((float)dh);
    for (int w = threadIdx.x; w < ROWS * p.heads * M; w += PXR_SIMT_THREADS) {
      const int r = w / (p.heads * M), hd = (w / M) % p.heads, a = w % M;
      float* row = A + (size_t)r * maxdim;
      float* q = row + a * D3 + hd * dh;
      float s[PXR_MAX_MODALITIES];
      float mx = -INFINITY;
      for (int b = 0; b < M; ++b) {
        const float* kk = row + b * D3 + D + hd * dh;
        float acc = 0.f;
        for (int d = 0; d < dh; ++d) acc += q[d] * kk[d];
        s[b] = acc * scale; mx = fmaxf(mx, s[b]);
      }
      float sum = 0.f;
      for (int b = 0; b < M; ++b) { s[b] = expf(s[b] - mx); sum += s[b]; }
      const float inv = 1.f / sum;
      for (int d = 0; d < dh; ++d) {
        float o = 0.f;
        for (int b = 0; b < M; ++b) o += s[b] * inv * row[b * D3 + 2 * D + hd * dh + d];
        q[d] = o;                                   // O overwrites this work item's own Q slice
      }
    }
    __syncthreads();
    for (int m = 0; m < M; ++m)
      linear_rows<ROWS>(A + m * D3, maxdim, D, p.attn_out_wt, p.attn_out_b, D, B + m * D, maxdim, -1);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < ROWS; r += PXR_SIMT_THREADS / 32) {
      float fused[16];                               // D <= 512
#pragma unroll
      for (int j = 0; j < 16; ++j) fused[j] = 0.f;
      for (int m = 0; m < M; ++m) {
        float y[16];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int d = lane + j * 32;
          y[j] = d < D ? X[(size_t)r * MD + m * D + d] + B[(size_t)r * maxdim + m * D + d] : 0.f;
          sum += y[j];
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mu = sum / D;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { const int d = lane + j * 32; if (d < D) var += (y[j] - mu) * (y[j] - mu); }
        for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rstd = rsqrtf(var / D + 1e-5f);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int d = lane + j * 32;
          if (d < D) fused[j] += (y[j] - mu) * rstd * p.ln_w[d] + p.ln_b[d];
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) { const int d = lane + j * 32; if (d < D) A[(size_t)r * maxdim + d] = fused[j] / M; }
    }
    __syncthreads();
    cur = A; ldcur = maxdim; Kcur = D; nxt = B;
  }

  // prediction MLP with eval-mode BatchNorm folded into the next Linear (multimodal.py:366-386)
  for (int l = 0; l < p.n_hidden; ++l) {
    linear_rows<ROWS>(cur, ldcur, Kcur, p.mlp_wt[l], p.mlp_b[l], p.hidden[l], nxt, maxdim, p.act);
    __syncthreads();
    cur = nxt; ldcur = maxdim; Kcur = p.hidden[l];
    nxt = (cur == A) ? B : A;
  }
  linear_small<ROWS>(cur, ldcur, Kcur, p.out_w, p.out_b, 1, G, 8);
  __syncthreads();
}

template <int ROWS>
__global__ void __launch_bounds__(PXR_SIMT_THREADS) score_simt_kernel(ScoreParams p) {
  extern __shared__ __align__(16) float smem[];
  const int MD = p.M * p.D, maxdim = p.maxdim;
  float* G = smem + (size_t)ROWS * MD + 2 * (size_t)ROWS * maxdim;
  long long* prow = reinterpret_cast<long long*>(G + ROWS * 8);
  const int64_t row0 = (int64_t)blockIdx.x * ROWS;
  if (threadIdx.x < ROWS) {
    const int64_t pair = row0 + threadIdx.x;
    long long u = -1, ir = 0;
    if (pair < p.n_pairs) { u = p.user_idx[pair]; ir = p.item_row[pair]; }
    if (u < 0) ir = 0;
    prow[threadIdx.x * 2] = u; prow[threadIdx.x * 2 + 1] = ir;
  }
  __syncthreads();
  score_rows<ROWS>(p, smem);
  if (threadIdx.x < ROWS) {
    const int64_t pair = row0 + threadIdx.x;
    if (pair < p.n_pairs) {
      float z = G[threadIdx.x * 8];
      float s = pxr_apply_final(z, p.fin);
      if (p.item_missing && p.item_missing[prow[threadIdx.x * 2 + 1]]) { s = 0.f; z = 0.f; }
      p.out[pair] = s;
      if (p.out_logit) p.out_logit[pair] = z;
    }
  }
}

// K3g + K3t fused: one block = one user x one item split.  The block sweeps its items ROWS at a time through the same
// literal forward and keeps the user's running top-K in shared memory (64-bit keys (score, ~index): ties -> lower index):
// a row enters the candidate buffer only if it beats the current K-th key; a full buffer is compacted by a bitonic sort.
// The users x items scores never reach HBM on this path either.
#define TOPKS_MAX_K 1024
struct TopkSimtParams {
  int64_t n_users; int S; int64_t rows_per_split; int k, kp, cap;
  float* out_scores; int32_t* out_idx;               // [S][n_users][k]
};

template <int ROWS>
__global__ void __launch_bounds__(PXR_SIMT_THREADS) score_topk_simt_kernel(ScoreParams p, TopkSimtParams tp) {
  extern __shared__ __align__(16) float smem[];
  const int MD = p.M * p.D, maxdim = p.maxdim;
  float* G = smem + (size_t)ROWS * MD + 2 * (size_t)ROWS * maxdim;
  long long* prow = reinterpret_cast<long long*>(G + ROWS * 8);
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(prow + ROWS * 2);      // [cap]
  __shared__ int s_cnt;
  __shared__ unsigned long long s_thr;
  const int64_t ul = blockIdx.x / tp.S;
  const int sp = blockIdx.x % tp.S;
  const int64_t lo = (int64_t)sp * tp.rows_per_split, hi = min(p.n_items, lo + tp.rows_per_split);
  const long long user = p.user_idx[ul];
  for (int i = threadIdx.x; i < tp.cap; i += PXR_SIMT_THREADS) buf[i] = 0ull;
  if (threadIdx.x == 0) { s_cnt = 0; s_thr = 0ull; }
  int64_t seen_lo = 0, seen_hi = 0;
  if (p.seen_indptr) { seen_lo = p.seen_indptr[ul]; seen_hi = p.seen_indptr[ul + 1]; }
  __syncthreads();
  auto compact = [&]() {            // keep the best k of the buffer, sorted; refresh the admission threshold
    bitonic_sort_desc(buf, tp.cap);
    for (int i = tp.k + threadIdx.x; i < tp.cap; i += PXR_SIMT_THREADS) buf[i] = 0ull;
    __syncthreads();
    if (threadIdx.x == 0) { s_cnt = min(s_cnt, tp.k); s_thr = s_cnt >= tp.k ? buf[tp.k - 1] : 0ull; }
    __syncthreads();
  };
  for (int64_t r0 = lo; r0 < hi; r0 += ROWS) {
    if (threadIdx.x < ROWS) {
      const int64_t ir = r0 + threadIdx.x;
      prow[threadIdx.x * 2] = ir < hi ? user : -1;
      prow[threadIdx.x * 2 + 1] = ir < hi ? ir : 0;
    }
    __syncthreads();
    score_rows<ROWS>(p, smem);
    if (s_cnt > tp.cap - ROWS) compact();            // block-uniform: s_cnt is only written between barriers
    if (threadIdx.x < ROWS) {
      const int64_t ir = r0 + threadIdx.x;
      if (ir < hi) {
        float s = pxr_apply_final(G[threadIdx.x * 8], p.fin);
        if (p.item_missing && p.item_missing[ir]) s = 0.f;
        const int32_t gi = (int32_t)(p.item_base + ir);
        bool seen = false;
        if (p.seen_indptr) {
          int64_t a = seen_lo, b = seen_hi;
          while (a < b) { const int64_t mid = (a + b) >> 1; if (p.seen_idx[mid] < gi) a = mid + 1; else b = mid; }
          seen = a < seen_hi && p.seen_idx[a] == gi;   // filtered (recommender.py:88-90)
        }
        const unsigned long long key = pxr_key(s, (uint32_t)gi);
        if (!seen && key > s_thr) buf[atomicAdd(&s_cnt, 1)] = key;
      }
    }
    __syncthreads();
  }
  compact();
  for (int i = threadIdx.x; i < tp.k; i += PXR_SIMT_THREADS) {
    const unsigned long long kv = buf[i];
    const int64_t o = ((int64_t)sp * tp.n_users + ul) * tp.k + i;
    tp.out_scores[o] = kv ? pxr_key_score(kv) : -INFINITY;
    tp.out_idx[o] = kv ? (int32_t)pxr_key_idx(kv) : -1;
  }
}

static int score_maxdim(const pxr_handle* h) {
  int md = h->cfg.embedding_dim;
  for (int l = 0; l < h->cfg.n_hidden; ++l) if (h->cfg.hidden[l] > md) md = h->cfg.hidden[l];
  if (h->cfg.fusion == PXR_FUSION_ATTENTION) { int q = h->M * 3 * h->cfg.embedding_dim; if (q > md) md = q; }
  return (md + 3) & ~3;
}

static size_t score_smem(const pxr_handle* h, int rows) {
  const int MD = h->M * h->cfg.embedding_dim;
  return ((size_t)rows * MD + 2 * (size_t)rows * score_maxdim(h) + (size_t)rows * 8) * sizeof(float) +
         (size_t)rows * 2 * sizeof(long long);
}

int pxr_simt_smem_rows(const pxr_handle* h, bool items_kernel) {
  if (items_kernel) { int a, b; size_t s; return items_rows_for(h, &a, &b, &s); }
  // 16 rows per block, not the largest that fits: two resident blocks per SM hide the weight-load latency better than one
  // block with twice the rows (262 144 pairs of the default MLP: 8.05 ms with 32 rows, 5.56 ms with 16, 6.32 ms with 8;
  // profiles/r02_rescore_rows.log).  PXR_SIMT_ROWS overrides for experiments.
  const int cands[4] = {32, 16, 8, 4};
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("PXR_SIMT_ROWS"); forced = e ? atoi(e) : 16; }
  for (int i = 0; i < 4; ++i) {
    if (forced && cands[i] > forced) continue;
    if (score_smem(h, cands[i]) <= (size_t)h->max_smem_optin - 1024) return cands[i];
  }
  return 0;
}

template <int ROWS>
static int launch_score(pxr_handle* h, const ScoreParams& p, cudaStream_t st) {
  const size_t smem = score_smem(h, ROWS);
  PXR_CUDA(h, cudaFuncSetAttribute(score_simt_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (p.n_pairs + ROWS - 1) / ROWS;
  if (blocks > 0x7fffffffLL) PXR_FAIL(h, PXR_ERR_INVALID, "too many pairs for one launch");
  score_simt_kernel<ROWS><<<(unsigned)blocks, PXR_SIMT_THREADS, smem, st>>>(p);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

static void fill_score_params(const pxr_handle* h, ScoreParams& p) {
  memset(&p, 0, sizeof(p));
  p.fusion = h->cfg.fusion; p.D = h->cfg.embedding_dim; p.M = h->M; p.n_hidden = h->cfg.n_hidden;
  for (int l = 0; l < p.n_hidden; ++l) { p.hidden[l] = h->cfg.hidden[l]; p.mlp_wt[l] = h->mlp[l].wt; p.mlp_b[l] = h->mlp[l].b; }
  p.heads = h->cfg.num_heads; p.act = h->cfg.activation; p.fin = h->cfg.final_activation; p.maxdim = score_maxdim(h);
  p.gate_w = h->gate.w; p.gate_b = h->gate.b;
  p.attn_in_wt = h->attn_in.wt; p.attn_in_b = h->attn_in.b; p.attn_out_wt = h->attn_out.wt; p.attn_out_b = h->attn_out.b;
  p.ln_w = h->ln_w; p.ln_b = h->ln_b;
  p.out_w = h->out.w; p.out_b = h->out.b;
  p.item_feats = h->item_feats; p.n_items = h->n_rows; p.item_base = h->item_base; p.item_missing = h->item_missing;
}

// explicit (user, item row) pairs -> scores (and logits)
int pxr_launch_score_simt(pxr_handle* h, const float* user_embedding, const int64_t* user_idx,
                          const int64_t* item_row, int64_t n_pairs, float* out, float* out_logit, cudaStream_t st) {
  if (n_pairs == 0) return PXR_OK;
  ScoreParams p;
  fill_score_params(h, p);
  p.user_emb = user_embedding; p.user_idx = user_idx; p.item_row = item_row; p.n_pairs = n_pairs;
  p.out = out; p.out_logit = out_logit;
  switch (pxr_simt_smem_rows(h, false)) {
    case 32: return launch_score<32>(h, p, st);
    case 16: return launch_score<16>(h, p, st);
    case 8: return launch_score<8>(h, p, st);
    case 4: return launch_score<4>(h, p, st);
  }
  PXR_FAIL(h, PXR_ERR_INVALID, "layer dims too large for the SIMT scoring kernel");
}

// ---- generic full-catalogue path: per-user blocks with the running top-K in shared memory (no dense score matrix)
static int topk_simt_kp(int k, int rows) { int kp = 1; while (kp < k || kp < rows) kp <<= 1; return kp; }

static int topk_simt_rows(const pxr_handle* h, int k) {
  const int cands[4] = {32, 16, 8, 4};
  const int base = pxr_simt_smem_rows(h, false);
  for (int i = 0; i < 4; ++i) {
    if (cands[i] > base) continue;
    const size_t sm = score_smem(h, cands[i]) + (size_t)2 * topk_simt_kp(k, cands[i]) * 8;
    if (sm <= (size_t)h->max_smem_optin - 1024) return cands[i];
  }
  return 0;
}

static int topk_simt_splits(const pxr_handle* h, int64_t n_users, int k) {
  // enough blocks to fill the SMs twice when the user block is small; bounded by K4's merge width and the tile count
  int64_t s = (2 * (int64_t)h->n_sm + n_users - 1) / (n_users > 0 ? n_users : 1);
  s = std::min<int64_t>(s, MERGE_MAX_KEYS_SIMT / k);
  s = std::min<int64_t>(s, (h->n_rows + 255) / 256);
  return (int)std::max<int64_t>(s, 1);
}

size_t pxr_simt_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k) {
  const int S = topk_simt_splits(h, n_users, k);
  return S > 1 ? pxr_align_up((size_t)S * n_users * k * 8, 256) + 256 : 256;
}

template <int ROWS>
static int launch_score_topk(pxr_handle* h, const ScoreParams& p, TopkSimtParams tp, cudaStream_t st) {
  tp.kp = topk_simt_kp(tp.k, ROWS); tp.cap = 2 * tp.kp;
  const size_t smem = score_smem(h, ROWS) + (size_t)tp.cap * 8;
  PXR_CUDA(h, cudaFuncSetAttribute(score_topk_simt_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = tp.n_users * tp.S;
  if (blocks > 0x7fffffffLL) PXR_FAIL(h, PXR_ERR_INVALID, "too many users for one launch");
  pxr_prof_begin(h, st);
  score_topk_simt_kernel<ROWS><<<(unsigned)blocks, PXR_SIMT_THREADS, smem, st>>>(p, tp);
  pxr_prof_end(h, st);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

int pxr_launch_score_topk_simt(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                               const int64_t* seen_indptr, const int32_t* seen_idx, int32_t k, float* out_scores,
                               int32_t* out_idx, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (k > TOPKS_MAX_K) PXR_FAIL(h, PXR_ERR_INVALID, "k=%d exceeds %d", k, TOPKS_MAX_K);
  ScoreParams p;
  fill_score_params(h, p);
  p.user_emb = user_embedding; p.user_idx = user_idx; p.seen_indptr = seen_indptr; p.seen_idx = seen_idx;
  TopkSimtParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.n_users = n_users; tp.k = k; tp.S = topk_simt_splits(h, n_users, k);
  tp.rows_per_split = (h->n_rows + tp.S - 1) / tp.S;
  tp.out_scores = out_scores; tp.out_idx = out_idx;
  if (tp.S > 1) {
    const size_t need = (size_t)tp.S * n_users * k;
    if (ws_bytes < need * 8) PXR_FAIL(h, PXR_ERR_WORKSPACE, "top-K workspace too small");
    tp.out_scores = (float*)ws; tp.out_idx = (int32_t*)((char*)ws + need * 4);
  }
  int rc;
  switch (topk_simt_rows(h, k)) {
    case 32: rc = launch_score_topk<32>(h, p, tp, st); break;
    case 16: rc = launch_score_topk<16>(h, p, tp, st); break;
    case 8: rc = launch_score_topk<8>(h, p, tp, st); break;
    case 4: rc = launch_score_topk<4>(h, p, tp, st); break;
    default: PXR_FAIL(h, PXR_ERR_INVALID, "layer dims / top_k too large for the SIMT scoring kernel");
  }
  if (rc) return rc;
  if (tp.S > 1) {
    rc = pxr_launch_merge(tp.out_scores, tp.out_idx, tp.S, n_users, k, out_scores, out_idx, st);
    h->launches++;
    if (rc) PXR_FAIL(h, rc, "top-K merge of %d item splits failed", tp.S);
  }
  return PXR_OK;
}

// ===========================================================================
// K3t: exact top-K of each row of a dense score block (masked entries = -inf).
// 64-bit keys (score, ~index) are all distinct, so an 8-pass radix select finds
// the exact K-th key and ties resolve to the lower item index by construction.
// ===========================================================================
#define TOPK_THREADS 256
#define TOPK_MAX_K 1024

__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = ((i & size) == 0);
          const unsigned long long a = keys[i], b = keys[j];
          if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[j] = a; }
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_rows_kernel(const float* __restrict__ scores, int64_t n_items,
                                                                  int64_t item_base, int k, float* out_scores,
                                                                  int32_t* out_idx) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long sel[TOPK_MAX_K];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_count, s_nvalid;
  const float* row = scores + (int64_t)blockIdx.x * n_items;
  if (threadIdx.x == 0) { s_prefix = 0ull; s_remaining = k; s_count = 0; s_nvalid = 0; }
  __syncthreads();
  // number of unmasked candidates
  int local = 0;
  for (int64_t i = threadIdx.x; i < n_items; i += TOPK_THREADS) local += (row[i] != -INFINITY);
  atomicAdd(&s_nvalid, local);
  __syncthreads();
  const int keff = min(k, s_nvalid);
  if (threadIdx.x == 0) s_remaining = keff;
  __syncthreads();
  if (keff > 0) {
    for (int pass = 7; pass >= 0; --pass) {
      for (int i = threadIdx.x; i < 256; i += TOPK_THREADS) hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int64_t i = threadIdx.x; i < n_items; i += TOPK_THREADS) {
        const float s = row[i];
        if (s == -INFINITY) continue;
        const unsigned long long key = pxr_key(s, (uint32_t)i);
        if (pass == 7 || (key >> (8 * (pass + 1))) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255ull], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int rem = s_remaining, b = 255;
        for (; b > 0; --b) { if ((int)hist[b] >= rem) break; rem -= hist[b]; }
        s_remaining = rem; s_prefix = (prefix << 8) | (unsigned long long)b;
      }
      __syncthreads();
    }
    const unsigned long long kth = s_prefix;
    for (int64_t i = threadIdx.x; i < n_items; i += TOPK_THREADS) {
      const float s = row[i];
      if (s == -INFINITY) continue;
      const unsigned long long key = pxr_key(s, (uint32_t)i);
      if (key >= kth) { const int slot = atomicAdd(&s_count, 1); if (slot < TOPK_MAX_K) sel[slot] = key; }
    }
  }
  __syncthreads();
  int n2 = 1;
  while (n2 < keff) n2 <<= 1;
  for (int i = keff + threadIdx.x; i < n2; i += TOPK_THREADS) sel[i] = 0ull;
  if (keff > 1) bitonic_sort_desc(sel, n2);
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += TOPK_THREADS) {
    float s = -INFINITY; int32_t id = -1;
    if (i < keff) { s = pxr_key_score(sel[i]); id = (int32_t)(pxr_key_idx(sel[i]) + item_base); }
    out_scores[(int64_t)blockIdx.x * k + i] = s;
    out_idx[(int64_t)blockIdx.x * k + i] = id;
  }
}

int pxr_launch_topk_rows(pxr_handle* h, const float* scores, int64_t n_users, int64_t n_items, int64_t item_base,
                         int32_t k, float* out_scores, int32_t* out_idx, cudaStream_t st) {
  if (k > TOPK_MAX_K) PXR_FAIL(h, PXR_ERR_INVALID, "k=%d exceeds %d", k, TOPK_MAX_K);
  if (n_users == 0) return PXR_OK;
  topk_rows_kernel<<<(unsigned)n_users, TOPK_THREADS, 0, st>>>(scores, n_items, item_base, k, out_scores, out_idx);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

// ===========================================================================
// K4: merge S per-shard top-K lists per user (keys distinct; k <= 64: registers + shuffles, k > 64: shared memory)
// ===========================================================================
#define MERGE_MAX_KEYS 4096
#define MERGE_MAX_WARPS 8

// Shared-memory variant (k > 64).  One warp per user.  Every per-shard list is already sorted by (score desc, index asc), so the S lists are merged
// pairwise in a tree, each 2-way merge keeping only the best k: output position r of a pair is found by a
// merge-path binary search (log2 k steps) and all positions are independent, so lanes take r = lane, lane + 32, ...
// Keys are the 64-bit composites of pxr_key (unique per item; 0 = padding, smallest), staged in shared memory.
// HBM traffic is the algorithmic 8 k S bytes in + 8 k bytes out per user (coalesced: a user's k entries of one
// shard are contiguous, consecutive users of a block too).
__global__ void __launch_bounds__(32 * MERGE_MAX_WARPS) merge_topk_kernel(const float* __restrict__ scores_in,
                                                                           const int32_t* __restrict__ idx_in, int n_shards,
                                                                           int64_t n_users, int k, int warps_per_block,
                                                                           float* __restrict__ out_scores,
                                                                           int32_t* __restrict__ out_idx) {
  extern __shared__ unsigned long long merge_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps_per_block) return;
  const int half_lists = (n_shards + 1) / 2;
  unsigned long long* bufA = merge_smem + (size_t)warp * (n_shards + half_lists) * k;   // n_shards lists
  unsigned long long* bufB = bufA + (size_t)n_shards * k;                               // ceil(n_shards / 2) lists
  for (int64_t u = (int64_t)blockIdx.x * warps_per_block + warp; u < n_users; u += (int64_t)gridDim.x * warps_per_block) {
    // stage the S lists: the loads of 8 elements per lane are issued together (the copy is latency-bound otherwise)
    // stage the S lists, 4 shards (8 independent loads per lane) at a time; no integer divisions in the loops
    for (int s0 = 0; s0 < n_shards; s0 += 4) {
      for (int j = lane; j < k; j += 32) {
        int32_t id[4]; float sc[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (s0 + m < n_shards) {
            const int64_t off = ((int64_t)(s0 + m) * n_users + u) * k + j;
            id[m] = idx_in[off]; sc[m] = scores_in[off];
          }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (s0 + m < n_shards) bufA[(s0 + m) * k + j] = id[m] < 0 ? 0ull : pxr_key(sc[m], (uint32_t)id[m]);
      }
    }
    __syncwarp();
    unsigned long long* src = bufA; unsigned long long* dst = bufB;
    int n = n_shards;
    while (n > 1) {
      const int pairs = n / 2;
      for (int pr = 0; pr < pairs; ++pr) {
        const unsigned long long* A = src + (2 * pr) * k;
        const unsigned long long* B = A + k;
        for (int r = lane; r < k; r += 32) {
          int lo = 0, hi = r;                    // i = number of the first r outputs that come from A (both lists hold k keys)
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (A[mid] > B[r - mid - 1]) lo = mid + 1; else hi = mid;
          }
          const int i = lo, j = r - lo;
          const unsigned long long a = i < k ? A[i] : 0ull, b2 = j < k ? B[j] : 0ull;
          dst[pr * k + r] = a > b2 ? a : b2;
        }
      }
      if (n & 1) for (int j = lane; j < k; j += 32) dst[pairs * k + j] = src[(n - 1) * k + j];
      __syncwarp();
      unsigned long long* tmp = src; src = dst; dst = tmp;
      n = pairs + (n & 1);
    }
    for (int j = lane; j < k; j += 32) {
      const unsigned long long key = src[j];
      out_scores[u * k + j] = key ? pxr_key_score(key) : -INFINITY;
      out_idx[u * k + j] = key ? (int32_t)pxr_key_idx(key) : -1;
    }
    __syncwarp();
  }
}

// Register variant for k <= 64 (every list fits a warp: slot j of a list lives in lane j & 31, register j >> 5).
// No shared memory: the S lists are folded into one accumulator list by bitonic top-64 merges.  Both lists are
// descending, so max(A[j], B[63 - j]) is a bitonic sequence that holds the best 64 keys of the union (B is
// reversed by one lane-mirror shuffle and a register swap); six compare-exchange stages (distance 32 inside the
// lane, 16..1 by shfl.xor) sort it again.  24 32-bit shuffles and ~60 integer instructions per 2-way merge instead
// of k merge-path searches through shared memory; the loads of MERGE_GROUP lists (2 * MERGE_GROUP scores + indices
// per lane) are issued together before the first merge, so a resident warp keeps 1.6 KB in flight at k = 50.
// (Issuing them one step ahead of the merges costs 20 registers and measured slower, 0.27 vs 0.24 ms: the kernel is
// bound by instruction issue -- 168 shuffles and ~600 integer instructions per user -- not by load latency.)
#define MERGE_GROUP 4
// one compare (2 ISETP on the 64-bit key) and one conditional move (2 SEL) per compare-exchange half: the lane keeps its
// own key or takes the partner's, never computes max and min separately
__device__ __forceinline__ void merge_cx(unsigned long long& x, unsigned long long partner, bool keep_max) {
  const bool take = (partner > x) == keep_max;
  x = take ? partner : x;
}

__device__ __forceinline__ void merge_top64(unsigned long long& a0, unsigned long long& a1, unsigned long long b0,
                                            unsigned long long b1, int lane) {
  const unsigned long long r0 = __shfl_xor_sync(0xffffffffu, b1, 31);      // reversed B: slot lane      <- B[63 - lane]
  const unsigned long long r1 = __shfl_xor_sync(0xffffffffu, b0, 31);      //             slot 32 + lane <- B[31 - lane]
  unsigned long long x0 = a0, x1 = a1;
  merge_cx(x0, r0, true);
  merge_cx(x1, r1, true);
  { const bool sw = x1 > x0; const unsigned long long hi = sw ? x1 : x0, lo = sw ? x0 : x1; x0 = hi; x1 = lo; }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, x0, d), p1 = __shfl_xor_sync(0xffffffffu, x1, d);
    const bool keep_max = (lane & d) == 0;                                  // descending: the lower lane keeps the larger key
    merge_cx(x0, p0, keep_max);
    merge_cx(x1, p1, keep_max);
  }
  a0 = x0; a1 = x1;
}

__global__ void __launch_bounds__(256) merge_topk_reg_kernel(const float* __restrict__ scores_in,
                                                             const int32_t* __restrict__ idx_in, int n_shards,
                                                             int64_t n_users, int k, float* __restrict__ out_scores,
                                                             int32_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool in0 = lane < k, in1 = lane + 32 < k;
  const int64_t list_stride = n_users * k;
  for (int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < n_users; u += warps) {
    unsigned long long a0 = 0ull, a1 = 0ull;
    const float* sp = scores_in + u * k + lane;                             // list s of this user: + s * list_stride
    const int32_t* ip = idx_in + u * k + lane;
    for (int s0 = 0; s0 < n_shards; s0 += MERGE_GROUP, sp += MERGE_GROUP * list_stride, ip += MERGE_GROUP * list_stride) {
      float sc[MERGE_GROUP][2]; int32_t id[MERGE_GROUP][2];
#pragma unroll
      for (int m = 0; m < MERGE_GROUP; ++m) {
        id[m][0] = id[m][1] = -1; sc[m][0] = sc[m][1] = 0.f;
        if (s0 + m < n_shards) {
          if (in0) { id[m][0] = __ldg(ip + m * list_stride); sc[m][0] = __ldg(sp + m * list_stride); }
          if (in1) { id[m][1] = __ldg(ip + m * list_stride + 32); sc[m][1] = __ldg(sp + m * list_stride + 32); }
        }
      }
#pragma unroll
      for (int m = 0; m < MERGE_GROUP; ++m) {
        if (s0 + m < n_shards) {                                            // warp-uniform
          const unsigned long long b0 = id[m][0] < 0 ? 0ull : pxr_key(sc[m][0], (uint32_t)id[m][0]);
          const unsigned long long b1 = id[m][1] < 0 ? 0ull : pxr_key(sc[m][1], (uint32_t)id[m][1]);
          if (s0 + m == 0) { a0 = b0; a1 = b1; } else merge_top64(a0, a1, b0, b1, lane);
        }
      }
    }
    if (in0) { out_scores[u * k + lane] = a0 ? pxr_key_score(a0) : -INFINITY; out_idx[u * k + lane] = a0 ? (int32_t)pxr_key_idx(a0) : -1; }
    if (in1) { out_scores[u * k + lane + 32] = a1 ? pxr_key_score(a1) : -INFINITY; out_idx[u * k + lane + 32] = a1 ? (int32_t)pxr_key_idx(a1) : -1; }
  }
}

int pxr_launch_merge(const float* scores_in, const int32_t* idx_in, int32_t n_shards, int64_t n_users, int32_t k,
                     float* out_scores, int32_t* out_idx, cudaStream_t st) {
  if ((int64_t)n_shards * k > MERGE_MAX_KEYS) return PXR_ERR_INVALID;
  if (n_users == 0) return PXR_OK;
  if (k <= 64) {
    int dev_ = 0, sms = 148;
    cudaGetDevice(&dev_);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_);
    const int64_t want_blocks = (n_users + 7) / 8;                          // 8 warps per block, one user per warp and pass
    const unsigned nb = (unsigned)std::min<int64_t>(want_blocks, (int64_t)sms * 8 * 4);
    merge_topk_reg_kernel<<<nb, 256, 0, st>>>(scores_in, idx_in, n_shards, n_users, k, out_scores, out_idx);
    return cudaGetLastError() == cudaSuccess ? PXR_OK : PXR_ERR_CUDA;
  }
  const size_t per_warp = (size_t)(n_shards + (n_shards + 1) / 2) * k * sizeof(unsigned long long);
  int wpb = (int)std::min<size_t>(MERGE_MAX_WARPS, (96 * 1024) / per_warp);
  if (wpb < 1) return PXR_ERR_INVALID;
  const size_t smem = per_warp * wpb;
  // per device and per function; idempotent (this path only serves k > 64)
  if (cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) return PXR_ERR_CUDA;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n_users + wpb - 1) / wpb;
  const unsigned blocks = (unsigned)std::min<int64_t>(want, (int64_t)n_sm * 8);
  merge_topk_kernel<<<blocks, 32 * wpb, smem, st>>>(scores_in, idx_in, n_shards, n_users, k, wpb, out_scores, out_idx);
  return cudaGetLastError() == cudaSuccess ? PXR_OK : PXR_ERR_CUDA;
}

// ===========================================================================
// Exact mode of the fused path: fp32 re-score of the candidate lists the 16-bit kernel kept (K3r).
// The fused kernel ranks with 16-bit operands; its 64-slot list per user is a superset of the exact top-K whenever the
// exact top-K lies inside the 16-bit top-64 (measured at catalogue scale: always, tests/test_gpu_parity.py).  The
// candidates are scored again with the literal fp32 arithmetic of pxr_score_pairs (score_simt_kernel) and ranked
// again with the reference's tie-break (stable sort over index order, recommender.py:105 -> lower index first).
// ===========================================================================
__global__ void rescore_prep_kernel(const int64_t* __restrict__ user_idx, const int32_t* __restrict__ list_idx, int64_t n_pairs,
                                    int L, int64_t item_base, int64_t n_rows, int64_t* __restrict__ pair_user,
                                    int64_t* __restrict__ pair_row) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const int32_t gi = list_idx[i];
  const int64_t r = (int64_t)gi - item_base;
  const bool ok = gi >= 0 && r >= 0 && r < n_rows;
  pair_user[i] = ok ? user_idx[i / L] : -1;        // -1: padding slot, score_simt_kernel computes on zeros and the sort drops it
  pair_row[i] = ok ? r : 0;
}

// full bitonic sort (descending) of 64 keys held two per lane: slot lane in x0, slot lane + 32 in x1
__device__ __forceinline__ void sort64_desc(unsigned long long& x0, unsigned long long& x1, int lane) {
#pragma unroll
  for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride == 32) {                                   // size == 64: partner is the other register of the lane
        if (x1 > x0) { const unsigned long long t = x0; x0 = x1; x1 = t; }
      } else {
        const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, x0, stride), p1 = __shfl_xor_sync(0xffffffffu, x1, stride);
        const bool lower = (lane & stride) == 0;
        const bool desc0 = size == 64 ? true : (size == 32 ? true : (lane & size) == 0);    // element index lane
        const bool desc1 = size == 64 ? true : (size == 32 ? false : (lane & size) == 0);   // element index lane + 32
        merge_cx(x0, p0, lower == desc0);
        merge_cx(x1, p1, lower == desc1);
      }
    }
  }
}

__global__ void __launch_bounds__(256) rescore_sort_kernel(const float* __restrict__ rescored, const int32_t* __restrict__ list_idx,
                                                           int64_t n_users, int k, int64_t item_base, int64_t n_rows,
                                                           float* __restrict__ out_scores, int32_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= n_users) return;
  const int32_t i0 = list_idx[u * 64 + lane], i1 = list_idx[u * 64 + 32 + lane];
  const bool ok0 = i0 >= 0 && i0 >= item_base && i0 < item_base + n_rows, ok1 = i1 >= 0 && i1 >= item_base && i1 < item_base + n_rows;
  unsigned long long x0 = ok0 ? pxr_key(rescored[u * 64 + lane], (uint32_t)i0) : 0ull;
  unsigned long long x1 = ok1 ? pxr_key(rescored[u * 64 + 32 + lane], (uint32_t)i1) : 0ull;
  sort64_desc(x0, x1, lane);
  if (lane < k) { out_scores[u * k + lane] = x0 ? pxr_key_score(x0) : -INFINITY; out_idx[u * k + lane] = x0 ? (int32_t)pxr_key_idx(x0) : -1; }
  if (lane + 32 < k) { out_scores[u * k + lane + 32] = x1 ? pxr_key_score(x1) : -INFINITY; out_idx[u * k + lane + 32] = x1 ? (int32_t)pxr_key_idx(x1) : -1; }
}

// Lists longer than 64 (top_k > 64: the fused kernel is run once per 64-slot page, score_tc.cu): a block per user sorts the
// L <= 1 024 keys (score, ~index) in shared memory with a bitonic network and writes the first k.
__global__ void __launch_bounds__(256) rescore_sort_block_kernel(const float* __restrict__ rescored, const int32_t* __restrict__ list_idx,
                                                                 int L, int NP2, int k, int64_t item_base, int64_t n_rows,
                                                                 float* __restrict__ out_scores, int32_t* __restrict__ out_idx) {
  __shared__ unsigned long long keys[1024];
  const int64_t u = blockIdx.x;
  for (int j = threadIdx.x; j < NP2; j += blockDim.x) {
    unsigned long long x = 0ull;
    if (j < L) {
      const int32_t gi = list_idx[u * L + j];
      if (gi >= 0 && gi >= item_base && gi < item_base + n_rows) x = pxr_key(rescored[u * L + j], (uint32_t)gi);
    }
    keys[j] = x;
  }
  __syncthreads();
  for (int size = 2; size <= NP2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < NP2 / 2; t += blockDim.x) {
        const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1)), j = i | stride;      // i < j, the pair of this step
        const bool desc = (i & size) == 0 || size == NP2;
        const unsigned long long a = keys[i], b = keys[j];
        if ((a < b) == desc) { keys[i] = b; keys[j] = a; }
      }
      __syncthreads();
    }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const unsigned long long x = j < NP2 ? keys[j] : 0ull;
    out_scores[u * k + j] = x ? pxr_key_score(x) : -INFINITY;
    out_idx[u * k + j] = x ? (int32_t)pxr_key_idx(x) : -1;
  }
}

// scratch per candidate: pair user (8) + pair row (8) + re-scored value (4)
size_t pxr_rescore_list_bytes(int64_t n_users, int32_t list_len) { return pxr_align_up((size_t)n_users * list_len * (8 + 8 + 4), 256); }

// list_idx: (n_users, L) candidate item indices (global, -1 padded), L = 64 or a multiple of 64 up to 1 024; k <= L
int pxr_launch_rescore(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                       const int32_t* list_idx, int32_t L, int32_t k, float* out_scores, int32_t* out_idx, void* ws, cudaStream_t st) {
  if (n_users == 0) return PXR_OK;
  if (L < 64 || L > 1024 || L % 64 || k > L) PXR_FAIL(h, PXR_ERR_INVALID, "re-score: list length %d / k %d not supported", L, k);
  const int64_t n_pairs = n_users * L;
  int64_t* pair_user = (int64_t*)ws;
  int64_t* pair_row = pair_user + n_pairs;
  float* resc = (float*)(pair_row + n_pairs);
  rescore_prep_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(user_idx, list_idx, n_pairs, L, h->item_base, h->n_rows, pair_user, pair_row);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  const int rc = pxr_launch_score_simt(h, user_embedding, pair_user, pair_row, n_pairs, resc, nullptr, st);
  if (rc) return rc;
  if (L == 64) {
    rescore_sort_kernel<<<(unsigned)((n_users + 7) / 8), 256, 0, st>>>(resc, list_idx, n_users, k, h->item_base, h->n_rows, out_scores, out_idx);
  } else {
    int np2 = 128;
    while (np2 < L) np2 <<= 1;
    rescore_sort_block_kernel<<<(unsigned)n_users, 256, 0, st>>>(resc, list_idx, L, np2, k, h->item_base, h->n_rows, out_scores, out_idx);
  }
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

// ===========================================================================
// K5: ranking metrics (reference tasks.py:567-635, 718-747; metrics.py:63-100)
// ===========================================================================
#define METRIC_COLS 9      // precision, recall, f1, hit_rate, ndcg, mrr, ndcg (metrics.py), precision hits/k (metrics.py:35), MAP (metrics.py:102-133)
#define METRIC_THREADS 128

struct MetricKs { int n; int k[PXR_MAX_KS]; };

__global__ void __launch_bounds__(METRIC_THREADS) metrics_user_kernel(const int32_t* __restrict__ topk, int k_stride,
                                                                       int64_t n_users, const int64_t* __restrict__ gt_indptr,
                                                                       const int32_t* __restrict__ gt_idx,
                                                                       const int32_t* __restrict__ recall_den, MetricKs ks,
                                                                       const double* __restrict__ discount,
                                                                       const double* __restrict__ ideal,
                                                                       double* __restrict__ block_sums) {
  __shared__ double red[METRIC_THREADS];
  const int64_t u = (int64_t)blockIdx.x * METRIC_THREADS + threadIdx.x;
  double vals[PXR_MAX_KS][METRIC_COLS];
  for (int a = 0; a < ks.n; ++a) for (int c = 0; c < METRIC_COLS; ++c) vals[a][c] = 0.0;
  if (u < n_users) {
    const int64_t g0 = gt_indptr[u], g1 = gt_indptr[u + 1];
    const int npos = (int)(g1 - g0);
    if (npos > 0) {                                   // "if not pos_set: continue" (tasks.py:589-591)
      const int32_t* rec = topk + u * k_stride;
      int hits = 0, nrec = 0, first = 0, ki = 0;
      double dcg = 0.0, apsum = 0.0;
      const int rden = recall_den ? recall_den[u] : npos;          // len(positive_items): raw rows, tasks.py:579
      for (int j = 0; j < k_stride && ki < ks.n; ++j) {
        const int32_t it = rec[j];
        if (it >= 0) {
          nrec++;
          bool hit = false;
          for (int64_t g = g0; g < g1; ++g) if (gt_idx[g] == it) { hit = true; break; }
          if (hit) { hits++; dcg += discount[j]; apsum += (double)hits / (double)(j + 1); if (!first) first = j + 1; }
        }
        while (ki < ks.n && j + 1 == ks.k[ki]) {     // cut-offs ascending
          const int k = ks.k[ki];
          const double prec = nrec > 0 ? (double)hits / (double)nrec : 0.0;   // denominator len(recs), tasks.py:577
          const double rec_ = rden > 0 ? (double)hits / (double)rden : 0.0;
          const double f1 = (prec + rec_) > 0.0 ? 2.0 * prec * rec_ / (prec + rec_) : 0.0;
          const double idcg = ideal[npos < k ? npos : k];
          vals[ki][0] = prec; vals[ki][1] = rec_; vals[ki][2] = f1; vals[ki][3] = hits > 0 ? 1.0 : 0.0;
          vals[ki][4] = idcg > 0.0 ? dcg / idcg : 0.0;
          vals[ki][5] = first ? 1.0 / (double)first : 0.0;
          vals[ki][6] = hits > 0 ? dcg / ideal[hits] : 0.0;
          vals[ki][7] = nrec > 0 ? (double)hits / (double)k : 0.0;           // metrics.py:29-35
          vals[ki][8] = hits > 0 ? apsum / (double)npos : 0.0;               // metrics.py:119-133 on the first k entries
          ki++;
        }
      }
    }
  }
  for (int a = 0; a < ks.n; ++a) {
    for (int c = 0; c < METRIC_COLS; ++c) {
      red[threadIdx.x] = vals[a][c];
      __syncthreads();
      for (int s = METRIC_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
      }
      if (threadIdx.x == 0) block_sums[(int64_t)blockIdx.x * (PXR_MAX_KS * METRIC_COLS) + a * METRIC_COLS + c] = red[0];
      __syncthreads();
    }
  }
}

// Fast variant for lists of at most 64 entries.  A warp owns a contiguous range of users (a multiple of 32, so every
// group of 32 lists starts 16-byte aligned) and takes them 32 at a time.  Their lists are one contiguous run of 32 k
// entries: the lanes stream it with coalesced 16-byte loads, ALL of a group's loads in flight at once (VB = ceil(k / 4)
// per lane: 6.4 KB per warp at k = 50) together with the first positive of each user and the next group's CSR offsets --
// the stage is bound by the chain of memory latencies per group, not by instructions (~10 warp instructions per user)
// nor by bytes in flight alone (ncu on the previous, batched version: 146 us for 223 MB with 1 - 2 KB in flight per warp and
// three dependent round trips per group).  The run is parked in shared memory (conflict-free 16-byte stores) and read back
// transposed: lane t scans the list of user ub + t against that user's positives (first one in a register) and builds the
// user's 64-bit hit and valid masks in registers -- no shuffles, no atomics.  Only users with a hit have non-zero
// metrics: those lanes then do the float64 arithmetic (every cut-off is a popcount of the hit mask; the discounted gain
// is summed left to right over the hit positions, the order of the reference loop, tasks.py:733-747); everybody else adds
// exact zeros, i.e. nothing.  Lane partial sums are reduced in a fixed order; blocks / warps own fixed user ranges =>
// deterministic result.
#define METRIC_WARPS 4
struct HitUser { unsigned long long hit, valid; int npos, rden; };

// the accuracy block of one user with at least one hit: adds its nine columns per cut-off to this lane's partial sums
template <int NKS>
__device__ __forceinline__ void hit_user_metrics(const HitUser& e, const MetricKs& ks, const double* __restrict__ discount,
                                                 const double* __restrict__ ideal, double (&sums)[NKS][METRIC_COLS]) {
  const unsigned long long hit = e.hit, valid = e.valid;
  const int npos = e.npos, rden = e.rden;
  const int first = __ffsll((long long)hit);
  double dcg = 0.0, apsum = 0.0;
  int nh = 0;
  unsigned long long rest = hit;
#pragma unroll
  for (int a = 0; a < NKS; ++a) {
    if (a < ks.n) {
      const int k = ks.k[a];
      const unsigned long long km = k >= 64 ? ~0ull : ((1ull << k) - 1ull);
      while (rest) {                            // extend the left-to-right sum to the hits below this cut-off
        const int j = __ffsll((long long)rest) - 1;
        if (j >= k) break;
        dcg += discount[j];
        apsum += (double)(++nh) / (double)(j + 1);
        rest &= rest - 1;
      }
      const int hits = __popcll(hit & km), nrec = __popcll(valid & km);
      const double prec = nrec > 0 ? (double)hits / (double)nrec : 0.0;   // denominator len(recs), tasks.py:577
      const double rec_ = rden > 0 ? (double)hits / (double)rden : 0.0;
      const double f1 = (prec + rec_) > 0.0 ? 2.0 * prec * rec_ / (prec + rec_) : 0.0;
      const double idcg = ideal[npos < k ? npos : k];
      sums[a][0] += prec; sums[a][1] += rec_; sums[a][2] += f1; sums[a][3] += hits > 0 ? 1.0 : 0.0;
      sums[a][4] += idcg > 0.0 ? dcg / idcg : 0.0;
      sums[a][5] += (first && first <= k) ? 1.0 / (double)first : 0.0;
      sums[a][6] += hits > 0 ? dcg / ideal[hits] : 0.0;
      sums[a][7] += nrec > 0 ? (double)hits / (double)k : 0.0;            // metrics.py:29-35
      sums[a][8] += hits > 0 ? apsum / (double)npos : 0.0;                // metrics.py:119-133 on the first k entries
    }
  }
}

// NKS: cut-offs (2 covers the usual @10 / @50; 8 = PXR_MAX_KS).  VEC: 16-byte loads (aligned lists), VB = loads per lane and batch
template <int NKS, bool VEC, int VB>
__global__ void __launch_bounds__(32 * METRIC_WARPS, (NKS <= 2 ? 4 : 2)) metrics_warp_kernel(const int32_t* __restrict__ topk, int k_stride,
                                                                         int64_t n_users, int64_t users_per_warp,
                                                                         const int64_t* __restrict__ gt_indptr,
                                                                         const int32_t* __restrict__ gt_idx,
                                                                         const int32_t* __restrict__ recall_den, MetricKs ks,
                                                                         const double* __restrict__ discount,
                                                                         const double* __restrict__ ideal,
                                                                         double* __restrict__ block_sums) {
  __shared__ double acc[METRIC_WARPS][PXR_MAX_KS * METRIC_COLS];
  __shared__ __align__(16) int32_t runs[METRIC_WARPS][32 * 64];      // the 32 lists of the warp's current group
  __shared__ HitUser hitq[METRIC_WARPS][64];                          // users with a hit waiting for a full-warp float64 pass
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double sums[NKS][METRIC_COLS];
#pragma unroll
  for (int a = 0; a < NKS; ++a)
#pragma unroll
    for (int c = 0; c < METRIC_COLS; ++c) sums[a][c] = 0.0;
  const int64_t w = (int64_t)blockIdx.x * METRIC_WARPS + warp;
  const int64_t u0 = w * users_per_warp, u1 = min(n_users, u0 + users_per_warp);
  int32_t* run = runs[warp];
  HitUser* hq = hitq[warp];
  int qn = 0;                                          // queued users (warp-uniform)
  constexpr int EPL = VEC ? 4 : 1;                     // entries per load
  int64_t g0n = 0; int nposn = 0;                      // lane t: CSR offsets of user ub + t, fetched one group ahead
  if (u0 + lane < u1) { g0n = gt_indptr[u0 + lane]; nposn = (int)(gt_indptr[u0 + lane + 1] - g0n); }
  for (int64_t ub = u0; ub < u1; ub += 32) {
    const int nb = (int)min((int64_t)32, u1 - ub);
    const int n_el = nb * k_stride;
    const int32_t* src = topk + ub * k_stride;
    // everything this group needs from memory is issued together: its list entries, the first positive of each user,
    // and the CSR offsets of the next group
    const int64_t g0 = g0n;                            // lane t: positives of user ub + t
    const int npos = nposn;
    int32_t pq0 = INT_MIN;                             // never equals a valid entry (>= 0)
    for (int base = 0; base < n_el; base += 32 * VB * EPL) {
      int32_t v[VB][EPL];
#pragma unroll
      for (int m = 0; m < VB; ++m) {
        const int e = base + EPL * (lane + 32 * m);
        if (VEC) {
          if (e + 3 < n_el) {
            const int4 q4 = __ldg(reinterpret_cast<const int4*>(src + e));
            v[m][0] = q4.x; v[m][1 % EPL] = q4.y; v[m][2 % EPL] = q4.z; v[m][3 % EPL] = q4.w;
          } else {
#pragma unroll
            for (int c = 0; c < EPL; ++c) v[m][c] = (e + c < n_el) ? __ldg(src + e + c) : -1;
          }
        } else {
          v[m][0] = e < n_el ? __ldg(src + e) : -1;
        }
      }
      if (base == 0) {
        if (npos > 0) pq0 = __ldg(gt_idx + g0);
        g0n = 0; nposn = 0;
        if (ub + 32 + lane < u1) { g0n = gt_indptr[ub + 32 + lane]; nposn = (int)(gt_indptr[ub + 32 + lane + 1] - g0n); }
      }
#pragma unroll
      for (int m = 0; m < VB; ++m) {
        const int e = base + EPL * (lane + 32 * m);
        if (VEC) { if (e < 32 * 64) *reinterpret_cast<int4*>(run + e) = make_int4(v[m][0], v[m][1 % EPL], v[m][2 % EPL], v[m][3 % EPL]); }
        else if (e < 32 * 64) run[e] = v[m][0];
      }
    }
    __syncwarp();
    unsigned long long hit = 0ull, valid = 0ull;
    if (lane < nb) {
      // eight entries at a time into 8-bit masks (constant shifts), then one 64-bit shift per chunk; the last chunk may
      // run into the next user's entries (inside the 32 x 64 buffer for every k <= 64): masked off below
      const int32_t* mine = run + lane * k_stride;
      for (int j0 = 0; j0 < k_stride; j0 += 8) {
        uint32_t h8 = 0u, v8 = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int32_t x = mine[j0 + i];
          v8 |= (uint32_t)(x >= 0) << i;
          h8 |= (uint32_t)(x == pq0) << i;                      // pq0 = INT_MIN without positives
        }
        valid |= (unsigned long long)v8 << j0;
        hit |= (unsigned long long)h8 << j0;
      }
      for (int q = 1; q < npos; ++q) {                          // further positives of this user (leave-one-out: none)
        const int32_t pq = __ldg(gt_idx + g0 + q);
        for (int j0 = 0; j0 < k_stride; j0 += 8) {
          uint32_t h8 = 0u;
#pragma unroll
          for (int i = 0; i < 8; ++i) h8 |= (uint32_t)(mine[j0 + i] == pq) << i;
          hit |= (unsigned long long)h8 << j0;
        }
      }
      valid &= k_stride >= 64 ? ~0ull : ((1ull << k_stride) - 1ull);
      hit &= valid;
    }
    __syncwarp();                                     // the run may be overwritten by the next group's stores
    // Users with a hit (a user without one, or without positives, adds exact zeros: tasks.py:589-591) are queued per warp
    // and their float64 arithmetic runs when 32 of them are waiting: a full warp per pass instead of the few divergent
    // hit lanes of every group (20 % of the users with a hit: the pass runs once per ~5 groups, not once per group)
    {
      const unsigned hm = __ballot_sync(0xffffffffu, hit != 0ull);
      if (hit) {
        HitUser& e = hq[qn + __popc(hm & ((1u << lane) - 1u))];
        e.hit = hit; e.valid = valid; e.npos = npos;
        e.rden = recall_den ? __ldg(recall_den + ub + lane) : npos;           // len(positive_items): raw rows, tasks.py:579
      }
      qn += __popc(hm);
      __syncwarp();
      if (qn >= 32) {
        qn -= 32;
        hit_user_metrics<NKS>(hq[qn + lane], ks, discount, ideal, sums);
        __syncwarp();
      }
    }
  }
  if (lane < qn) hit_user_metrics<NKS>(hq[lane], ks, discount, ideal, sums);       // the users still waiting
  for (int i = lane; i < PXR_MAX_KS * METRIC_COLS; i += 32) acc[warp][i] = 0.0;      // cut-offs beyond NKS / ks.n
  __syncwarp();
#pragma unroll
  for (int a = 0; a < NKS; ++a)
#pragma unroll
    for (int c = 0; c < METRIC_COLS; ++c) {
      double v = a < ks.n ? sums[a][c] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) acc[warp][a * METRIC_COLS + c] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < PXR_MAX_KS * METRIC_COLS; i += blockDim.x) {
    double t = 0.0;
    for (int ww = 0; ww < METRIC_WARPS; ++ww) t += acc[ww][i];
    block_sums[(int64_t)blockIdx.x * (PXR_MAX_KS * METRIC_COLS) + i] = t;
  }
}

#define METRIC_FINAL_THREADS 256
__global__ void __launch_bounds__(METRIC_FINAL_THREADS) metrics_final_kernel(const double* __restrict__ block_sums, int64_t n_blocks, int n_ks,
                                                                            double* out) {
  // one block per (k, column): thread i sums blocks i, i + 256, ... in order, then a fixed-shape tree => deterministic
  __shared__ double part[METRIC_FINAL_THREADS];
  const int t = blockIdx.x;
  if (t >= n_ks * METRIC_COLS) return;
  double s = 0.0;
  for (int64_t b = threadIdx.x; b < n_blocks; b += METRIC_FINAL_THREADS) s += block_sums[b * (PXR_MAX_KS * METRIC_COLS) + t];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = METRIC_FINAL_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[t] = part[0];
}

int pxr_launch_metrics(const int32_t* topk_idx, int32_t k_stride, int64_t n_users, const int64_t* gt_indptr,
                       const int32_t* gt_idx, const int32_t* recall_den, const int32_t* ks, int32_t n_ks, const double* discount,
                       const double* ideal, double* out_sums, void* ws, cudaStream_t st) {
  MetricKs mk; mk.n = n_ks;
  for (int i = 0; i < n_ks; ++i) { mk.k[i] = ks[i]; if (i && ks[i] <= ks[i - 1]) return PXR_ERR_INVALID; if (ks[i] > k_stride || ks[i] <= 0) return PXR_ERR_INVALID; }
  int64_t blocks = (n_users + METRIC_THREADS - 1) / METRIC_THREADS;
  if (blocks == 0) { cudaMemsetAsync(out_sums, 0, sizeof(double) * n_ks * METRIC_COLS, st); return PXR_OK; }
  if (k_stride <= 64) {
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    // one wave of resident blocks (4 per SM for <= 2 cut-offs, 2 otherwise): a warp then walks many groups of 32 users and
    // its fixed costs (the float64 reduction of its partial sums) are paid once; never more than the workspace holds
    blocks = std::min<int64_t>(blocks, (int64_t)n_sm * (n_ks <= 2 ? 4 : 2));
    const int64_t warps = blocks * METRIC_WARPS;
    const int64_t upw = ((n_users + warps - 1) / warps + 31) / 32 * 32;      // a multiple of 32: every group of lists starts 16-byte aligned
    const bool vec = (reinterpret_cast<uintptr_t>(topk_idx) & 15) == 0;
    // aligned lists: one batch of ceil(k / 4) 16-byte loads per lane covers a whole group of 32 lists; otherwise scalar loads, 8 per batch
    // 35 KB of static shared memory per block: ask for the large carve-out, or the default split leaves one block per SM
#define PXR_METRICS_LAUNCH(NK, V, VB) do { cudaFuncSetAttribute(metrics_warp_kernel<NK, V, VB>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared); \
      metrics_warp_kernel<NK, V, VB><<<(unsigned)blocks, 32 * METRIC_WARPS, 0, st>>>(topk_idx, k_stride, n_users, upw, gt_indptr, gt_idx, recall_den, mk, discount, ideal, (double*)ws); } while (0)
#define PXR_METRICS_VB(NK) do { if (!vec) PXR_METRICS_LAUNCH(NK, false, 8); else if (k_stride <= 16) PXR_METRICS_LAUNCH(NK, true, 4); \
      else if (k_stride <= 32) PXR_METRICS_LAUNCH(NK, true, 8); else if (k_stride <= 52) PXR_METRICS_LAUNCH(NK, true, 13); else PXR_METRICS_LAUNCH(NK, true, 16); } while (0)
    if (n_ks <= 2) PXR_METRICS_VB(2); else PXR_METRICS_VB(PXR_MAX_KS);
#undef PXR_METRICS_VB
#undef PXR_METRICS_LAUNCH
  } else {
    metrics_user_kernel<<<(unsigned)blocks, METRIC_THREADS, 0, st>>>(topk_idx, k_stride, n_users, gt_indptr, gt_idx, recall_den, mk,
                                                                     discount, ideal, (double*)ws);
  }
  metrics_final_kernel<<<n_ks * METRIC_COLS, METRIC_FINAL_THREADS, 0, st>>>((const double*)ws, blocks, n_ks, out_sums);
  return cudaGetLastError() == cudaSuccess ? PXR_OK : PXR_ERR_CUDA;
}
